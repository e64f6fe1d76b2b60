#!/usr/bin/env python
"""Instruction histogram of the kernels of an object whose mangled name matches a regex.
Usage: python tools/sass_hist.py <object-or-so> <name-regex> [--dump]"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], re.compile(sys.argv[2])
    dump = "--dump" in sys.argv
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        if not pat.search(name):
            continue
        mix = collections.Counter()
        n = 0
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*([^;]*);", f):
            mix[m.group(3).split(".")[0]] += 1
            n += 1
            if dump:
                print(m.group(1), m.group(2) or "", m.group(3), m.group(4))
        print(name[:150])
        print(" ", n, "instructions:", " ".join(f"{k}:{v}" for k, v in mix.most_common(45)))


if __name__ == "__main__":
    main()
