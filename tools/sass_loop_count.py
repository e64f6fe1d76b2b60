#!/usr/bin/env python
"""Instruction mix of the innermost tile loop of every k_voice_audio instantiation in an object file.

Usage: python tools/sass_loop_count.py <object-or-so> [kernel-name-regex]
The tile loop is taken to be the longest backward branch; counts are per loop body (= one tile of SPT samples/thread).
"""
import collections
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"k_voice_audio")
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if not pat.search(name):
            continue
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        ins = []
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*([^;]*);", f):
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
        best = None
        for addr, op, args in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", args)
                if t:
                    tgt = int(t.group(1), 16)
                    if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                        # ignore the outer persistent loop (branches back to the queue pop near the top)
                        if tgt > 0x200:
                            best = (tgt, addr)
        if not best:
            continue
        body = [i for i in ins if best[0] <= i[0] <= best[1]]
        mix = collections.Counter()
        for _, op, _ in body:
            base = op.split(".")[0]
            mix[base] += 1
        spt = re.search(r"k_voice_audio<\(int\)(\d+), \(int\)(\d+)", dem)
        per = int(spt.group(2)) if spt else 1
        print(f"{dem[:110]}\n  loop 0x{best[0]:x}-0x{best[1]:x}: {len(body)} instr/tile = {len(body)/per:.1f}/sample")
        print("  " + " ".join(f"{k}:{v}" for k, v in mix.most_common(40)))


if __name__ == "__main__":
    main()
