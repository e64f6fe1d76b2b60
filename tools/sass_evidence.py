#!/usr/bin/env python
"""SASS evidence for the tensor-core / TMA kernels -> profiles/sass_<kernel>.txt (tracked).

    python tools/sass_evidence.py           # cuobjdump -sass of the in-tree libias_b200.so, no GPU needed

For k_gram_tc and k_bwd_tc: the tcgen05 / TMEM / bulk-copy mnemonics B200_PROFILING.md names as proof (UTCHMMA =
tcgen05.mma kind::tf32, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM
alloc / dealloc, SYNCS = mbarrier), with counts and every matching line.  For the packed-fp32 kernels: FFMA2 counts."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "inverse-audio-synthesis_b200", "ias_b200", "libias_b200.so")
PROF = os.path.join(ROOT, "profiles")
TC = re.compile(r"UTCHMMA|UTCQMMA|UTCMMA|LDTM|STTM|UBLKCP|UTCBAR|UTCATOMSWS|SYNCS|UTMALDG|FENCE\.VIEW")
TARGETS = {"k_gram_tc": TC, "k_bwd_tc": TC, "k_voice_audioILi128ELi16ELi4ELb1ELb0": re.compile(r"FFMA2|FADD2|FMUL2|DADD|F2F|MUFU"),
           "k_pqmf_analysisILi3ELi63ELi4ENS0_6TapsCMILi3ELi63EEELb1": re.compile(r"UBLKCP|SYNCS|FFMA2|LDS|STG|LDCU|REDUX|REDG")}


def main():
    import bench

    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for key, pat in TARGETS.items():
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            if key not in name:
                continue
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            lines = [ln.strip() for ln in f.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/", ln) and pat.search(ln)]
            cnt = collections.Counter(pat.search(ln).group(0) for ln in lines)
            total = len(re.findall(r"/\*[0-9a-f]{4}\*/\s+", f))
            short = re.sub(r"[^a-z0-9_]", "", key.split("IL")[0])
            out = [f"# cuobjdump -sass libias_b200.so ({os.path.basename(LIB)}, sources {bench.source_hash()}): {dem[:160]}",
                   f"# {total} instructions; matched mnemonics: " + ", ".join(f"{k} x{v}" for k, v in cnt.most_common()), ""]
            out += lines if key.startswith(("k_gram", "k_bwd")) else lines[:40] + [f"... ({len(lines)} matching lines)"]
            path = os.path.join(PROF, f"sass_{short}.txt")
            open(path, "w").write("\n".join(out) + "\n")
            print(path, dict(cnt))
            break


if __name__ == "__main__":
    main()
