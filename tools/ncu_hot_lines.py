#!/usr/bin/env python
"""Hot source lines of one ncu report: python tools/ncu_hot_lines.py <report.ncu-rep> [top-n]
Reads `ncu --page source --print-source cuda,sass` (needs -lineinfo + --import-source on) and prints, per source line,
its share of executed warp instructions and of stall samples."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname, hdr, out = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if len(r) >= 2 and r[0] == "File Path":
            fname, hdr = r[1].split("/")[-1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr) or not r[0]:
            continue
        d = dict(zip(hdr, r))
        try:
            out.append((int(d["Instructions Executed"]), int(d["# Samples"]), fname, r[0], r[1].strip()[:110]))
        except (ValueError, KeyError):
            continue
    tot, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print(f"total warp instructions {tot}, samples {ts}")
    for o in sorted(out, reverse=True)[:top]:
        print(f"{o[0] / tot * 100:5.1f}% instr {o[1] / ts * 100:5.1f}% samples  {o[2]}:{o[3]}  {o[4]}")


if __name__ == "__main__":
    main()
