#!/usr/bin/env python
"""Issue-side roofline model of k_voice_audio ("report min(HBM, issue) honestly") -> profiles/issue_model.json.

The kernel is bound by the dispatch port of the four warp schedulers per SM, not by HBM.  This tool
counts the SASS instructions of the kernel's tile loop on the path 85 % of the voices take (no MIDI clamp; an object
compiled with -DIAS_AUDIO_COUNT_NOCLAMP_ONLY so the loop holds one pitch-pass variant, for counting only) and prices
them in dispatch cycles per warp instruction with the costs measured by tools/micro/ffma2_bench.cu on B200:

    packed fp32 (FFMA2 / FADD2 / FMUL2)  2      half-rate ALU pipe (FSEL, FMNMX, LEA, ISETP, IADD3, LOP3, SHF, MOV ...)  2
    FP64 (DADD)                          2      XU (MUFU, F2F, I2F, F2I)                                                  2
    scalar fp32, loads / stores, shuffles, branches, uniform-datapath instructions                                     1

bench.py reads the result (tied to the source hash) and divides the modelled cycles by what SMs x 4 schedulers x clock
offer during the measured launch time: roofline.issue.achieved_frac.

    python tools/issue_model.py            # cross-compiles voice.cu to a temporary object, no GPU needed
"""
import collections
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "inverse-audio-synthesis_b200", "csrc")
KERNEL = r"k_voice_audioILi128ELi16ELi4ELb1ELb0"
SPT = 16
TWO = {"FFMA2", "FADD2", "FMUL2", "DADD", "DFMA", "DMUL", "MUFU", "F2F", "I2F", "F2I", "I2FP", "FRND",
       "FSEL", "FMNMX", "FMNMX3", "LEA", "ISETP", "FSETP", "IADD3", "LOP3", "SHF", "MOV", "IMAD", "SEL", "VIMNMX", "PRMT",
       "CS2R", "HFMA2", "PLOP3", "IABS", "VIADD"}


def loop_body(sass_text):
    for f in re.split(r"\n\s*Function : ", sass_text)[1:]:
        name = f.split("\n", 1)[0].strip()
        if not re.search(KERNEL, name):
            continue
        ins = [(int(m.group(1), 16), m.group(3), m.group(4))
               for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*([^;]*);", f)]
        best = None
        for addr, op, args in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", args)
                if t:
                    tgt = int(t.group(1), 16)
                    if 0x200 < tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                        best = (tgt, addr)
        return [i for i in ins if best[0] <= i[0] <= best[1]]
    raise SystemExit("kernel not found")


def main():
    import bench  # source_hash()

    with tempfile.TemporaryDirectory() as tmp:
        obj = os.path.join(tmp, "voice_count.o")
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                        "-Xcompiler", "-fPIC", "-fmad=false", "-DIAS_POW_NOINLINE", "-DIAS_AUDIO_COUNT_NOCLAMP_ONLY", "-c",
                        os.path.join(CSRC, "voice.cu"), "-o", obj], check=True)
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    body = loop_body(txt)
    mix = collections.Counter(op.split(".")[0] for _, op, _ in body)
    cycles = sum(n * (2 if k in TWO else 1) for k, n in mix.items())
    out = {
        "kernel": "k_voice_audio<128,16,4> tile loop, no-clamp path (85 % of voices)",
        "src_sha256": bench.source_hash(),
        "warp_instructions_per_tile": len(body),
        "instr_per_sample": len(body) / SPT,
        "dispatch_cycles_per_sample": cycles / SPT,
        "mix_per_tile": dict(mix.most_common()),
        "two_cycle_classes": sorted(TWO & set(mix)),
        "note": "per warp and sample (one warp instruction covers 32 samples' worth of one per-sample operation); costs from "
                "tools/micro/ffma2_bench.cu; per-tile overheads (scan, barrier, loads) are included, the per-voice prologue, "
                "the silent-tail fill and the normalise pass are not",
    }
    path = os.path.join(ROOT, "profiles", "issue_model.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("src_sha256", "instr_per_sample", "dispatch_cycles_per_sample")}))


if __name__ == "__main__":
    main()
