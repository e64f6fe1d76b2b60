#!/usr/bin/env python
"""Issue-side roofline model of k_voice_audio ("report min(HBM, issue) honestly") -> profiles/issue_model.json.

The kernel is bound on the instruction side, not by HBM.  This tool counts the SASS instructions of the
kernel's tile loop on the path 85 % of the voices take (no MIDI clamp; an object compiled with
-DIAS_AUDIO_COUNT_NOCLAMP_ONLY so the loop holds one pitch-pass variant, for counting only) and turns the mix into cycles
of every execution pipe of a scheduler (one quarter of an SM), with the occupancies measured by
tools/micro/dispatch_mix.cu on B200 (profiles/dispatch_mix_r2.log):

    FMA pipe   packed fp32 (FFMA2 / FADD2 / FMUL2) 2 cycles (8 FFMA2 take 16.3 cycles: exactly two FFMA), scalar fp32 1
    XU pipe    MUFU, F2F, F2I, I2F 8 cycles (4 lanes per cycle; 4 MUFU take 32 cycles, a f32->f64->f32 round trip 17)
    ALU pipe   FSEL, FMNMX, LEA, ISETP, IADD3, LOP3, SHF, MOV ... 2 cycles (half rate)
    FP64 pipe  DADD 2 cycles (4 DADD take 8.1 cycles)
    dispatch   every instruction 1 cycle

The pipes run concurrently (FFMA2 x8 + DADD x4: 16.5 cycles; + MUFU x2: 17.9; + LEA x4: 17.6), so the bound of the tile
loop is the busiest pipe -- the FMA pipe.  bench.py divides that bound by what the schedulers offer during the measured
launch time: roofline.issue.achieved_frac (= FMA-pipe utilisation the kernel would need at minimum / time it takes).

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
KERNEL = r"k_voice_audio_spILi128ELi16ELi4ELb0"  # the default audio kernel (software-pipelined)
SPT = 16
PACKED = {"FFMA2", "FADD2", "FMUL2"}
SCALAR_FP32 = {"FFMA", "FADD", "FMUL"}
XU = {"MUFU", "F2F", "I2F", "F2I", "I2FP", "FRND"}
FP64 = {"DADD", "DFMA", "DMUL"}
ALU = {"FSEL", "FMNMX", "FMNMX3", "LEA", "ISETP", "FSETP", "IADD3", "LOP3", "SHF", "MOV", "IMAD", "SEL", "VIMNMX", "PRMT",
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
    pipes = {
        "fma": sum(2 * n for k, n in mix.items() if k in PACKED) + sum(n for k, n in mix.items() if k in SCALAR_FP32),
        "xu": sum(8 * n for k, n in mix.items() if k in XU),
        "alu": sum(2 * n for k, n in mix.items() if k in ALU),
        "fp64": sum(2 * n for k, n in mix.items() if k in FP64),
        "dispatch": len(body),
    }
    bound = max(pipes, key=pipes.get)
    out = {
        "kernel": "k_voice_audio_sp<128,16,4> tile loop (pass 2 of tile t + pass 1 of tile t+1), no-clamp path (85 % of voices)",
        "src_sha256": bench.source_hash(),
        "warp_instructions_per_tile": len(body),
        "instr_per_sample": len(body) / SPT,
        "pipe_cycles_per_sample": {k: v / SPT for k, v in pipes.items()},
        "bound_pipe": bound,
        "bound_cycles_per_sample": pipes[bound] / SPT,
        "mix_per_tile": dict(mix.most_common()),
        "note": "per warp and sample (one warp instruction covers 32 samples' worth of one per-sample operation), cycles of one "
                "scheduler's pipes; occupancies from tools/micro/dispatch_mix.cu; per-tile overheads (scan, barrier, loads) are "
                "included, the per-voice prologue, the silent-tail fill and the normalise pass are not",
    }
    path = os.path.join(ROOT, "profiles", "issue_model.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("src_sha256", "instr_per_sample", "pipe_cycles_per_sample", "bound_pipe")}))


if __name__ == "__main__":
    main()
