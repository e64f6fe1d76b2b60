#!/bin/bash
# BASELINE config 5 (512 x 30 s) and the reproducible-noise variant with the final build
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3n
timeout 900 python bench.py --seconds 30 --batch-per-gpu 512 --steps 20 --warmup 3 --no-cpu-baseline --no-noise-variant > gpurun_out/bench_c5_g1_$TAG.json 2> gpurun_out/bench_c5_g1_$TAG.err; echo "config 5 exit $?"
timeout 600 python bench.py --reproducible --steps 50 --no-cpu-baseline --no-noise-variant > gpurun_out/bench_repro_$TAG.json 2> gpurun_out/bench_repro_$TAG.err; echo "reproducible exit $?"
python - <<PY
import json
for f in ("bench_c5_g1_$TAG", "bench_repro_$TAG"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_ok"))
        ob = d["parity"]["oracle_batch"]
        print("   voices<=1e-4", ob["voices_le_1e-4"], "of", ob["sounds"], "median", ob["audio_median"], "max", ob["audio_max"], "pqmf", ob["pqmf_rel"], "loss e2e", ob["loss4_rel_end_to_end"])
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -2 gpurun_out/bench_c5_g1_$TAG.err
