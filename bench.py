#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 front end: 4 s @ 44.1 kHz sounds/sec through synth -> PQMF -> VICReg.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference-style CPU path (oracle) on host cores

One step = one pass of the hot path over one batch of synthetic sounds per GPU:
    Voice(batch_idx) [seed 78 params/sound on device -> control-rate stage -> audio render]
      -> PQMF(N=3).analysis -> harness bridge (abs-mean pool to 256 + fixed projections; NOT a reference component)
      -> all-gather of the [B_local, 2*256] embeddings (N > 1) -> VICReg loss (D = 256, tcgen05 Gram).
Workload: BASELINE.json configs[3] per-GPU shard -- 1024 sounds per GPU (global batch 1024*N; 8192 at N = 8), weak
scaling.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is derived.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import torch

import harness  # sets sys.path for the package and the oracle

T_4S = 176400
D = harness.EMBED_DIM
ALGO_BYTES_PER_SOUND = 12 * T_4S + 8 * D  # synth writes 4T, PQMF reads 4T + writes 4T, loss reads 8D


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=1024)
    ap.add_argument("--bands", type=int, default=3)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="e2e leg: eager launches instead of a CUDA graph replay")
    ap.add_argument("--no-pipeline", dest="pipeline", action="store_false",
                    help="by default a step renders the batch whose control stage the previous step prepared and, on a "
                         "side stream under this batch's PQMF / exchange / loss, seeds and prepares the next batch "
                         "(Voice.prepare: seed -> ADSR -> control -> schedule; same kernels, bit-identical results, "
                         "tests/test_gpu_e2e.py); --no-pipeline runs every stage of a batch back to back "
                         "(measured on B200: 1.249 vs 1.223 ms/step)")
    ap.set_defaults(pipeline=True)
    ap.add_argument("--pipeline-depth", type=int, default=1, choices=[1, 2],
                    help="1 (default): the whole control stage of batch k+1 under PQMF / bridge / loss of batch k; 2: the "
                         "control stage in two halves -- LFO / modulation / records of batch k+1 under the PQMF analysis of "
                         "batch k, then seeding + ADSR envelopes of batch k+2 under the bridge / loss kernels of batch k "
                         "(bit-identical, measured slower: 1.133 vs 1.116 ms/step, the envelopes end up on the critical "
                         "path of the step's join)")
    ap.add_argument("--gather", default="stats", choices=["stats", "peer", "nccl"],
                    help="N>1 embedding exchange: stats = every rank reduces its own rows and pushes a 0.4 MB summary "
                         "(mean, second moments, Gram) to its peers over NVLink from inside the loss kernels; peer = the "
                         "statistics kernel reads every peer's embeddings over NVLink (symmetric memory); nccl = "
                         "FullGatherLayer all-gather through torch.distributed (NCCL)")
    ap.add_argument("--cpu-sample", type=int, default=128,
                    help="sounds per CPU chunk (BASELINE configs[0]: 128); the reference arm renders --batch-per-gpu "
                         "sounds per step in chunks of this size, the in-run cpu_baseline leg one chunk per step")
    ap.add_argument("--reproducible", action="store_true",
                    help="SynthConfig(reproducible=True): noise rows repeat with period 32, so the table (22.6 MB) lives in "
                         "L2.  Default is the reference's own setting, reproducible: False (conf/config.yaml:38): a [B,T] "
                         "noise table (722 MB at 1024 x 4 s) that the render streams from HBM every step")
    ap.add_argument("--no-nonreproducible", "--no-noise-variant", dest="no_noise_variant", action="store_true",
                    help="skip the e2e variant with the other noise mode")
    ap.add_argument("--normalize", default="defer", choices=["defer", "kernel"],
                    help="normalize_if_clipping of the rendered audio: defer = Voice(normalize='defer'): the render leaves "
                         "the raw mix and a [B] factor (1/peak of a clipping row, else 1) that PQMF.analysis applies to the "
                         "bands it writes (the filter bank is linear), so the clipping rows (13 %% of the voices) are not "
                         "read and written a second time; kernel = Voice(normalize=True): the audio tensor itself is "
                         "normalised by a second pass inside the render (what Voice.forward returns to a caller that wants "
                         "the audio)")
    ap.add_argument("--main-priority", action="store_true",
                    help="capture the step on a high-priority stream, so that the kernels of the batch in flight (PQMF, "
                         "bridge, loss) are scheduled ahead of the next batch's control stage on the side stream")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (outside the timed region)")
    ap.add_argument("--parity-sounds", type=int, default=0,
                    help="sounds of the parity block's CPU oracle render (default: 128 at 4 s, 16 for long clips)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons while the timed region runs.  NVML from a thread every 5 ms (the driver's
    --steps 20 timed region is ~25 ms: `nvidia-smi -lms 100` sees it once); nvidia-smi only if NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReasons bits
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []      # nvidia-smi: (arrival wall-clock, csv line)
        self.samples = []   # NVML: (wall-clock, sm_mhz, reasons bitmask)
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.t_begin = None
        self.t_end = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)
            self.thread = threading.Thread(target=self._pump_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump_nvml(self):
        pynvml, h = self.nvml
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                     int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while (self.proc is not None or self.nvml is not None) and not (self.rows or self.samples) and \
                time.time() - t0 < timeout:
            time.sleep(0.02)

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        self.stop_flag = True
        lo = (self.t_begin or 0.0)
        if self.nvml is not None:
            hi = (self.t_end or 1e30) + 0.002
            inside = [(m, r) for t, m, r in self.samples if lo <= t <= hi] or [(m, r) for _, m, r in self.samples[-3:]]
            sm = sorted(m for m, _ in inside)
            mask = 0
            for _, r in inside:
                mask |= r
            reasons = sorted(k for k, b in self.BITS.items() if mask & b)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms period, samples inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        hi = (self.t_end or 1e30) + 0.15  # a sample describes the ~100 ms before it arrives
        rows = [r for t, r in self.rows if lo <= t <= hi] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ----------------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle on the host cores)
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_rate(chunk: int, bands: int, seconds: float, steps: int, warmup: int, sounds_per_step: int = 0,
                  reproducible: bool = False):
    """Oracle front end on the host cores.  One step = `sounds_per_step` sounds (default: one chunk) rendered and
    filtered in chunks of `chunk` sounds (bounds the host memory: the torch restatement of the synth keeps ~25
    [chunk, T] fp32 intermediates alive), then the bridge projections and ONE VICReg loss over the step's embeddings.
    -> (sounds/s, ms per step, per-stage ms per step)."""
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import vicreg as OV

    sounds_per_step = sounds_per_step or chunk
    nchunks = max(1, (sounds_per_step + chunk - 1) // chunk)
    times = []
    stage = {}
    for i in range(warmup + steps):
        t = {}
        t0 = time.perf_counter()
        xs, ys = [], []
        for c in range(nchunks):
            tc = {}
            n = min(chunk, sounds_per_step - c * chunk)
            # chunk c of step i = batch number i * nchunks + c of a batch-size-`chunk` Voice: the same sound ids as one
            # batch of `sounds_per_step` sounds (ids are batch_idx * B + row) when chunk divides it
            r = harness.oracle_front_end(i * nchunks + c, n, N=bands, seconds=seconds, timings=tc, torch_ops=True,
                                         cfg_batch=sounds_per_step, reproducible=reproducible)
            xs.append(r["x"])
            ys.append(r["y"])
            for k, v in tc.items():
                t[k] = t.get(k, 0.0) + v
        if nchunks > 1:  # the loss of the step sees all of its sounds (the per-chunk value above is discarded)
            tl = time.perf_counter()
            OV.loss_torch(torch.cat(xs), torch.cat(ys), sounds_per_step, harness.EMBED_DIM)
            t["vicreg"] = t.get("vicreg", 0.0) + time.perf_counter() - tl
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            for k, v in t.items():
                stage[k] = stage.get(k, 0.0) + v
    total = sum(times)
    return (sounds_per_step * len(times) / total, total / len(times) * 1e3,
            {k: v / len(times) * 1e3 for k, v in stage.items()})


def workload_name(bands: int, seconds: float, per_gpu: int, world: int, exchange: str) -> str:
    return ("BASELINE configs[%d] shard: front end synth->PQMF(N=%d)->VICReg(D=256), %d sounds x %g s @ 44.1 kHz per GPU, "
            "global batch %d, embedding exchange %s" % (4 if seconds > 4.0 else 3, bands, per_gpu, seconds,
                                                        per_gpu * world, exchange))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = args.batch_per_gpu
    value, ms, stage = cpu_step_rate(args.cpu_sample, args.bands, args.seconds, args.steps, args.warmup,
                                     sounds_per_step=per_step, reproducible=args.reproducible)
    cores = torch.get_num_threads()
    sample = (f"{per_step} sounds x {args.seconds:g} s per step in chunks of {args.cpu_sample} (BASELINE configs[0] is one "
              f"chunk), oracle/ CPU path: torch fp32 Voice restatement -> torch conv1d PQMF -> bridge -> torch VICReg ops "
              f"on the step's {per_step} embeddings, {cores} threads")
    line = {
        "impl": "reference", "metric": "4s@44.1kHz sounds/sec (synth->PQMF->VICReg)", "value": value,
        "unit": "sounds/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.bands, args.seconds, per_step, 1, "n/a (CPU reference arm: one "
                                             "process, one shard; rates are per sound)"),
                   "per_gpu_batch": per_step, "seconds": args.seconds, "bands": args.bands,
                   "cpu_chunk": args.cpu_sample},
        "cpu_baseline": {"value": value, "unit": "sounds/s", "cores": cores, "kind": "port", "sample": sample,
                         "stage_ms": stage},
        "e2e": {"value": value, "unit": "sounds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def source_hash() -> str:
    """sha256 over the kernel sources of the build being benchmarked (ties committed ncu captures to a build)."""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(harness.PKG, "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".cpp")):
            h.update(name.encode())
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()[:16]


def parity_block(args, rank, world, dev, voice, gram, vic, wa, wp):
    """Results against the CPU oracle, OUTSIDE every timed region (SURVEY 8c tolerances).  Every N:
    the four loss terms each rank computed through the route that was timed (statistics exchange / peer gather / NCCL)
    against the float64 oracle on the rank-ordered concatenation of all ranks' embeddings (gathered here with a plain
    NCCL all-gather) -- the N-rank == 1-rank semantics of FullGatherLayer (vicreg.py:38-39,79-95).  Rank 0 additionally
    renders a small batch on the CPU oracle and checks audio, PQMF bands and loss terms of the GPU path against it.
    -> (dict for the JSON line, list of violated bounds)."""
    import numpy as np
    import torch.distributed as dist

    import ias_b200
    from oracle import vicreg as OV

    B = voice.batch_size
    failures = []
    out = {"where": "after the timed regions, not included in any reported time"}
    # ---- exchange route: N ranks == oracle on the concatenation ----
    batch_no = 900_001
    audio, params, _ = voice(batch_no * world + rank)
    _, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice.row_scale)
    with torch.no_grad():
        l4 = torch.stack(vic.loss(x, y)).double()
    xy = torch.cat([x, y], dim=1).contiguous()
    if world > 1:
        xy_all = torch.empty((world,) + tuple(xy.shape), dtype=xy.dtype, device=dev)
        dist.all_gather_into_tensor(xy_all.view(-1), xy.view(-1))
        l4_all = torch.empty((world, 4), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(l4_all.view(-1), l4.contiguous())
    else:
        xy_all, l4_all = xy[None], l4[None]
    if rank == 0:
        xa = xy_all[:, :, :D].reshape(world * B, D).double().cpu().numpy()
        ya = xy_all[:, :, D:].reshape(world * B, D).double().cpu().numpy()
        got = l4_all.cpu().numpy()
        rel = np.zeros(4)
        for r in range(world):
            want = np.array(OV.loss(xa, ya, world * B, D, local_rows=slice(r * B, (r + 1) * B)))
            rel = np.maximum(rel, np.abs(got[r] - want) / np.abs(want))
        same = bool(np.all(got[:, 2:] == got[0, 2:]))
        out["exchange"] = {
            "route": (args.gather if world > 1 else "single process"), "ranks": world, "global_batch": world * B,
            "loss4_rel": [float(v) for v in rel], "std_cov_identical_across_ranks": same,
            "oracle": "oracle/vicreg.py float64 on the concatenation of every rank's embeddings (plain NCCL all-gather)",
            "bound": 1e-4}
        if not (np.all(rel <= 1e-4) and same):
            failures.append("exchange loss4_rel %s identical=%s" % (rel, same))
        # ---- rank 0: a small batch against the CPU oracle ----
        P = args.parity_sounds or (128 if args.seconds <= 4.0 else 16)
        # reproducible mode renders multiples of 32 (the first P rows are compared); same noise mode as the timed run
        Bv = (P + 31) // 32 * 32 if args.reproducible else P
        ref = harness.oracle_front_end(0, P, N=args.bands, seconds=args.seconds, cfg_batch=P,
                                       reproducible=args.reproducible)
        cfgP = ias_b200.SynthConfig(batch_size=Bv, reproducible=args.reproducible, sample_rate=44100,
                                    buffer_size_seconds=args.seconds)
        vP = ias_b200.Voice(synthconfig=cfgP, normalize=voice.normalize).to(dev)
        vcfg = types.SimpleNamespace(dim=D, embeddim=D, vicreg=types.SimpleNamespace(
            mlp="8-8-%d", batch_size=P, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
        vicP = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity(), gather=False)
        a, prm, _ = vP(0)
        a, prm = a[:P].contiguous(), prm[:P].contiguous()
        scaleP = vP.row_scale[:P].contiguous() if vP.row_scale is not None else None
        # deferred normalisation: the audio tensor is the raw mix; what is compared is raw * factor
        a_cmp = a if scaleP is None else a * scaleP[:, None]
        err = (a_cmp.cpu() - ref["audio"]).abs().max(dim=1)[0]
        bands_ref_in = gram(ref["audio"].unsqueeze(1).to(dev)).cpu()
        pq = float((bands_ref_in - ref["bands"]).abs().max() / ref["bands"].abs().max())
        want = np.array(ref["loss4"])
        with torch.no_grad():
            g_or = np.array([float(v) for v in vicP.loss(ref["x"].to(dev), ref["y"].to(dev))])
            _, xg, yg = harness.analysis_bridge(gram, a, prm, wa, wp, scaleP)
            g_e2e = np.array([float(v) for v in vicP.loss(xg, yg)])
        rel_or = np.abs(g_or - want) / np.abs(want)
        rel_e2e = np.abs(g_e2e - want) / np.abs(want)
        out["oracle_batch"] = {
            "sounds": P, "seconds": args.seconds, "params_bit_equal": bool(torch.equal(prm.cpu(), ref["params"])),
            "voices_le_1e-4": int((err <= 1e-4).sum()), "audio_median": float(err.median()),
            "audio_max": float(err.max()), "pqmf_rel": pq, "loss4_rel_on_oracle_embeddings": [float(v) for v in rel_or],
            "loss4_rel_end_to_end": [float(v) for v in rel_e2e],
            "bounds": {"audio": 1e-4, "pqmf_rel": 1e-5, "loss4_rel": 1e-4},
            "note": "oracle = oracle/ on this box's CPU (torch fp32 Voice restatement, parity unpinned w.r.t. torchsynth; "
                    "numpy float64 PQMF and loss).  Voices above 1e-4 trace to 1-ulp LFO cosine differences amplified by "
                    "the pitch path (DESIGN.md 4); the end-to-end loss terms inherit them."}
        if not out["oracle_batch"]["params_bit_equal"]:
            failures.append("seeded parameters differ from torch's CPU generator")
        # 4 s clips: the fp64 phase scan is exact and the median voice agrees to ~3e-7.  Longer clips: partial sums
        # exceed 2^53 * 2^-33, a few phase arguments per million differ by one ulp and the median rises
        # (measured 1.1e-5 at 30 s): bounded by the north-star 1e-4 there.
        median_bound = 1e-5 if args.seconds <= 4.0 else 1e-4
        out["oracle_batch"]["bounds"]["audio_median"] = median_bound
        if float(err.median()) > median_bound:
            failures.append("audio median error %g" % float(err.median()))
        if pq > 1e-5:
            failures.append("pqmf_rel %g" % pq)
        if not np.all(rel_or <= 1e-4):
            failures.append("loss4 on oracle embeddings %s" % rel_or)
    return out, failures


def run_ours(args):
    import torch.distributed as dist

    import ias_b200
    from ias_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = ias_b200.lib()
    _lib.check(lib.ias_device_check(local), "ias_device_check")

    B = args.batch_per_gpu
    T = int(args.seconds * 44100)
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=args.reproducible, sample_rate=44100,
                               buffer_size_seconds=args.seconds)
    voice = ias_b200.Voice(synthconfig=cfg, normalize="defer" if args.normalize == "defer" else True).to(dev)
    gram = ias_b200.PQMF(N=args.bands).to(dev)
    vcfg = types.SimpleNamespace(dim=D, embeddim=D, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B * world, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(dev)
    exchange_desc = "n/a at N=1"
    if world > 1:
        if args.gather == "stats":
            ias_b200.use_fused_gather(ias_b200.StatsExchange(D, dev))
            exchange_desc = ("statistics exchange: per-rank mean / second moments / tcgen05 Gram pushed to every peer over "
                             "NVLink by the loss kernels (0.39 MB per rank and step), pooled on arrival; no NCCL on the "
                             "data path (NCCL carries only process-group setup and the timing all-reduce)")
        elif args.gather == "peer":
            ias_b200.use_fused_gather(ias_b200.EmbeddingExchange(B, D, dev))
            exchange_desc = ("fused into the loss kernels: the column-statistics kernel reads every peer's embeddings over "
                             "NVLink peer memory; no NCCL on the data path")
        else:
            exchange_desc = "NCCL all-gather of the embeddings (FullGatherLayer through torch.distributed)"

    def step(i: int):
        audio, params, _ = voice(i * world + rank)       # sound ids [ (i*W + r) * B, ... ): SURVEY 8(d) config 4
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice.row_scale)
        with torch.no_grad():
            return vic.loss(x, y)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total_steps = args.warmup + args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    idx_dev = torch.zeros(1, dtype=torch.int64, device=dev)  # this rank's batch number, resident on the device

    def step_index(idx):
        audio, params, _ = voice(idx)                    # a device-resident index is read by the seeding kernel
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice.row_scale)
        with torch.no_grad():
            return vic.loss(x, y)

    def step_dev():
        return torch.stack(step_index(idx_dev))

    def step_dev_advance():
        out = step_dev()
        idx_dev.add_(world)                              # next step's batch number, computed on the device
        return out

    pipe_side = torch.cuda.Stream(device=dev)  # (a high-priority stream measured slower: 1.243 vs 1.222 ms/step)
    pipelined = args.pipeline

    def step_pipelined(next_from_host=None, result_to_host=None):
        """One step of the software-pipelined front end: renders the batch whose parameters and control stage the
        previous step prepared (audio stage only), then -- on a side stream, while PQMF / bridge / loss of this batch
        run -- takes the next batch number (advanced on the device, or copied from pinned host memory) and prepares
        it: seed -> ADSR -> control -> schedule.  Same kernels, same results; the latency-bound control stage hides
        behind the bandwidth-bound tail of the step."""
        audio, params, _ = voice(idx_dev, prepared=True)
        cur = torch.cuda.current_stream()
        pipe_side.wait_stream(cur)

        def next_index():
            if next_from_host is not None:
                idx_dev.copy_(next_from_host, non_blocking=True)   # H2D: the loader runs ahead of the render
            else:
                idx_dev.add_(world)

        if args.pipeline_depth == 1:
            with torch.cuda.stream(pipe_side):
                next_index()
                voice.prepare(idx_dev)
            bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice.row_scale)
        else:
            with torch.cuda.stream(pipe_side):
                voice.prepare_modulation()                         # batch k+1, under the PQMF analysis of batch k
            analysis_done = torch.cuda.Event()
            bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice.row_scale,
                                                  after_analysis=lambda: analysis_done.record(cur))
            with torch.cuda.stream(pipe_side):
                pipe_side.wait_event(analysis_done)                # batch k+2: seed + envelopes, under bridge / loss
                next_index()
                voice.prepare_envelopes(idx_dev)
        with torch.no_grad():
            out = torch.stack(vic.loss(x, y))
        if result_to_host is not None:
            result_to_host.copy_(out, non_blocking=True)           # D2H of the step's result
        cur.wait_stream(pipe_side)
        return out

    def prime(first_batch: int, second_batch=None):
        """Fill the pipeline: batch ``first_batch`` ready to render and, two deep, the envelopes of the batch after."""
        idx_dev.fill_(first_batch)
        if args.pipeline_depth == 1:
            voice.prepare(idx_dev)
            return
        voice.prepare_envelopes(idx_dev)
        voice.prepare_modulation()
        idx_dev.fill_(first_batch + world if second_batch is None else second_batch)
        voice.prepare_envelopes(idx_dev)

    def capture(fn):
        """fn() captured in a CUDA graph (after a side-stream warm-up, as torch requires); None if capture fails."""
        if args.no_graph:
            return None, None, "eager launches (--no-graph)"
        try:
            keep = idx_dev.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            sync()
            g = torch.cuda.CUDAGraph()
            if args.main_priority:
                cap_stream = torch.cuda.Stream(device=dev, priority=-1)
                cap_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.graph(g, stream=cap_stream):
                    out = fn()
            else:
                with torch.cuda.graph(g):
                    out = fn()
            idx_dev.copy_(keep)
            return g, out, "one CUDA graph replay per step"
        except Exception as exc:  # e.g. a collective that cannot be captured
            torch.cuda.synchronize()
            return None, None, "eager launches (graph capture failed: %s)" % type(exc).__name__

    # ---- device-resident timing: `value` ----
    # W warm-up steps, then exactly K steps between two CUDA events on the launch stream.  Each step is one replay of
    # the captured step (seed -> control -> schedule -> audio -> PQMF -> bridge -> loss, batch number advanced on the
    # device), so launch gaps between the ~14 kernels do not count against the GPU.
    for i in range(args.warmup):
        step(i)
    sync()
    step_value = step_pipelined if pipelined else step_dev_advance
    if pipelined:
        prime(args.warmup * world + rank)
    graph_v, out_v, value_mode = capture(step_value)
    if pipelined:
        value_mode += ("; control stage of batch k+1 (Voice.prepare) overlapped with PQMF/bridge/loss of batch k"
                       if args.pipeline_depth == 1 else
                       "; modulation stage of batch k+1 under the PQMF analysis of batch k, seeding + envelopes of batch "
                       "k+2 under its bridge / loss kernels (Voice.prepare_modulation / prepare_envelopes)")
        prime(args.warmup * world + rank)
    else:
        idx_dev.fill_(args.warmup * world + rank)
    for _ in range(3):
        if graph_v is not None:
            graph_v.replay()
        else:
            out_v = step_value()
    if pipelined:
        prime(args.warmup * world + rank)
    else:
        idx_dev.fill_(args.warmup * world + rank)
    sync()
    if rank == 0:
        sampler.wait_first_sample()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    sampler.mark_begin()
    e0.record()
    for i in range(args.steps):
        if graph_v is not None:
            graph_v.replay()
        else:
            out_v = step_value()
    e1.record()
    sync()
    sampler.mark_end()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    loss_vals = [float(o) for o in out_v]

    # ---- per-kernel times: a separate eager pass with the library's event brackets on (not part of `value`) ----
    lib.ias_prof_reset()
    lib.ias_prof_enable(1)
    prof_steps = min(args.steps, 20)
    for i in range(args.warmup, args.warmup + prof_steps):
        step(i)
    sync()
    launches_per_step = int(lib.ias_prof_launches(-1)) // prof_steps
    launches = launches_per_step * args.steps
    lib.ias_prof_enable(0)
    kern = {}
    import ctypes

    for k in range(lib.ias_prof_kernel_count()):
        tot, n = ctypes.c_double(0), ctypes.c_longlong(0)
        _lib.check(lib.ias_prof_read(k, ctypes.byref(tot), ctypes.byref(n)), "ias_prof_read")
        if n.value:
            kern[lib.ias_prof_kernel_name(k).decode()] = {"ms_per_launch": tot.value / n.value, "launches": n.value}

    # ---- end to end through the public API with host buffers: `e2e` ----
    # Per step: the batch number arrives in pinned host memory and is copied to the device (what Lightning does with
    # the reference's integer DataLoader, runsetup.py:46-48), the step runs on it, and the four loss scalars are read
    # back to the host (which synchronises every step).  Voice(batch_idx) accepts the device-resident index
    # (ias_voice_seed_params_dev), so nothing in the step needs the host and it replays from a CUDA graph.
    # The two copies are part of the replayed graph (memcpy nodes on the pinned buffers), so a step costs the host one
    # store, one graph launch and one stream synchronisation.
    batch_numbers = (torch.arange(args.warmup, total_steps, dtype=torch.int64) * world + rank).tolist()
    host_in = torch.zeros(1, dtype=torch.int64).pin_memory()
    host_out = torch.empty(4, dtype=torch.float32).pin_memory()

    def step_e2e_plain():
        idx_dev.copy_(host_in, non_blocking=True)      # H2D of the step's input
        out = step_dev()
        host_out.copy_(out, non_blocking=True)         # D2H of the step's result
        return out

    def step_e2e_pipelined():
        return step_pipelined(next_from_host=host_in, result_to_host=host_out)

    step_e2e = step_e2e_pipelined if pipelined else step_e2e_plain

    def prime_e2e():
        prime(batch_numbers[0], batch_numbers[min(1, len(batch_numbers) - 1)])

    if pipelined:
        host_in[0] = batch_numbers[0]
        prime_e2e()
    graph, static_out, e2e_mode = capture(step_e2e)
    if pipelined:
        e2e_mode += ("; the loader runs %d batch(es) ahead: step k copies batch number k+%d in and prepares it on a side "
                     "stream" % (args.pipeline_depth, args.pipeline_depth))
    stream = torch.cuda.current_stream()

    def run_e2e_step(j: int):
        # the DataLoader's integer arrives in pinned host memory: this step's batch (plain) or the next one's (pipelined)
        ahead = args.pipeline_depth if pipelined else 0
        host_in[0] = batch_numbers[min(j + ahead, len(batch_numbers) - 1)]
        if graph is not None:
            graph.replay()
        else:
            step_e2e()
        stream.synchronize()                           # the result is on the host: the step is over

    if pipelined:
        prime_e2e()
    for j in range(min(3, args.steps)):  # warm the replay path
        run_e2e_step(j)
    if pipelined:
        prime_e2e()
    sync()
    e0.record()
    for j in range(args.steps):
        run_e2e_step(j)
    e1.record()
    sync()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_last = [float(v) for v in host_out]

    # ---- variant: parameters supplied by the host (pinned [78,B] block copied in every step, seeding skipped) ----
    host_params = torch.rand((78, B), generator=torch.Generator().manual_seed(4321)).pin_memory()

    def step_host_params():
        voice._store.copy_(host_params, non_blocking=True)
        audio = voice.output()
        bands, x, y = harness.analysis_bridge(gram, audio, voice.params01(), wa, wp, voice.row_scale)
        with torch.no_grad():
            o = vic.loss(x, y)
        host_out.copy_(torch.stack(o))

    for j in range(3):  # warm-up: the eager path allocates its own audio / band buffers (the graphs own theirs)
        step_host_params()
    sync()
    e0.record()
    for j in range(args.steps):
        step_host_params()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_params_ms = float(t.item())

    # ---- variant: the other noise mode (default run: reproducible=True, the 32-row L2-resident table) ----
    e2e_nr_ms, nr_mode = None, None
    if not args.no_noise_variant:
        cfg_nr = ias_b200.SynthConfig(batch_size=B, reproducible=not args.reproducible, sample_rate=44100,
                                      buffer_size_seconds=args.seconds)
        voice_nr = ias_b200.Voice(synthconfig=cfg_nr, normalize=voice.normalize).to(dev)

        def step_e2e_nr():
            idx_dev.copy_(host_in, non_blocking=True)
            audio, params, _ = voice_nr(idx_dev)
            bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice_nr.row_scale)
            with torch.no_grad():
                out = torch.stack(vic.loss(x, y))
            host_out.copy_(out, non_blocking=True)
            return out

        graph_nr, _, nr_mode = capture(step_e2e_nr)

        def run_nr(j):
            host_in[0] = batch_numbers[j]
            if graph_nr is not None:
                graph_nr.replay()
            else:
                step_e2e_nr()
            stream.synchronize()

        for j in range(min(3, args.steps)):
            run_nr(j)
        sync()
        e0.record()
        for j in range(args.steps):
            run_nr(j)
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_nr_ms = float(t.item())
        del voice_nr, graph_nr

    clocks = sampler.stop() if rank == 0 else None
    parity, parity_failures = (None, [])
    if not args.no_parity:
        parity, parity_failures = parity_block(args, rank, world, dev, voice, gram, vic, wa, wp)
    def teardown():
        # graphs that captured collectives (--gather nccl) must be gone before the communicator is, and every rank
        # leaves together: rank 0 is still busy with the CPU oracle of the parity block while the others are done
        nonlocal graph_v, graph
        graph_v = graph = None
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return

    sounds = B * world * args.steps
    value = sounds / (ms_total * 1e-3)
    peaks = {}
    peak_src = "fallback"
    try:
        peaks = json.load(open(os.path.join(harness.ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    scale_T = T / T_4S
    va = kern.get("k_voice_audio", {"ms_per_launch": float("nan"), "launches": 0})
    # dominant kernel: k_voice_audio.  Algorithmic bytes per launch: 4*T audio written per sound, plus
    # 4*T noise read per sound when the noise table is the reference's [B,T] one (reproducible=False: 722 MB, streamed
    # from HBM); with the 32-row reproducible table the noise is L2 resident and not counted.
    noise_bytes = 0.0 if args.reproducible else 4.0 * T * B
    voice_bytes = 4.0 * T * B + noise_bytes
    achieved = voice_bytes / (va["ms_per_launch"] * 1e-3) / 1e9
    # roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
    # (profiles/capture_k_voice_audio.json, written by tools/summarize_ncu.py) -- used only if the capture was taken
    # from the very sources being benchmarked and on this workload shape; null otherwise.
    # roofline.issue: the kernel is bound on the instruction side, not by HBM (SURVEY 8d "report min(HBM, issue)"):
    # cycles its tile loop needs on the busiest execution pipe (tools/issue_model.py: SASS of this build priced with the
    # pipe occupancies measured on B200) against what SMs x 4 schedulers x clock offer during the measured launch time.
    traffic, issue, capture_note = None, None, "no capture of this build"
    src = source_hash()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    clk = float((clocks or {}).get("sm_mhz") or 1965.0)
    try:
        mdl = json.load(open(os.path.join(harness.ROOT, "profiles", "issue_model.json")))
        if mdl.get("src_sha256") == src and T % 16 == 0:
            need = mdl["bound_cycles_per_sample"] * float(B) * T / 32.0
            avail = sms * 4 * clk * 1e6 * (va["ms_per_launch"] * 1e-3)
            issue = {"instr_per_sample": mdl["instr_per_sample"], "pipe_cycles_per_sample": mdl["pipe_cycles_per_sample"],
                     "bound_pipe": mdl["bound_pipe"], "bound_cycles_per_sample": mdl["bound_cycles_per_sample"],
                     "achieved_frac": need / avail, "cycles_per_sample_measured": avail / (float(B) * T / 32.0),
                     "sm_mhz_used": clk, "schedulers": sms * 4,
                     "note": "achieved_frac = cycles the busiest issue-side resource (bound_pipe: the dispatch port, one "
                             "warp instruction per cycle; the FMA pipe needs 95 %% of that) needs for B*T samples / (SMs x 4 "
                             "schedulers x clock x measured launch time); an upper estimate: the ~11 %% of samples in silent "
                             "tails are zero-filled, not rendered.  1.0 would be a kernel issuing one instruction per cycle "
                             "and scheduler.  Occupancy scaling (1..4 CTAs per SM: 5.10 / 3.25 / 2.81 / 2.55 ms per 3552 "
                             "full voices) and the variants with more warps or no serial section between tiles (all slower) "
                             "are in DESIGN.md 3.1 / 3.2.  Per warp and sample; model: tools/issue_model.py" % ()}
    except Exception:
        pass
    try:
        cap = json.load(open(os.path.join(harness.ROOT, "profiles", "capture_k_voice_audio.json")))
        if cap.get("src_sha256") != src:
            capture_note = "profiles/capture_k_voice_audio.json is of sources %s, this build is %s" % (
                cap.get("src_sha256"), src)
        elif not (B == cap.get("B") and T == cap.get("T")):
            capture_note = "capture is of B=%s T=%s" % (cap.get("B"), cap.get("T"))
        else:
            traffic = cap.get("dram_bytes_per_launch")
            capture_note = "ncu --set full capture %s of this build" % cap.get("run")
            if issue is not None:
                issue["ncu"] = {k: cap.get(k) for k in ("inst_executed_per_launch", "issue_active_pct", "pipes_pct",
                                                        "gpu_time_us")}
                if cap.get("inst_executed_per_launch"):
                    issue["ncu"]["instr_per_sample_executed"] = cap["inst_executed_per_launch"] * 32.0 / (float(B) * T)
    except Exception:
        pass
    line = {
        "metric": "4s@44.1kHz sounds/sec (synth->PQMF->VICReg)",
        "value": value,
        "unit": "sounds/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": workload_name(args.bands, args.seconds, B, world, exchange_desc),
            "per_gpu_batch": B, "global_batch": B * world, "seconds": args.seconds, "bands": args.bands,
            "exchange": args.gather if world > 1 else None,
            "noise": ("reproducible=True: 32-row table, L2 resident" if args.reproducible else
                      "reproducible=False, the reference's setting (conf/config.yaml:38): [B,T] noise table, %.0f MB "
                      "streamed from HBM every step" % (4e-6 * B * T)) + "; the other mode is timed beside it as "
                                                                            "e2e_noise_variant",
            "normalize": ("deferred: the render leaves the raw mix + a [B] factor (1/peak of a clipping row, else 1), the "
                          "PQMF analysis applies it to the bands it writes (Voice(normalize='defer'), "
                          "PQMF.analysis(row_scale=)); bands / loss as with the in-render pass to rounding, checked by the "
                          "parity block" if args.normalize == "defer" else
                          "in the render (Voice(normalize=True)): second pass over the clipping rows"),
            "l2": "inputs larger than L2: %.0f MB audio + %.0f MB bands per step vs 126 MB L2" % (
                4e-6 * B * T, 4e-6 * B * T),
            "bridge": "abs-mean pool to 256 bins + 2 torch matmuls (harness, not a reference component), inside the "
                      "step; device path taken: %s" % harness.LAST_BRIDGE_PATH,
            "rates": "sounds/s: a rate per sound, comparable across batch sizes (the CPU reference arm renders the same "
                     "per-GPU batch in chunks of %d sounds)" % args.cpu_sample,
        },
        "roofline": {
            "bound": "hbm", "kernel": "k_voice_audio_sp (profiled as k_voice_audio)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": capture_note, "issue": issue,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": voice_bytes,
            "algorithmic_bytes_note": "4*T*B audio written" + ("" if args.reproducible else
                                                               " + 4*T*B noise read ([B,T] table, larger than L2)"),
            "frac_audio_write_only": 4.0 * T * B / (va["ms_per_launch"] * 1e-3) / 1e9 / hbm_peak,
            "ms_per_launch": va["ms_per_launch"],
            "note": "k_voice_audio_sp is bound on the instruction side (see `issue`), not by HBM (DESIGN.md 3.1); "
                    "step-level fraction below",
            "step_frac": value / world * ALGO_BYTES_PER_SOUND * scale_T / 1e9 / hbm_peak,
            "step_algorithmic_bytes_per_sound": ALGO_BYTES_PER_SOUND * scale_T,
            "step_frac_with_noise_read": (None if args.reproducible else
                                          value / world * (ALGO_BYTES_PER_SOUND * scale_T + 4.0 * T) / 1e9 / hbm_peak),
        },
        "kernels": kern,
        "e2e": {"value": sounds / (e2e_ms * 1e-3), "unit": "sounds/s", "h2d_bytes_per_step": 8,
                "d2h_bytes_per_step": 16,
                "mode": e2e_mode + " (H2D and D2H copies are nodes of the graph; stream synchronised every step)", "loss4_last_step": e2e_last,
                "note": "public API Voice(batch_idx)->PQMF->VICReg.loss; step input is the batch number (pinned host -> "
                        "device, read there by the seeding kernel), parameters are seeded on the device; result = 4 loss "
                        "scalars read back every step"},
        "e2e_host_params": {"value": sounds / (e2e_params_ms * 1e-3), "unit": "sounds/s",
                            "h2d_bytes_per_step": 78 * B * 4, "d2h_bytes_per_step": 16,
                            "note": "same, but the [78,B] parameter block comes from pinned host memory every step"},
        "e2e_noise_variant": (None if e2e_nr_ms is None else {
            "value": sounds / (e2e_nr_ms * 1e-3), "unit": "sounds/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 16,
            "mode": nr_mode, "reproducible": not args.reproducible,
            "note": "same step with SynthConfig(reproducible=%s): %s" % (
                not args.reproducible, "32-row noise table, L2 resident" if not args.reproducible else
                "[B,T] noise table streamed from HBM")}),
        "parity": parity,
        "parity_ok": not parity_failures,
        "src_sha256": src,
        "gpu_launches": launches,
        "gpu_launches_note": "%d kernel launches of libias_b200.so per step (counted by the library in the eager "
                             "profiling pass) x %d timed steps; the timed region issues them as %s" % (
                                 launches_per_step, args.steps, value_mode),
        "value_mode": value_mode,
        "clocks": clocks,
        "loss4_last_step": loss_vals,
    }
    if not args.no_cpu_baseline and world == 1:
        # bounded sample: 30 steps x 128 sounds, about 10 s of host work on the box's 16 cores
        # bounded sample: 30 steps x 128 four-second sounds (16 thirty-second ones), about 10 s of host work
        cpu_steps = 30
        chunk = args.cpu_sample if args.seconds <= 4.0 else max(8, int(args.cpu_sample * 4.0 / args.seconds) // 8 * 8)
        v, ms, stage = cpu_step_rate(chunk, args.bands, args.seconds, steps=cpu_steps, warmup=2,
                                     reproducible=args.reproducible)
        line["cpu_baseline"] = {
            "value": v, "unit": "sounds/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{chunk} sounds x {args.seconds:g} s per step x {cpu_steps} steps (BASELINE configs[0]) through "
                      "oracle/ (torch fp32 Voice restatement -> torch conv1d PQMF -> bridge -> torch VICReg ops)",
            "ms_per_step": ms, "stage_ms": stage}
    print(json.dumps(line), flush=True)
    teardown()
    if parity_failures:
        print("bench.py: PARITY FAILED: " + "; ".join(parity_failures), file=sys.stderr)
        sys.exit(3)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
