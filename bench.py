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
    ap.add_argument("--pipeline", action="store_true",
                    help="overlap the next batch's control stage (Voice.prepare, side stream) with this batch's PQMF / "
                         "loss; measured +2.3 %% (1.250 -> 1.222 ms/step), off by default so a step is self-contained")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused = loss kernels read peers' embeddings over NVLink (symmetric memory); "
                         "nccl = FullGatherLayer all-gather through torch.distributed")
    ap.add_argument("--cpu-sample", type=int, default=128, help="sounds per CPU-baseline step (config 1: 128)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []  # (arrival wall-clock, csv line)
        self.proc = None
        self.t_begin = None
        self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        lo = (self.t_begin or 0.0)
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
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle on the host cores)
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_rate(sample: int, bands: int, seconds: float, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    stage = {}
    for i in range(warmup + steps):
        t = {}
        t0 = time.perf_counter()
        harness.oracle_front_end(i, sample, N=bands, seconds=seconds, timings=t, torch_ops=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            for k, v in t.items():
                stage[k] = stage.get(k, 0.0) + v
    total = sum(times)
    return sample * len(times) / total, total / len(times) * 1e3, {k: v / len(times) * 1e3 for k, v in stage.items()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, stage = cpu_step_rate(args.cpu_sample, args.bands, args.seconds, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = (f"{args.cpu_sample} sounds x {args.seconds:g} s per step (BASELINE configs[0]), oracle/ CPU path: torch fp32 "
              f"Voice restatement -> torch conv1d PQMF -> bridge -> torch VICReg ops, {cores} threads")
    line = {
        "impl": "reference", "metric": "4s@44.1kHz sounds/sec (synth->PQMF->VICReg)", "value": value,
        "unit": "sounds/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "front end synth->PQMF(N=%d)->VICReg(D=256), %g s voices, CPU sample of %d sounds/step"
                   % (args.bands, args.seconds, args.cpu_sample)},
        "cpu_baseline": {"value": value, "unit": "sounds/s", "cores": cores, "kind": "port", "sample": sample,
                         "stage_ms": stage},
        "e2e": {"value": value, "unit": "sounds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
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
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=args.seconds)
    voice = ias_b200.Voice(synthconfig=cfg).to(dev)
    gram = ias_b200.PQMF(N=args.bands).to(dev)
    vcfg = types.SimpleNamespace(dim=D, embeddim=D, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B * world, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(dev)
    if world > 1 and args.gather == "fused":
        ias_b200.use_fused_gather(ias_b200.EmbeddingExchange(B, D, dev))

    def step(i: int):
        audio, params, _ = voice(i * world + rank)       # sound ids [ (i*W + r) * B, ... ): SURVEY 8(d) config 4
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp)
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
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp)
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
        with torch.cuda.stream(pipe_side):
            if next_from_host is not None:
                idx_dev.copy_(next_from_host, non_blocking=True)   # H2D: the loader runs one batch ahead
            else:
                idx_dev.add_(world)
            voice.prepare(idx_dev)
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp)
        with torch.no_grad():
            out = torch.stack(vic.loss(x, y))
        if result_to_host is not None:
            result_to_host.copy_(out, non_blocking=True)           # D2H of the step's result
        cur.wait_stream(pipe_side)
        return out

    def prime(first_batch: int):
        idx_dev.fill_(first_batch)
        voice.prepare(idx_dev)

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
        value_mode += "; control stage of batch k+1 (Voice.prepare) overlapped with PQMF/bridge/loss of batch k"
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
    if pipelined:
        host_in[0] = batch_numbers[0]
        prime(batch_numbers[0])
    graph, static_out, e2e_mode = capture(step_e2e)
    if pipelined:
        e2e_mode += "; the loader runs one batch ahead: step k copies batch number k+1 in and prepares it on a side stream"
    stream = torch.cuda.current_stream()

    def run_e2e_step(j: int):
        # the DataLoader's integer arrives in pinned host memory: this step's batch (plain) or the next one's (pipelined)
        host_in[0] = batch_numbers[min(j + 1, len(batch_numbers) - 1)] if pipelined else batch_numbers[j]
        if graph is not None:
            graph.replay()
        else:
            step_e2e()
        stream.synchronize()                           # the result is on the host: the step is over

    if pipelined:
        prime(batch_numbers[0])
    for j in range(min(3, args.steps)):  # warm the replay path
        run_e2e_step(j)
    if pipelined:
        prime(batch_numbers[0])
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
        bands, x, y = harness.analysis_bridge(gram, audio, voice.params01(), wa, wp)
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

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
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
    # dominant kernel: k_voice_audio.  Algorithmic bytes per launch = 4*T written per sound (noise table is the
    # 32-row L2-resident one in reproducible mode) x B sounds.
    voice_bytes = 4.0 * T * B
    achieved = voice_bytes / (va["ms_per_launch"] * 1e-3) / 1e9
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
    try:
        if B == 1024 and T == T_4S:
            traffic = json.load(open(os.path.join(harness.ROOT, "profiles", "traffic.json")))["k_voice_audio"][
                "per_launch_bytes"]
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
            "workload": ("BASELINE configs[3] shard: front end synth->PQMF(N=%d)->VICReg(D=256), %d sounds x %g s @ 44.1 kHz "
                         "per GPU, global batch %d, embedding all-gather %s" % (
                             args.bands, B, args.seconds, B * world, ("fused into the loss kernels over NVLink peer memory" if args.gather == "fused" else "over NCCL")
                             if world > 1 else "n/a at N=1")),
            "per_gpu_batch": B, "global_batch": B * world, "seconds": args.seconds, "bands": args.bands,
            "noise": "reproducible (32-row table)",
            "l2": "inputs larger than L2: 722 MB audio + 722 MB bands per step vs 126 MB L2",
            "bridge": "abs-mean pool to 256 bins (epilogue of k_pqmf_analysis + k_pool_finalize) + 2 torch matmuls (harness, not a reference component), inside the step",
        },
        "roofline": {
            "bound": "hbm", "kernel": "k_voice_audio", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": voice_bytes,
            "ms_per_launch": va["ms_per_launch"],
            "note": "k_voice_audio is instruction-issue bound, not HBM bound (DESIGN.md); step-level fraction below",
            "step_frac": value / world * ALGO_BYTES_PER_SOUND * scale_T / 1e9 / hbm_peak,
            "step_algorithmic_bytes_per_sound": ALGO_BYTES_PER_SOUND * scale_T,
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
        cpu_steps = 30
        v, ms, stage = cpu_step_rate(args.cpu_sample, args.bands, args.seconds, steps=cpu_steps, warmup=2)
        line["cpu_baseline"] = {
            "value": v, "unit": "sounds/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{args.cpu_sample} sounds x {args.seconds:g} s per step x {cpu_steps} steps (BASELINE configs[0]) through "
                      "oracle/ (torch fp32 Voice restatement -> torch conv1d PQMF -> bridge -> torch VICReg ops)",
            "ms_per_step": ms, "stage_ms": stage}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
