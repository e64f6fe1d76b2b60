#!/usr/bin/env python
"""Turn one GPU session's scratch output (gpurun_out/*_<tag>.*) into the tracked summaries under profiles/.

    python tools/summarize_ncu.py r01e

Writes profiles/<tag>_ncu_full_summary.md (selected `ncu --set full` metrics of every prof_*_<tag>.ncu-rep),
profiles/<tag>_ncu_launch_list.md (per-kernel totals of the launch-list pass and each kernel's share of the step),
copies the bench lines / test logs, and refreshes profiles/traffic.json (DRAM bytes per launch, read by bench.py).
"""
import collections
import csv
import glob
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return None
    head, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(head, units, vals)}


def write_capture_json(tag, m, rep_name):
    """profiles/capture_k_voice_audio.json: what bench.py's roofline.traffic / roofline.issue.ncu quote, tied to the
    source hash of the session's own bench line (a capture of other sources is ignored by bench.py)."""
    def val(k, default=None):
        try:
            return float(m[k][1].replace(",", ""))
        except Exception:
            return default

    src, B, T = None, 1024, 176400
    for cand in (f"bench_{tag}.json", f"plain_{tag}.log"):
        try:
            for line in open(os.path.join(OUT, cand)):
                if line.startswith("{"):
                    d = json.loads(line)
                    src = d.get("src_sha256")
                    B = d["config"]["per_gpu_batch"]
                    T = int(d["config"]["seconds"] * 44100)
        except Exception:
            pass
        if src:
            break
    rd = val("dram__bytes_read.sum", 0.0) * UNIT_BYTES.get(m.get("dram__bytes_read.sum", ("byte",))[0], 1)
    wr = val("dram__bytes_write.sum", 0.0) * UNIT_BYTES.get(m.get("dram__bytes_write.sum", ("byte",))[0], 1)
    out = {
        "run": tag, "capture": rep_name, "src_sha256": src, "B": B, "T": T,
        "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
        "gpu_time_us": val("gpu__time_duration.sum"),
        "inst_executed_per_launch": val("smsp__inst_executed.sum"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipes_pct": {p: val(f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active") for p in ("fma", "alu", "xu", "lsu")},
        "fp64_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "registers": val("launch__registers_per_thread"),
    }
    json.dump(out, open(os.path.join(PROF, "capture_k_voice_audio.json"), "w"), indent=1)


def full_summaries(tag):
    lines = [f"# ncu --set full summaries, run {tag} (B200, bench.py --steps 3 --warmup 3 --no-cpu-baseline, "
             "one launch per kernel)\n"]
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(OUT, f"prof_*_{tag}.ncu-rep"))):
        m = raw_page(rep)
        if m is None:
            lines.append(f"\n## {os.path.basename(rep)}: empty capture\n")
            continue
        name = m.get("Kernel Name", ("", "?"))[1]
        lines.append(f"\n## {name[:110]}  ({os.path.basename(rep)})\n")
        for k in METRICS:
            if k in m:
                lines.append(f"{k:<80} {m[k][0]:<16} {m[k][1]}")
        stalls = []
        for k, (u, v) in m.items():
            if k.startswith(STALL_PREFIX) and k.endswith(STALL_SUFFIX):
                try:
                    stalls.append((float(v), k[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append("top stall reasons (warps per issue-active cycle): " +
                     ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))
        if "k_voice_audio" in name:
            write_capture_json(tag, m, os.path.basename(rep))
        try:
            rd = float(m["dram__bytes_read.sum"][1]) * UNIT_BYTES[m["dram__bytes_read.sum"][0]]
            wr = float(m["dram__bytes_write.sum"][1]) * UNIT_BYTES[m["dram__bytes_write.sum"][0]]
            hit = re.search(r"\b(k_[a-z0-9_]+)", name)
            short = hit.group(1) if hit else name.split("(")[0].strip()
            traffic[short] = {"per_launch_bytes": int(rd + wr), "read": int(rd), "write": int(wr),
                              "capture": os.path.basename(rep)}
        except Exception:
            pass
    if len(lines) > 1:  # a session without full captures (tools/gpu_final.sh) leaves no empty summary behind
        with open(os.path.join(PROF, f"{tag}_ncu_full_summary.md"), "w") as f:
            f.write("\n".join(lines) + "\n")
    if traffic:
        path = os.path.join(PROF, "traffic.json")
        old = {}
        try:
            old = json.load(open(path))
        except Exception:
            pass
        old.update(traffic)
        old["_run"] = tag
        json.dump(old, open(path, "w"), indent=1)
    return traffic


def launch_list(tag):
    src = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    if not rows:
        return
    head = rows[0]
    try:
        i_name, i_val, i_unit = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    except ValueError:
        return
    tot = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[i_unit], 1.0)
        name = r[i_name].split("(")[0][-70:]
        t = tot.setdefault(name, [0.0, 0])
        t[0] += v * scale
        t[1] += 1
    total = sum(v[0] for v in tot.values())
    out = [f"# ncu launch list, run {tag}: gpu__time_duration.sum per kernel (cold-cache, serialised), share of all launches\n",
           "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f} % |")
    with open(os.path.join(PROF, f"{tag}_ncu_launch_list.md"), "w") as f:
        f.write("\n".join(out) + "\n")


def copy_logs(tag):
    for pat in (f"bench_{tag}.json", f"bench_ref_{tag}.json", f"smoke_{tag}.log", f"test_*_{tag}.log",
                f"bench_g*_{tag}.json", f"multi_check_{tag}.log", f"topo_{tag}.txt", f"configs_{tag}.json",
                f"configs_{tag}.jsonl"):
        for p in glob.glob(os.path.join(OUT, pat)):
            if os.path.getsize(p) < 200_000:
                shutil.copy(p, os.path.join(PROF, os.path.basename(p)))


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    t = full_summaries(tag)
    launch_list(tag)
    copy_logs(tag)
    print("summarised", tag, {k: v["per_launch_bytes"] for k, v in t.items()})
