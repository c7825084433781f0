"""Summarise one `ncu --set full` capture of the streaming-fold stepper (text to stdout) and,
with --write-counters, refresh profiles/kernel_counters.json -- the per-launch instruction and
DRAM-byte counts bench.py's roofline block uses, keyed by the sha256 of the kernel source.

    python profiles/tools_ncu_summary.py <rep.ncu-rep> <job samples> [--days 564]
           [--stepper-samples N] [--sha <sha256 | file written on the GPU box>] [--write-counters <source tag>]

<job samples> normalises "warp-instr per warp-day" like the round-1 summaries (whole job);
--stepper-samples is the number of samples the captured launch itself stepped (job minus the
pilot's, whose rows are folded from the pilot matrix), used for the per-launch counters.
"""
import argparse
import csv
import hashlib
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = os.path.join(ROOT, "covid19-vaccination-model_b200", "csrc", "vaxmc_kernels.cuh")

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("samples", type=float, nargs="?", default=1e6)
ap.add_argument("--days", type=int, default=564)
ap.add_argument("--stepper-samples", type=float, default=None)
ap.add_argument("--sha", default=None)
ap.add_argument("--write-counters", default=None, metavar="SOURCE_TAG")
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
d = {n: (v[i], u[i]) for i, n in enumerate(h)}
keys = ["gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
inst = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
print("warp-instr per warp-day", inst / (a.samples * a.days / 32))
for n in h:
    if "issue_stalled" in n and n.endswith(".ratio") and "not_issued" not in n and float(d[n][0]) > 0.05:
        print(f"{n:90s} {d[n][0]}")


def to_bytes(val, unit):
    x = float(val.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


if a.write_counters:
    sha = a.sha
    if sha and os.path.exists(sha):
        sha = open(sha).read().split()[0]
    if not sha:
        sha = hashlib.sha256(open(KERNELS, "rb").read()).hexdigest()
    stepped = a.stepper_samples if a.stepper_samples else a.samples
    out = {
        "kernel": "vx_sim_kernel<SIM_FOLD>",
        "kernels_sha256": sha,
        "source": a.write_counters,
        "workload": f"USA scenario, job of {a.samples:g} samples x {a.days} days; the captured launch stepped {stepped:g}",
        "sample_days_per_launch": stepped * a.days,
        "warp_inst_per_launch": inst,
        "warp_inst_per_warp_day": inst / (stepped * a.days / 32),
        "dram_bytes_per_launch": to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"]),
        "ncu_duration_ms": float(d["gpu__time_duration.sum"][0]),
        "ncu_issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "ncu_threads_per_inst": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
    }
    with open(os.path.join(ROOT, "profiles", "kernel_counters.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("wrote profiles/kernel_counters.json for", sha[:16])
