"""Turns gpurun_out/launches_*.csv and *.ncu-rep into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> (e.g. r01_v3)"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(tag):
    path = os.path.join(GP, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("fmmb::<unnamed>::", "").replace("void ", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e3)
    with open(os.path.join(OUT, "launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                "`python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (plan build + 5 device matvecs + serialised "
                "roofline steps + host-buffer steps).  Times are cold-cache and serialised: compare shares.\n\n"
                "| kernel | launches | mean us | total us |\n|---|---:|---:|---:|\n" % tag)
        for k, v in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f |\n" % (k[:70], len(v), sum(v) / len(v), sum(v)))
    os.system("cp %s %s" % (path, os.path.join(OUT, "launches_%s.csv" % tag)))


def longest(H, rows):
    i = H.index("gpu__time_duration.sum")
    return max((r for r in rows if len(r) > i), key=lambda r: float(r[i].replace(",", "")))


def report(rep, tag):
    path = os.path.join(GP, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    V = longest(H, rows[2:])             # several launches captured: summarise the longest one
    with open(os.path.join(OUT, rep + ".md"), "w") as f:
        f.write("# %s\n\n`ncu --set full --clock-control none --import-source on`, one launch of the kernel inside "
                "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5` (N=1M, P=8); when the capture holds several "
                "launches of the kernel, the longest one.\n\n" % rep)
        name = V[H.index("Kernel Name")]
        f.write("kernel: `%s`\n\n| metric | unit | value |\n|---|---|---:|\n" % name)
        for k in KEYS:
            if k in H:
                i = H.index(k)
                f.write("| %s | %s | %s |\n" % (k, U[i], V[i]))


def traffic(reps, tag):
    """profiles/traffic_rNN.json: DRAM bytes per launch of the top kernels (read by bench.py for roofline.traffic)."""
    import json
    out = {}
    for key, rep in reps.items():
        path = os.path.join(GP, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode()
        rows = list(csv.reader(io.StringIO(raw)))
        H, U = rows[0], rows[1]
        V = longest(H, rows[2:])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

        def val(name):
            i = H.index(name)
            return float(V[i].replace(",", "")) * scale[U[i]]
        out[key] = {"kernel": V[H.index("Kernel Name")][:80], "source": rep + ".ncu-rep",
                    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                    "duration_ms_under_ncu": float(V[H.index("gpu__time_duration.sum")].replace(",", "")) *
                    {"ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(U[H.index("gpu__time_duration.sum")].replace("second", "s")
                                                          .replace("msecond", "ms"), 1.0)}
    if out:
        json.dump(out, open(os.path.join(OUT, "traffic_%s.json" % tag.split("_")[0]), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    reps = [a for a in sys.argv[2:] if "=" not in a]
    for rep in reps:
        report(rep, tag)
    traffic(dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a), tag)
