#!/usr/bin/env python
"""Turns the raw ncu output of profiles/capture.sh (gpurun_out/<tag>_*) into the committed summaries:

    profiles/<tag>_launches_summary.txt   per-kernel count / total / share of the bench command's launch list
    profiles/<tag>_ncu_summary.txt        key counters of the `ncu --set full` capture, one block per kernel
    profiles/traffic.json                 dram bytes per launch for the kernels bench.py reports a roofline for

Runs on the CPU box (ncu -i reads the report without a GPU):  python profiles/summarize.py r01
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    n = name.replace("jabd::", "")
    return n.split("(")[0]


def launches():
    path = os.path.join(src, tag + "_launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(dst, tag + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 40 --warmup 3 "
                "--no-cpu-baseline --no-extras\n# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("%-60s %8s %12s %10s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.1f %10.2f %6.1f%%\n" % (k[:60], n, t / 1e3, t / n / 1e3, 100 * t / tot))
        # the command also replays per-phase graphs (prep alone, prep+match, encode alone), so launch COUNTS differ per kernel;
        # one step = one launch of each of the three kernels: shares of a step from the per-launch averages
        step = [(k, agg[k][1] / agg[k][0]) for k in ("assign_prep_kernel", "assign_match_kernel", "match_encode_kernel") if k in agg]
        if len(step) == 3:
            st = sum(v for _, v in step)
            f.write("\n# share of one step (avg per-launch time of its three kernels; bench.py's live phases are in <tag>_bench.json 'phases'):\n")
            for k, v in step:
                f.write("#   %-24s %6.2f us  %5.1f%%\n" % (k, v / 1e3, 100 * v / st))
    print("wrote", tag + "_launches_summary.txt")


def full():
    rep = os.path.join(src, tag + "_full.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    traffic = {}
    seen = {}
    with open(os.path.join(dst, tag + "_ncu_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, command: python profiles/prof_run.py all\n"
                "# (one profiled pass after warm-up.  cfg2 target assignment + MultiBox loss fwd/bwd, smooth-L1 and DIoU box terms: 32 x 640^2\n"
                "#  images, 16,800 priors; cfg3 detection: 16 x 1024^2, 43,008 priors; decode: 64 x 1024^2; WIDER AP counters: 256 images)\n"
                "# under ncu every launch is replayed ~40x with caches flushed: durations are not bench numbers\n")
        for r in rows[2:]:
            k = short(r[ki])
            seen[k] = seen.get(k, 0) + 1
            # prof_run.py launches the assignment twice: work list of 64-GT items (a call that runs alone), then of 192-GT items
            # (what the lanes of jabd_assign_batches run) -- both instances of the matching kernel are listed
            second = seen[k] == 2 and k == "assign_match_kernel"
            if seen[k] > 1 and not second:
                continue
            f.write("\n== %s%s\n" % (r[ki][:150], "   [second launch: JABD_ASSIGN_TUNE(192, 192, 100)]" if second else ""))
            rd = wr = None
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("   %-72s %16s %s\n" % (m, r[i], units[i]))
                    if m == "dram__bytes_read.sum":
                        rd = float(r[i]) * SCALE.get(units[i], 1.0)
                    if m == "dram__bytes_write.sum":
                        wr = float(r[i]) * SCALE.get(units[i], 1.0)
            if rd is not None and wr is not None and not second:
                traffic[k + "_bytes_per_launch"] = rd + wr
                traffic[k + "_read_write"] = [rd, wr]
    traffic["how"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch (profiles/%s_ncu_summary.txt); "
                      "writes that are still in the 126 MB L2 when the launch ends are not counted by this counter" % tag)
    json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1, sort_keys=True)
    print("wrote", tag + "_ncu_summary.txt, traffic.json")


launches()
full()
