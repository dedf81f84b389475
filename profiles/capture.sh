#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): plain runs first, then the ncu passes of the same command lines.
#   profiles/capture.sh r01        -> gpurun_out/r01_*.{json,csv,ncu-rep,log}
# Summaries are extracted afterwards on the CPU box with profiles/summarize.py and committed under profiles/.
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extras"
PROF="python profiles/prof_run.py all"
# 1. the bench line itself (never under a profiler)
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { echo "bench failed"; tail -5 $out/${tag}_bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# 2. launch list of the bench command: every kernel with its device time (cold-cache, serialised: compare shares)
$BENCH > $out/${tag}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $BENCH > $out/${tag}_ncu_launch.log 2>&1
# 3. full capture of the hot kernels on BASELINE shapes (prof_run.py: cfg2 assign + loss, cfg3 detect, decode, AP eval):
#    one profiled pass after warm-up (cudaProfilerStart/Stop inside prof_run.py)
$PROF > $out/${tag}_plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"assign_|match_encode|detect_kernel|detect_multi_kernel|decode_kernel|mbl_|wider_eval_kernel" -c 30 \
    -f -o $out/${tag}_full $PROF > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
