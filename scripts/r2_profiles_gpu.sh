#!/bin/bash
# Round-2 profile captures on ONE B200 (how the profiles/r2_* files were made).  Every ncu run follows a plain run of
# the same command line that exited 0.  usage (from the repo root on the GPU box): bash scripts/r2_profiles_gpu.sh
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strings"
timeout 120 $BENCH > $O/r2_prof_bench_plain.json 2> $O/r2_prof_bench_plain.err &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_bench_4096.csv $BENCH > $O/r2_prof_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 3 -c 1 -o $O/prof_r2_blind_rotate $BENCH > $O/r2_prof_ncu2.log 2>&1
echo "throughput kernel rc=$?"
LAT="python scripts/gpu_latency.py 148 296"
timeout 120 $LAT > $O/r2_prof_lat_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_wide_kernel -s 2 -c 1 -o $O/prof_r2_wide_single $LAT > $O/r2_prof_ncu3.log 2>&1
echo "latency single rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_wide2_kernel -s 2 -c 1 -o $O/prof_r2_wide_pair $LAT > $O/r2_prof_ncu4.log 2>&1
echo "latency pair (296) rc=$?"
FULL="python scripts/gpu_latency.py 4096"
timeout 120 $FULL > $O/r2_prof_full_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_wide2_kernel -s 2 -c 1 -o $O/prof_r2_wide_pair_4096 $FULL > $O/r2_prof_ncu5.log 2>&1
echo "latency pair (4096) rc=$?"
tail -2 $O/r2_prof_full_plain.log
