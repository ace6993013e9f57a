#!/bin/bash
# Round-2 (second half) captures of the throughput kernel with the rolled step + tensor-memory twiddles on ONE B200.
# Every ncu run follows a plain run of the same command line that exited 0.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strings"
timeout 120 $BENCH > $O/r2b_prof_bench_plain.json 2> $O/r2b_prof_bench_plain.err &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches_bench_4096.csv $BENCH > $O/r2b_prof_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 3 -c 1 -f -o $O/prof_r2b_blind_rotate $BENCH > $O/r2b_prof_ncu2.log 2>&1
echo "throughput kernel rc=$?"
ncu -i $O/prof_r2b_blind_rotate.ncu-rep --page raw --csv > $O/r2b_blind_rotate_ncu_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ks_gemm_tc_kernel -s 3 -c 1 -f -o $O/prof_r2b_ks_gemm_tc $BENCH > $O/r2b_prof_ncu3.log 2>&1
echo "keyswitch GEMM rc=$?"
ncu -i $O/prof_r2b_ks_gemm_tc.ncu-rep --page raw --csv > $O/r2b_ks_gemm_tc_ncu_raw.csv 2>/dev/null
