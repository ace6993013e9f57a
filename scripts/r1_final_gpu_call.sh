#!/bin/bash
# One bounded GPU call at the end of round 1 (about 9 minutes of box time):
#   1-2  A/B of the default blind rotation and the slim-prologue build (same box, same clocks)
#   3    kernel tests against the slim build
#   4    the whole GPU suite against the default build
#   5    the default bench line
#   6    trace check (every arena block vs the plaintext interpretation) with whatever time is left
# Everything lands in gpurun_out/r1z_*; every step has its own timeout so the call cannot hang the box.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
SLIM=$PWD/fhestring_b200/libfhestr_engine_slim.so
SHORT="--steps 6 --warmup 3 --no-cpu-baseline --no-contains"
timeout 150 python bench.py $SHORT > $O/r1z_bench_default_short.json 2> $O/r1z_bench_default_short.err
echo "default short: rc=$? t=$SECONDS"
FHESTR_ENGINE_LIB=$SLIM timeout 100 python bench.py $SHORT > $O/r1z_bench_slim_short.json 2> $O/r1z_bench_slim_short.err
echo "slim short: rc=$? t=$SECONDS"
FHESTR_ENGINE_LIB=$SLIM timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r1z_pytest_slim_kernels.log 2>&1
echo "slim kernel tests: rc=$? t=$SECONDS"; tail -3 $O/r1z_pytest_slim_kernels.log
timeout 240 python -m pytest tests -m gpu -x -q > $O/r1z_pytest_gpu.log 2>&1
echo "gpu suite: rc=$? t=$SECONDS"; tail -3 $O/r1z_pytest_gpu.log
if [ $((540 - SECONDS)) -gt 60 ]; then
  timeout $((540 - SECONDS)) python bench.py > $O/r1z_bench_default.json 2> $O/r1z_bench_default.err
  echo "default bench: rc=$? t=$SECONDS"
fi
if [ $((540 - SECONDS)) -gt 40 ]; then
  timeout $((540 - SECONDS)) python scripts/gpu_trace_check.py ALL --repeat 100 --quiet > $O/r1z_trace_check.log 2>&1
  echo "trace check: rc=$? t=$SECONDS"; tail -2 $O/r1z_trace_check.log
fi
head -c 600 $O/r1z_bench_default_short.json; echo; head -c 600 $O/r1z_bench_slim_short.json; echo
