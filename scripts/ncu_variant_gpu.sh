#!/bin/bash
# One `ncu --set full` capture of the throughput blind-rotation kernel of a library variant (build.py --variant TAG),
# after a plain run of the same command; the raw page is exported on the box (gpurun_out/ncu_<tag>_raw.csv).
#   gpurun -- 'bash scripts/ncu_variant_gpu.sh compact [more tags]'
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strings"
libof() { [ "$1" = default ] && echo $PWD/fhestring_b200/libfhestr_engine.so || echo $PWD/fhestring_b200/libfhestr_engine_$1.so; }
for tag in "$@"; do
  export FHESTR_ENGINE_LIB=$(libof $tag)
  timeout 120 $BENCH > $O/ncu_${tag}_plain.json 2> $O/ncu_${tag}_plain.err || { echo "$tag: plain run failed"; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 3 -c 1 -f -o $O/ncu_$tag $BENCH > $O/ncu_${tag}.log 2>&1
  echo "$tag ncu rc=$?"
  ncu -i $O/ncu_$tag.ncu-rep --page raw --csv > $O/ncu_${tag}_raw.csv 2>/dev/null
done
