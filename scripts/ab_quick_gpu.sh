#!/bin/bash
# Short A/B of kernel variants on one GPU box: the PBS bench only (no suite), one line per tag.
#   gpurun -- 'bash scripts/ab_quick_gpu.sh default q15 ...'   -> gpurun_out/abq_<tag>.json
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
SHORT="--steps 6 --warmup 3 --no-cpu-baseline --no-contains"
libof() { [ "$1" = default ] && echo $PWD/fhestring_b200/libfhestr_engine.so || echo $PWD/fhestring_b200/libfhestr_engine_$1.so; }
for tag in "$@"; do
  lib=$(libof $tag)
  [ -f $lib ] || { echo "missing $lib"; continue; }
  FHESTR_ENGINE_LIB=$lib timeout 90 python bench.py $SHORT > $O/abq_$tag.json 2> $O/abq_$tag.err
  echo "$tag rc=$? t=$SECONDS $(python -c "import json; d=json.loads(open('$O/abq_$tag.json').read().strip().splitlines()[-1]); print(round(d['value']), 'PBS/s, blind rotation', round(d['roofline']['ms_per_launch'],3), 'ms, frac', round(d['roofline']['frac'],4), 'verified', d['verified_decrypt'])" 2>&1 | tail -1)"
done
