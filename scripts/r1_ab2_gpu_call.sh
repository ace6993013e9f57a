#!/bin/bash
# Last (about one minute) GPU call of round 1: neighbours of the winning variant, cvt4 repeated as the noise estimate.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
SHORT="--steps 6 --warmup 3 --no-cpu-baseline --no-contains"
for tag in slim_cvt4 slim_cvt4_i2f slim_cvt8 slim_cvt3 slim_cvt6 slim_cvt4; do
  [ $SECONDS -gt 62 ] && break
  lib=$PWD/fhestring_b200/libfhestr_engine_$tag.so
  FHESTR_ENGINE_LIB=$lib timeout 30 python bench.py $SHORT > $O/r1x_ab_${tag}_$SECONDS.json 2> /dev/null
  echo "$tag rc=$? t=$SECONDS $(python -c "import json,glob; f=sorted(glob.glob('$O/r1x_ab_${tag}_*.json'))[-1]; d=json.loads(open(f).read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['ms_per_launch'],3), d['verified_decrypt'])" 2>&1)"
done
