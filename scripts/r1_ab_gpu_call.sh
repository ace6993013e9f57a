#!/bin/bash
# Second bounded GPU call of round 1 (about 4 minutes): A/B of the FHESTR_BR_* kernel variants on one box, then the
# whole GPU suite, the default bench line and an ncu capture for the fastest one.  Outputs: gpurun_out/r1y_*.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
SHORT="--steps 6 --warmup 3 --no-cpu-baseline --no-contains"
for tag in default slim slim_pf4 slim_pf12 slim_cvt1 slim_cvt2 slim_cvt4 slim_i2f slim_cvt2_i2f; do
  lib=$PWD/fhestring_b200/libfhestr_engine_$tag.so
  [ $tag = default ] && lib=$PWD/fhestring_b200/libfhestr_engine.so
  [ -f $lib ] || { echo "missing $lib"; continue; }
  FHESTR_ENGINE_LIB=$lib timeout 60 python bench.py $SHORT > $O/r1y_ab_$tag.json 2> $O/r1y_ab_$tag.err
  echo "$tag rc=$? t=$SECONDS $(python -c "import json,sys; d=json.loads(open('$O/r1y_ab_$tag.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['ms_per_launch'],3), d['verified_decrypt'])" 2>&1)"
done
WIN=$(python - <<'PY'
import glob, json, os
best, tag = 0, "default"
for f in glob.glob("gpurun_out/r1y_ab_*.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("verified_decrypt") and d["value"] > best:
            best, tag = d["value"], os.path.basename(f)[len("r1y_ab_"):-len(".json")]
    except Exception:
        pass
print(tag)
PY
)
echo "winner: $WIN"; echo $WIN > $O/r1y_winner.txt
WLIB=$PWD/fhestring_b200/libfhestr_engine_$WIN.so
[ $WIN = default ] && WLIB=$PWD/fhestring_b200/libfhestr_engine.so
FHESTR_ENGINE_LIB=$WLIB timeout 120 python -m pytest tests -m gpu -x -q > $O/r1y_pytest_gpu_winner.log 2>&1
echo "gpu suite on $WIN: rc=$? t=$SECONDS"; tail -2 $O/r1y_pytest_gpu_winner.log
FHESTR_ENGINE_LIB=$WLIB timeout 90 python bench.py > $O/r1y_bench_winner.json 2> $O/r1y_bench_winner.err
echo "bench on $WIN: rc=$? t=$SECONDS"
if [ $SECONDS -lt 200 ]; then
  FHESTR_ENGINE_LIB=$WLIB timeout 100 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 4 -c 1 \
    -f -o $O/prof_br_r1y_winner python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-contains > $O/r1y_ncu_full.log 2>&1
  echo "ncu full: rc=$? t=$SECONDS"
fi
if [ $SECONDS -lt 240 ]; then
  FHESTR_ENGINE_LIB=$WLIB timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1y_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-contains > $O/r1y_ncu_launches.log 2>&1
  echo "ncu launches: rc=$? t=$SECONDS"
fi
