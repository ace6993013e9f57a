#!/bin/bash
# A/B of blind-rotation kernel variants on ONE GPU box (how profiles/r1_final2_ab_variants.md was measured).
#   here:        python fhestring_b200/build.py --variant p8_cvt4 FHESTR_BR_PREFETCH=8 FHESTR_BR_CVT_FP64=4   (per variant)
#   on the box:  gpurun -- 'bash scripts/ab_variants_gpu.sh default p8_cvt4 ...'
# Each tag runs the short PBS bench (about 7 s) against fhestring_b200/libfhestr_engine_<tag>.so ("default" is the
# shipped library); the fastest verified one then runs the whole GPU suite.  Every step has its own timeout.
# Outputs: gpurun_out/ab_<tag>.json, gpurun_out/ab_winner.txt, gpurun_out/ab_pytest_winner.log
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
SHORT="--steps 6 --warmup 3 --no-cpu-baseline --no-contains"
libof() { [ "$1" = default ] && echo $PWD/fhestring_b200/libfhestr_engine.so || echo $PWD/fhestring_b200/libfhestr_engine_$1.so; }
for tag in "$@"; do
  lib=$(libof $tag)
  [ -f $lib ] || { echo "missing $lib"; continue; }
  FHESTR_ENGINE_LIB=$lib timeout 60 python bench.py $SHORT > $O/ab_$tag.json 2> $O/ab_$tag.err
  echo "$tag rc=$? t=$SECONDS $(python -c "import json; d=json.loads(open('$O/ab_$tag.json').read().strip().splitlines()[-1]); print(round(d['value']), 'PBS/s, blind rotation', round(d['roofline']['ms_per_launch'],3), 'ms, verified', d['verified_decrypt'])" 2>&1)"
done
WIN=$(python - <<'PY'
import glob, json, os
best, tag = 0, "default"
for f in glob.glob("gpurun_out/ab_*.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("verified_decrypt") and d["value"] > best:
            best, tag = d["value"], os.path.basename(f)[len("ab_"):-len(".json")]
    except Exception:
        pass
print(tag)
PY
)
echo "winner: $WIN"; echo $WIN > $O/ab_winner.txt
FHESTR_ENGINE_LIB=$(libof $WIN) timeout 180 python -m pytest tests -m gpu -x -q > $O/ab_pytest_winner.log 2>&1
echo "gpu suite on $WIN: rc=$? t=$SECONDS"; tail -2 $O/ab_pytest_winner.log
