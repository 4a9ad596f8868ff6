#!/bin/bash
# A/B of the on-demand threshold on named workloads:  LMS="0 1024" WLS="C1_tum_tracking C4_large" VIEWS=2 bash tools/quick_lazy2.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for WL in ${WLS:-C1_tum_tracking C2_replica_mapping}; do
for LM in ${LMS:-0 1024}; do
V=""; if [ "$WL" = "C4_large" ]; then V="--views ${VIEWS:-2}"; fi
GSR_LAZY_MIN=$LM python bench.py --workload $WL $V --steps ${STEPS:-50} --warmup 3 --no-cpu-baseline --engines 1 > gpurun_out/b_${WL}_$LM.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b_${WL}_$LM.json'))
r=d['roofline']
print('$WL lazy_min=$LM', 'ms/step %.4f' % d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), r.get('stages') or r.get('stages_one_view'))
PY
done; done
