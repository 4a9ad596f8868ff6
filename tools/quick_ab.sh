#!/bin/bash
# A/B of an environment switch on a workload:  VAR=GSR_NO_OVERLAP VALS="1 0" WL=C1_tum_tracking bash tools/quick_ab.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for V in ${VALS:-1 0}; do
if [ "$V" = "0" ]; then unset ${VAR}; else export ${VAR}=$V; fi
python bench.py --workload ${WL:-C1_tum_tracking} --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline > gpurun_out/ab_$V.json 2>gpurun_out/ab.err || tail -5 gpurun_out/ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab_$V.json'))
r=d['roofline']
print('${VAR}=$V', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e'].get('ms_per_step'), {k: v['ms'] if isinstance(v, dict) else v for k, v in (r.get('stages') or r.get('stages_one_view')).items()})
PY
done
