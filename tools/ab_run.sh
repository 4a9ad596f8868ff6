#!/bin/bash
# Parity suite, then bench lines for a list of "ENV=VAL:workload[:views]" variants (ENV "-" = product defaults).
#   VARIANTS="-:C1_tum_tracking GSR_NO_PDL_FWD=1:C1_tum_tracking -:C4_large:2" bash tools/ab_run.sh
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
i=0
for V in $VARIANTS; do
i=$((i+1))
IFS=: read -r ENVV WL VIEWS <<< "$V"
EXTRA=""; [ -n "$VIEWS" ] && EXTRA="--views $VIEWS"
( [ "$ENVV" != "-" ] && export $ENVV
python bench.py --workload $WL $EXTRA --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline > $O/abr_$i.json 2>$O/abr.err || tail -5 $O/abr.err )
python - <<PY
import json
d=json.load(open('$O/abr_$i.json'))
r=d['roofline']
print('$V', 'ms/step %.4f' % d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), {k: v['ms'] if isinstance(v, dict) else v for k, v in (r.get('stages') or r.get('stages_one_view')).items()})
PY
done
