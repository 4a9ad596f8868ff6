#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_r2m.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2m.log
B="python bench.py --no-cpu-baseline --steps 60 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:v['ms'] for k,v in d['roofline'].get('stages_one_view', d['roofline'].get('stages')).items()})
c=d.get('also_C1')
if c: print('   also_C1 ms/step %.4f e2e %.4f' % (c['ms_per_step'], c['e2e']['ms_per_step']), {k:v['ms'] for k,v in c['roofline']['stages'].items()})
PY
}
for rep in 1 2; do
$B > $O/r2m_tma$rep.json 2>$O/r2m.err; show $O/r2m_tma$rep.json "TMA rows (rep $rep)"
GSR_NO_TMA=1 $B > $O/r2m_notma$rep.json 2>>$O/r2m.err; show $O/r2m_notma$rep.json "vector loads (rep $rep)"
done
tail -3 $O/r2m.err
