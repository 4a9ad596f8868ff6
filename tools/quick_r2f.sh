#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_r2f.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2f.log
B="python bench.py --no-cpu-baseline --steps 40 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:v['ms'] for k,v in d['roofline'].get('stages_one_view', d['roofline'].get('stages')).items()})
c=d.get('also_C1')
if c: print('   also_C1 ms/step %.4f e2e %.4f' % (c['ms_per_step'], c['e2e']['ms_per_step']), {k:v['ms'] for k,v in c['roofline']['stages'].items()})
PY
}
$B > $O/r2f_base.json 2>$O/r2f.err; show $O/r2f_base.json "C2 spatial order"
GSR_NO_SPATIAL_ORDER=1 $B --no-also-c1 > $O/r2f_old.json 2>>$O/r2f.err; show $O/r2f_old.json "C2 plain scatter"
$B --no-also-c1 --as-rank-of 8 > $O/r2f_rank8.json 2>>$O/r2f.err; show $O/r2f_rank8.json "C2 rank0-of-8 share"
$B --no-also-c1 --workload C4_large --views 2 --steps 10 > $O/r2f_c4.json 2>>$O/r2f.err; show $O/r2f_c4.json "C4 v2 new"
GSR_NO_SPATIAL_ORDER=1 $B --no-also-c1 --workload C4_large --views 2 --steps 10 > $O/r2f_c4_plain.json 2>>$O/r2f.err; show $O/r2f_c4_plain.json "C4 v2 plain"
$B --no-also-c1 --workload C3_batched_tracking --views 8 --steps 20 > $O/r2f_c3.json 2>>$O/r2f.err; show $O/r2f_c3.json "C3 v8 new"
GSR_NO_SPATIAL_ORDER=1 $B --no-also-c1 --workload C3_batched_tracking --views 8 --steps 20 > $O/r2f_c3_old.json 2>>$O/r2f.err; show $O/r2f_c3_old.json "C3 v8 plain"
tail -3 $O/r2f.err
