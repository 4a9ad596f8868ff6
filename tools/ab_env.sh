#!/bin/bash
# A/B of one environment switch on the default bench line:  VAR=GSR_NO_PDL_FWD bash tools/ab_env.sh
O=gpurun_out
VAR=${VAR:-GSR_NO_PDL_FWD}
python -m pytest tests -m gpu -x -q > $O/pytest_ab2.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_ab2.log
B="python bench.py --no-cpu-baseline --steps 60 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))
c=d.get('also_C1')
if c: print('   also_C1 ms/step %.4f e2e %.4f' % (c['ms_per_step'], c['e2e']['ms_per_step']))
PY
}
for rep in 1 2; do
$B > $O/ab2_on$rep.json 2>$O/ab2.err; show $O/ab2_on$rep.json "default (rep $rep)"
env $VAR=1 $B > $O/ab2_off$rep.json 2>>$O/ab2.err; show $O/ab2_off$rep.json "$VAR=1 (rep $rep)"
done
$B --no-also-c1 --as-rank-of 8 > $O/ab2_r8.json 2>>$O/ab2.err; show $O/ab2_r8.json "rank0-of-8 default"
env $VAR=1 $B --no-also-c1 --as-rank-of 8 > $O/ab2_r8off.json 2>>$O/ab2.err; show $O/ab2_r8off.json "rank0-of-8 $VAR=1"
tail -3 $O/ab2.err
