#!/bin/bash
O=gpurun_out
python -m pytest tests/test_config_scale_gpu.py tests/test_engine_gpu.py -m gpu -x -q > $O/pytest_dc.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_dc.log
B="python bench.py --no-cpu-baseline --steps 40 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:v['ms'] for k,v in d['roofline']['stages_one_view'].items()})
c=d.get('also_C1')
if c: print('   also_C1 ms/step %.4f e2e %.4f' % (c['ms_per_step'], c['e2e']['ms_per_step']))
PY
}
$B > $O/dc_on.json 2>$O/dc.err; show $O/dc_on.json "C2 default"


tail -3 $O/dc.err
