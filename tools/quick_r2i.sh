#!/bin/bash
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-also-c1 --steps 40 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['check'])
except Exception as e: print(sys.argv[2], 'FAILED', e)
PY
}
$B > $O/r2i_g.json 2>$O/r2i.err; show $O/r2i_g.json "C2 N=1 graph"; tail -2 $O/r2i.err
$B --no-graph > $O/r2i_e.json 2>$O/r2i.err; show $O/r2i_e.json "C2 N=1 eager"
$B --as-rank-of 8 > $O/r2i_g8.json 2>$O/r2i.err; show $O/r2i_g8.json "rank0-of-8 graph"; tail -2 $O/r2i.err
$B --as-rank-of 8 --no-graph > $O/r2i_e8.json 2>$O/r2i.err; show $O/r2i_e8.json "rank0-of-8 eager"
$B --workload C3_batched_tracking --views 8 --steps 20 > $O/r2i_c3.json 2>$O/r2i.err; show $O/r2i_c3.json "C3 v8 graph"; tail -2 $O/r2i.err
python -m pytest tests/test_engine_gpu.py tests/test_slam_ops_gpu.py tests/test_binning_gpu.py tests/test_slam_golden.py tests/test_script_chain.py tests/test_render_glue.py -m gpu -x -q 2>&1 | tail -4
