#!/bin/bash
O=gpurun_out
python -m pytest tests/test_engine_gpu.py tests/test_slam_ops_gpu.py -m gpu -x -q 2>&1 | tail -4
B="python bench.py --no-cpu-baseline --no-also-c1 --steps 40 --warmup 3"
$B > $O/r2k_g.json 2>$O/r2k.err; python -c "
import json; d=json.load(open('$O/r2k_g.json')); print('C2 N=1 graph ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['check'])"; tail -2 $O/r2k.err
