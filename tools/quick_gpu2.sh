#!/bin/bash
# quick iteration incl. the C2 window: parity tests + short C1 and C2 benches with stage times
bash tools/quick_gpu.sh
python bench.py --workload C2_replica_mapping --steps 5 --warmup 3 > gpurun_out/c2_n1.json 2> gpurun_out/c2_n1.err; tail -3 gpurun_out/c2_n1.err
python -c "import json; d=json.load(open('gpurun_out/c2_n1.json')); print('C2', d['value'], d['ms_per_step'], d['roofline']['stages_one_view'])"
