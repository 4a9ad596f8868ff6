#!/bin/bash
# quick iteration: parity tests + a short bench with stage times
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err; tail -3 gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms'] for k,v in d['roofline']['stages'].items()})"
