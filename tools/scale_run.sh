#!/bin/bash
# multi-GPU bench lines on ONE box:  gpurun --gpus 8 -- 'bash tools/scale_run.sh r2b'
# RUNS: space-separated "workload:N:steps[:views]" entries
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
RUNS=${RUNS:-"C2_replica_mapping:8:30 C4_large:8:10 C3_batched_tracking:8:20 C2_replica_mapping:4:30 C2_replica_mapping:2:30"}
for R in $RUNS; do
  IFS=: read WL N STEPS VIEWS <<< "$R"
  S=${WL:0:2}
  EXTRA=""; [ -n "$VIEWS" ] && EXTRA="--views $VIEWS"
  F=$O/scale_${S}${VIEWS:+_v$VIEWS}_n${N}_$TAG
  timeout ${RUN_TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline $EXTRA > $F.json 2> $F.err
  echo "$WL n=$N rc=$?"
  python - <<PY
import json
try:
    d = json.load(open('$F.json'))
    print('  value %.2f %s  ms/step %.4f  e2e ms %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['ms_per_step']), d['details'].get('collective', '')[-40:], d['details'].get('units_per_rank', [''])[0])
except Exception as e:
    print('  no line:', e); print(open('$F.err').read()[-1500:])
PY
done
