#!/bin/bash
# multi-GPU bench lines on one box:  N=8 bash tools/scale_run.sh  (under gpurun --gpus N)
N=${N:-8}
O=gpurun_out
for WL in ${WLS:-C1_tum_tracking C2_replica_mapping C3_batched_tracking}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --steps ${STEPS:-30} --warmup 3 --no-cpu-baseline > $O/scale_${WL}_n$N.json 2> $O/scale_${WL}_n$N.err || tail -5 $O/scale_${WL}_n$N.err
  python - <<PY
import json
d=json.load(open('$O/scale_${WL}_n$N.json'))
print('$WL n=$N', 'value %.1f %s' % (d['value'], d['unit']), 'ms/step %.4f' % d['ms_per_step'], 'e2e ms %.4f' % d['e2e']['ms_per_step'])
PY
done
