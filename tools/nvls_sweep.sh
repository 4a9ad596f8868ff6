#!/bin/bash
# gpurun --gpus N -- 'N=8 bash tools/nvls_sweep.sh'
N=${N:-8}
for U in 4 8 2; do for SZ in 7000544 42000544; do
GSR_REDUCE_UNROLL=$U timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/nvls_allreduce_check.py $SZ 2>> gpurun_out/nvls_n$N.err | grep '^{' | sed "s/^/unroll=$U /" | tee -a gpurun_out/nvls_n$N.jsonl
done; done
