#!/usr/bin/env python
"""N-rank check + timing of the NVSwitch gradient reduction (gsr_window_allreduce) against dist.all_reduce:
    timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/nvls_allreduce_check.py [n_floats]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = "cuda:%d" % local
dist.init_process_group("nccl", device_id=torch.device(dev))
from diff_gaussian_rasterization.window import SwitchReducer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 7_000_544
out = {"world": world, "n_floats": n}
red = SwitchReducer.create(n, dev)
if red is None:
    out["unavailable"] = "symmetric / multicast memory could not be set up: %s" % SwitchReducer.last_error
else:
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    worst = 0.0
    for rep in range(3):
        x = torch.randn(n, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        red.buffer.copy_(x)
        torch.cuda.synchronize()
        dist.barrier()
        red.all_reduce()
        torch.cuda.synchronize()
        worst = max(worst, float((red.buffer - ref).abs().max() / ref.abs().max()))
    out["max_rel_err_vs_nccl"] = worst
    out["timed_out"] = red.timed_out()
    # every rank holds the same bits?
    chk = red.buffer.double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    out["identical_on_all_ranks"] = all(float(a) == float(lst[0]) for a in lst)
    st = torch.cuda.current_stream()
    for name, fn in (("switch_kernel_ms", red.all_reduce), ("nccl_ms", lambda: dist.all_reduce(red.buffer))):
        for ctas in ((16, 32, 64, 148) if name.startswith("switch") else (0,)):
            if ctas:
                red.ctas = ctas
            for _ in range(5):
                fn()
            ts = []
            for _ in range(20):
                dist.barrier()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st); fn(); b.record(st)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name + ("_%dctas" % ctas if ctas else "")] = round(float(t), 4)
    out["timed_out_after_timing"] = red.timed_out()
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
