"""gsr_window_allreduce (csrc/window_reduce.cu: the window's gradient sum inside the NVSwitch, multimem.ld_reduce + multimem.st over
symmetric memory) against dist.all_reduce (NCCL) on two ranks -- needs two GPUs with multicast memory, so it is skipped on the
one-GPU test box; run it with `gpurun --gpus 2 -- 'python -m pytest tests/test_switch_reduce_gpu.py -m gpu -q'`
(results of the 2- and 8-GPU runs: profiles/r2_switch_reduce_n*.json*)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n_floats", [1_000_003 // 4 * 4 + 96, 7_000_544])
def test_switch_reduction_matches_nccl_on_two_ranks(n_floats):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "nvls_allreduce_check.py"), str(n_floats)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    if "unavailable" in out:
        pytest.skip(out["unavailable"])
    assert out["world"] == 2 and not out["timed_out"] and not out["timed_out_after_timing"]
    assert out["identical_on_all_ranks"]
    assert out["max_rel_err_vs_nccl"] <= 1e-6
