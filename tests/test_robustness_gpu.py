"""Failure paths of the no-sync step: tiny / empty binning workspaces, the bounded spin of the compositing
backward, and replays of a tracking iteration behind convergence."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import rel_err

pytestmark = pytest.mark.gpu


def _small():
    import scenes as S
    from diff_gaussian_rasterization.engine import RasterEngine

    cfg = dict(W=160, H=128, fx=150.0, fy=150.0, cx=80.0, cy=64.0, P=3000, sh_degree=0)
    sc = S.make_scene(cfg, seed=3)
    sc["scales"] = sc["scales"] * 2.0
    t = S.to_torch(sc, "cuda")
    eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                       cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)
    cam = S.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], S.base_pose())
    eng.set_camera(RasterEngine.pack_camera(*(torch.from_numpy(cam[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda())
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"], seed=4)
    eng.dL_dcolor.copy_(torch.from_numpy(dc)); eng.dL_ddepth.copy_(torch.from_numpy(dd))
    return cfg, sc, eng


def _shrink(eng, capacity):
    from diff_gaussian_rasterization import _cabi

    L = _cabi.load()
    eng.capacity = int(capacity)
    eng.bin_bytes = L.gsr_binning_bytes(eng.P, eng.W, eng.H, eng.capacity)
    eng.binning = torch.empty((max(eng.bin_bytes, 1),), dtype=torch.uint8, device=eng.dev)
    eng.graph_fwd = eng.graph_bwd = eng.graph_all = None


def test_nosync_forward_rejects_an_empty_binning_workspace():
    """capacity 0 with visible Gaussians used to walk non-empty ranges through an empty point_list (ADVICE: medium)."""
    cfg, sc, eng = _small()
    R = eng.calibrate()
    assert R > 0
    _shrink(eng, 0)
    with pytest.raises(RuntimeError, match="capacity must be positive"):
        eng.launch_forward()
    torch.cuda.synchronize()


@pytest.mark.parametrize("capacity", [1, 37])
def test_nosync_forward_with_a_tiny_workspace_overflows_cleanly_and_recovers(capacity):
    cfg, sc, eng = _small()
    R = eng.calibrate()
    eng.step(use_graph=False)
    torch.cuda.synchronize()
    want_tau, want_color = eng.g_tau.clone(), eng.color.clone()
    _shrink(eng, capacity)
    eng.launch_forward()
    eng.launch_backward()
    torch.cuda.synchronize()                      # no fault: every list access stays inside the workspace
    need, ov = eng.header()
    assert ov and need == R
    got = eng.step_checked(use_graph=False)       # grows the workspace and re-runs
    assert got == R
    assert torch.equal(eng.color, want_color) and rel_err(eng.g_tau.cpu().numpy(), want_tau.cpu().numpy()) <= 1e-5


def test_backward_rejects_a_binning_capacity_out_of_range():
    """The capacity handed to the backward carves up the binning workspace: a negative one must not turn into a huge offset."""
    cfg, sc, eng = _small()
    eng.calibrate()
    eng.step(use_graph=False)
    cap = eng.capacity
    try:
        eng.capacity = -1
        with pytest.raises(Exception, match="capacity out of range"):
            eng.launch_backward()
    finally:
        eng.capacity = cap
    eng.launch_backward()
    torch.cuda.synchronize()
    assert not eng.header()[1]


def test_backward_spin_timeout_is_reported_as_an_error_not_as_overflow():
    """upstream_ready that never arrives: the compositing backward gives up after its bounded spin, flags the header's
    spin_timeout word (not the capacity-overflow word) and the host sees GSR_ERR_TIMEOUT."""
    cfg, sc, eng = _small()
    eng.calibrate()
    never = torch.zeros(1, dtype=torch.int32, device="cuda")
    eng.launch_forward()
    eng.launch_backward(upstream_ready=never)
    torch.cuda.synchronize()
    st = eng.status()
    assert st["spin_timeout"] == 2 and st["overflow"] == 0
    with pytest.raises(RuntimeError, match="gave up waiting"):
        eng.header()
    assert float(eng.grad_flat.abs().max()) == 0.0      # tiles that gave up contribute nothing
    # the next forward clears the word; a flag that is set lets the step through
    ready = torch.ones(1, dtype=torch.int32, device="cuda")
    eng.launch_forward()
    eng.launch_backward(upstream_ready=ready)
    torch.cuda.synchronize()
    assert eng.status()["spin_timeout"] == 0
    eng.header()
    assert float(eng.grad_flat.abs().max()) > 0.0


def test_tracking_step_is_a_noop_once_converged():
    """The reference breaks out of its loop on the converged iteration (slam_frontend.py:180,192-193): replays of the captured
    iteration behind it leave pose, exposure, Adam state and counters untouched (ADVICE: TrackingLoop polls every 10)."""
    import scenes as S
    from diff_gaussian_rasterization import slam_ops as SO

    cam = S.make_camera(64, 48, 60.0, 60.0, 31.5, 23.5, S.base_pose())
    pose = SO.PoseState(cam["w2c"][:3, :3], cam["w2c"][:3, 3], cam["projmatrix_raw"], device="cuda")
    block = torch.zeros(52, dtype=torch.float32, device="cuda")
    tau = torch.tensor([1e-3, -2e-3, 5e-4, 1e-3, 2e-3, -1e-3], dtype=torch.float32, device="cuda")
    sums = torch.tensor([0.1, 1e-3, -2e-3, 0.0], dtype=torch.float32, device="cuda")
    SO.tracking_step(pose, tau, sums, block, converged_threshold=1e9)       # converges on its first iteration
    torch.cuda.synchronize()
    st = pose.status.cpu().numpy()
    assert st[0] == 1 and st[1] == 1 and st[2] == 1
    keep = [x.clone() for x in (pose.RT, pose.exposure, pose.adam, pose.status, block)]
    for _ in range(3):
        SO.tracking_step(pose, tau, sums, block, converged_threshold=1e9)
    torch.cuda.synchronize()
    for a, b in zip(keep, (pose.RT, pose.exposure, pose.adam, pose.status, block)):
        assert torch.equal(a, b)
    pose.reset_optimizer()                                                   # next frame: steps apply again
    SO.tracking_step(pose, tau, sums, block, converged_threshold=0.0)
    torch.cuda.synchronize()
    assert not torch.equal(keep[0], pose.RT) and int(pose.status[1]) == 1
