"""Rows f2 / f3 of SURVEY.md §8 pinned to the reference's own modules: tests/golden/slam_golden.npz holds outputs of
utils/slam_utils.py (five loss variants + autograd gradients), utils/pose_utils.py (SO3_exp, V, SE3_exp, update_pose),
utils/camera_utils.py (camera tensors) and the tracking optimiser of utils/slam_frontend.py:132-162, produced by
tests/golden/make_slam_golden.py importing those modules in the build container.

CPU: oracle/slam_ref.py (the restatement the other GPU tests use as checker) must reproduce them.
GPU: gsr_slam_loss, the loss in the forward's epilogue (through slam_loss's arithmetic) and gsr_tracking_step must too."""
import os

import numpy as np
import pytest
import torch

from common import rel_err
from oracle import slam_ref as R

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slam_golden.npz"))
MODES = ["track_rgbd", "track_mono", "map_rgbd", "map_mono", "map_init"]


def _loss_inputs():
    W, H = (int(x) for x in G["loss_shape"])
    flat = torch.from_numpy(G["loss_inputs"])
    sizes = [3 * H * W, H * W, H * W, 3 * H * W, H * W, H * W]
    parts = torch.split(flat, sizes)
    color, depth, opacity, gt, gtd, gmask = (p.reshape(-1, H, W) for p in parts)
    return W, H, color, depth, opacity, gt, gtd, gmask > 0.5


@pytest.mark.parametrize("mode", MODES)
def test_oracle_losses_reproduce_the_reference_module(mode):
    W, H, color, depth, opacity, gt, gtd, gmask = _loss_inputs()
    thr, alpha, a0, b0 = (float(x) for x in G["loss_params"])
    img, dep = color.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    a, b = torch.tensor([a0], requires_grad=True), torch.tensor([b0], requires_grad=True)
    mono = mode.endswith("mono")
    if mode.startswith("track"):
        loss = R.loss_tracking(img, dep, opacity, gt, gtd, gmask, a, b, thr, alpha, mono)
    else:
        loss = R.loss_mapping(img, dep, gt, gtd, a, b, thr, alpha, mono, initialization=(mode == "map_init"))
    loss.backward()
    assert abs(loss.item() - float(G[mode + "_loss"])) <= 1e-7
    np.testing.assert_array_equal(img.grad.numpy(), G[mode + "_dcolor"])       # same torch ops in the same order: identical
    np.testing.assert_array_equal((torch.zeros_like(dep) if dep.grad is None else dep.grad).numpy(), G[mode + "_ddepth"])
    dab = [0.0 if a.grad is None else a.grad.item(), 0.0 if b.grad is None else b.grad.item()]
    np.testing.assert_allclose(dab, G[mode + "_dab"], rtol=1e-6, atol=1e-9)


def test_oracle_se3_exp_reproduces_the_reference_module():
    for i, tau in enumerate(G["se3_taus"]):
        t = torch.from_numpy(tau)
        np.testing.assert_array_equal(R.SE3_exp(t).numpy(), G["se3_exp"][i])
        np.testing.assert_array_equal(R.SO3_exp(t[3:]).numpy(), G["so3_exp"][i])
        np.testing.assert_array_equal(R.V(t[3:]).numpy(), G["V"][i])


def test_oracle_tracking_optimiser_and_pose_update_reproduce_the_reference_modules():
    T0 = torch.from_numpy(G["track_T0"])
    Rm, Tm = T0[:3, :3].clone(), T0[:3, 3].clone()
    proj = torch.from_numpy(G["track_proj"])
    rot, trans = torch.zeros(3, requires_grad=True), torch.zeros(3, requires_grad=True)
    ea, eb = torch.zeros(1, requires_grad=True), torch.zeros(1, requires_grad=True)
    opt = torch.optim.Adam([{"params": [rot], "lr": 0.003}, {"params": [trans], "lr": 0.001}, {"params": [ea], "lr": 0.01}, {"params": [eb], "lr": 0.01}])
    for it, gr in enumerate(G["track_grads"]):
        gr = torch.from_numpy(gr)
        rot.grad, trans.grad, ea.grad, eb.grad = gr[3:6].clone(), gr[0:3].clone(), gr[6:7].clone(), gr[7:8].clone()
        with torch.no_grad():
            opt.step()
            Rm, Tm, conv = R.update_pose(Rm, Tm, trans.detach(), rot.detach())
            rot.zero_(); trans.zero_()
        np.testing.assert_allclose(Rm.numpy(), G["track_R"][it], rtol=0, atol=1e-7)
        np.testing.assert_allclose(Tm.numpy(), G["track_T"][it], rtol=0, atol=1e-7)
        assert conv == bool(G["track_converged"][it])
        wvt, full, center = R.camera_tensors(Rm, Tm, proj)
        np.testing.assert_allclose(wvt.numpy(), G["track_wvt"][it], rtol=0, atol=1e-6)
        np.testing.assert_allclose(full.numpy(), G["track_full"][it], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(center.numpy(), G["track_center"][it], rtol=0, atol=1e-6)
        np.testing.assert_allclose([ea.item(), eb.item()], G["track_exposure"][it], rtol=1e-6, atol=1e-9)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("mode", MODES)
def test_loss_kernel_against_the_reference_module(mode):
    from diff_gaussian_rasterization import slam_ops as S

    W, H, color, depth, opacity, gt, gtd, gmask = _loss_inputs()
    thr, alpha, a0, b0 = (float(x) for x in G["loss_params"])
    mono, tracking, init = mode.endswith("mono"), mode.startswith("track"), mode == "map_init"
    ws = S.LossWorkspace(W, H)
    cu = lambda t: t.cuda().contiguous()
    sums = S.slam_loss(ws, cu(color), cu(depth), cu(opacity), cu(gt), None if mono else cu(gtd),      # initialization only drops the exposure (:92-99)
                       cu(gmask.to(torch.uint8)) if tracking else None, None if init else torch.tensor([a0, b0]).cuda(), thr, alpha,
                       tracking=tracking).cpu().numpy()
    ref = float(G[mode + "_loss"])
    assert abs(sums[0] - ref) <= 2e-6 * abs(ref)
    assert rel_err(ws.dL_dcolor.cpu().numpy(), G[mode + "_dcolor"]) <= 1e-6
    assert rel_err(ws.dL_ddepth.cpu().numpy(), G[mode + "_ddepth"]) <= 1e-6
    if not init:
        np.testing.assert_allclose(sums[1:3], G[mode + "_dab"], rtol=1e-4, atol=1e-8)


@pytest.mark.gpu
def test_tracking_step_kernel_against_the_reference_modules():
    """Adam (slam_frontend.py:132-162) + update_pose (pose_utils.py:76-93) + camera tensors (camera_utils.py:96-109) of eight
    consecutive iterations, the last two with zero gradients."""
    from diff_gaussian_rasterization import slam_ops as S

    T0 = G["track_T0"]
    pose = S.PoseState(T0[:3, :3], T0[:3, 3], G["track_proj"])
    block = torch.zeros(52, dtype=torch.float32, device="cuda")
    for it, gr in enumerate(G["track_grads"]):
        sums = torch.tensor([0.0, gr[6], gr[7], 0.0], dtype=torch.float32, device="cuda")
        S.tracking_step(pose, torch.from_numpy(gr[:6].copy()).cuda(), sums, block)
        rt = pose.RT.cpu().numpy()
        assert rel_err(rt[:9].reshape(3, 3), G["track_R"][it]) <= 2e-6, it
        assert rel_err(rt[9:], G["track_T"][it]) <= 2e-6, it
        np.testing.assert_allclose(pose.exposure.cpu().numpy(), G["track_exposure"][it], rtol=1e-4, atol=2e-6)
        blk = block.cpu().numpy()
        assert rel_err(blk[0:16].reshape(4, 4), G["track_wvt"][it]) <= 2e-6
        assert rel_err(blk[16:32].reshape(4, 4), G["track_full"][it]) <= 2e-6
        np.testing.assert_array_equal(blk[32:48].reshape(4, 4), G["track_proj"])
        assert rel_err(blk[48:51], G["track_center"][it]) <= 1e-5
        st = pose.status.cpu().numpy()
        if not G["track_converged"][:it + 1].any():
            assert st[0] == 0 and st[1] == it + 1
    assert bool(st[0]) == bool(G["track_converged"].any())
    # a frame whose first gradient is exactly zero converges at once (|tau| = 0) and stays where it is
    pose = S.PoseState(T0[:3, :3], T0[:3, 3], G["track_proj"])
    assert bool(G["still_converged"])
    for _ in range(3):
        S.tracking_step(pose, torch.zeros(6, device="cuda"), None, block)
    st = pose.status.cpu().numpy()
    assert st[0] == 1 and st[2] == 1
    rt = pose.RT.cpu().numpy()
    assert rel_err(rt[:9].reshape(3, 3), G["still_R"]) <= 1e-6 and rel_err(rt[9:], G["still_T"]) <= 1e-6
