#!/usr/bin/env python
"""Generates tests/golden/slam_golden.npz by IMPORTING the reference's own modules (build container only; the reference
tree does not travel to the GPU box):

    utils/slam_utils.py    get_loss_tracking / get_loss_mapping (all five variants)          :56-128
    utils/pose_utils.py    SO3_exp, V, SE3_exp, update_pose                                  :26-93
    utils/camera_utils.py  Camera.world_view_transform / full_proj_transform / camera_center :96-109
    utils/slam_frontend.py the parameter groups of the tracking optimiser                    :132-162 (torch.optim.Adam)

Every value is computed by the reference code on CPU tensors (float32, like the application) with torch autograd for the
gradients; the only stubs are the containers the functions read from (`config` dict, a viewpoint whose `original_image.cuda()`
returns the CPU tensor).  Run:  python tests/golden/make_slam_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from utils import pose_utils as PU  # noqa: E402
from utils import slam_utils as SU  # noqa: E402
from utils.camera_utils import Camera  # noqa: E402
from gaussian_splatting.utils.graphics_utils import getProjectionMatrix2  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}


class _OnCpu:
    """stands in for a tensor whose .cuda() the reference calls (no GPU in the build container)"""

    def __init__(self, t):
        self.t = t

    def cuda(self):
        return self.t


def images(W, H, seed):
    g = torch.Generator().manual_seed(seed)
    color = torch.rand((3, H, W), generator=g)
    depth = torch.rand((1, H, W), generator=g) * 4
    opacity = torch.rand((1, H, W), generator=g) * 0.2 + 0.85          # straddles the 0.95 opacity mask
    gt = torch.rand((3, H, W), generator=g)
    gt[:, : H // 8] = 0.001                                              # below the rgb boundary threshold
    gtd = torch.rand((1, H, W), generator=g) * 4
    gtd[:, :, : W // 10] = 0.0                                           # invalid depth
    gmask = torch.rand((1, H, W), generator=g) > 0.3
    return color, depth, opacity, gt, gtd, gmask


# ---------------------------------------------------------------- losses (f3)
W, H = 53, 37
color, depth, opacity, gt, gtd, gmask = images(W, H, 3)
thr, alpha, a0, b0 = 0.01, 0.9, 0.07, -0.03
out["loss_inputs"] = np.concatenate([x.reshape(-1).numpy().astype(np.float32) for x in (color, depth, opacity, gt, gtd, gmask.float())])
out["loss_shape"] = np.array([W, H], np.int32)
out["loss_params"] = np.array([thr, alpha, a0, b0], np.float64)
for mode in ("track_rgbd", "track_mono", "map_rgbd", "map_mono", "map_init"):
    mono = mode.endswith("mono")
    config = {"Training": {"monocular": mono, "rgb_boundary_threshold": thr, "alpha": alpha}}
    vp = types.SimpleNamespace(original_image=_OnCpu(gt), depth=gtd[0].numpy(), grad_mask=gmask,
                               exposure_a=torch.tensor([a0], requires_grad=True), exposure_b=torch.tensor([b0], requires_grad=True))
    img, dep = color.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    if mode.startswith("track"):
        loss = SU.get_loss_tracking(config, img, dep, opacity, vp)
    else:
        loss = SU.get_loss_mapping(config, img, dep, vp, opacity, initialization=(mode == "map_init"))
    loss.backward()
    z = lambda g, like: (torch.zeros_like(like) if g is None else g).numpy()
    out[mode + "_loss"] = np.float64(loss.item())
    out[mode + "_dcolor"] = z(img.grad, img)
    out[mode + "_ddepth"] = z(dep.grad, dep)
    out[mode + "_dab"] = np.array([z(vp.exposure_a.grad, vp.exposure_a)[0], z(vp.exposure_b.grad, vp.exposure_b)[0]], np.float64)

# ---------------------------------------------------------------- SE3_exp (f2)
taus = np.array([[0.01, -0.02, 0.03, 0.004, -0.003, 0.002],
                 [0.3, 0.1, -0.2, 0.5, -0.4, 0.3],
                 [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
                 [1e-3, 2e-3, -1e-3, 3e-6, -2e-6, 1e-6],          # angle < 1e-5: the series branch
                 [-0.05, 0.02, 0.01, 0.0, 0.0, 1.2]], np.float32)
out["se3_taus"] = taus
out["se3_exp"] = np.stack([PU.SE3_exp(torch.from_numpy(t)).numpy() for t in taus])
out["so3_exp"] = np.stack([PU.SO3_exp(torch.from_numpy(t[3:])).numpy() for t in taus])
out["V"] = np.stack([PU.V(torch.from_numpy(t[3:])).numpy() for t in taus])

# ---------------------------------------------------------------- tracking optimiser + update_pose + camera tensors (f2)
fx, fy, cx, cy, Wc, Hc = 517.306408, 516.469215, 318.643040, 255.313989, 640, 480
proj = getProjectionMatrix2(znear=0.01, zfar=100.0, fx=fx, fy=fy, cx=cx, cy=cy, W=Wc, H=Hc).transpose(0, 1)   # slam_frontend.py:318-327
c, s = np.cos(0.3), np.sin(0.3)
T0 = torch.tensor([[c, 0, s, 0.1], [0, 1, 0, -0.2], [-s, 0, c, 0.5], [0, 0, 0, 1]], dtype=torch.float32)
cam = Camera(0, None, None, T0.clone(), proj, fx, fy, cx, cy, 0.0, 0.0, Hc, Wc, T0.clone(), device="cpu")
lr = {"cam_rot_delta": 0.003, "cam_trans_delta": 0.001}                                       # configs/rgbd/tum/base_config.yaml
opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": lr["cam_rot_delta"]}, {"params": [cam.cam_trans_delta], "lr": lr["cam_trans_delta"]},
                        {"params": [cam.exposure_a], "lr": 0.01}, {"params": [cam.exposure_b], "lr": 0.01}])
g = torch.Generator().manual_seed(11)
steps = 8
grads, Rs, Ts, expo, conv, wvt, full, center = [], [], [], [], [], [], [], []
for it in range(steps):
    scale = 1e-2 if it < steps - 2 else 0.0          # the last two steps: zero gradients (momentum still moves the pose)
    gtau = torch.randn(6, generator=g) * scale       # rasterizer order: [rho, theta]
    gexp = torch.randn(2, generator=g) * scale
    cam.cam_rot_delta.grad, cam.cam_trans_delta.grad = gtau[3:].clone(), gtau[:3].clone()
    cam.exposure_a.grad, cam.exposure_b.grad = gexp[:1].clone(), gexp[1:].clone()
    with torch.no_grad():
        opt.step()
        converged = PU.update_pose(cam)
    grads.append(torch.cat([gtau, gexp]).numpy())
    Rs.append(cam.R.numpy().copy()); Ts.append(cam.T.numpy().copy())
    expo.append([cam.exposure_a.item(), cam.exposure_b.item()])
    conv.append(bool(converged))
    wvt.append(cam.world_view_transform.detach().numpy().copy())
    full.append(cam.full_proj_transform.detach().numpy().copy())
    center.append(cam.camera_center.detach().numpy().copy())
out["track_T0"] = T0.numpy()
out["track_proj"] = proj.numpy()
out["track_grads"] = np.stack(grads)
out["track_R"] = np.stack(Rs); out["track_T"] = np.stack(Ts)
out["track_exposure"] = np.array(expo, np.float64)
out["track_converged"] = np.array(conv)
out["track_wvt"] = np.stack(wvt); out["track_full"] = np.stack(full); out["track_center"] = np.stack(center)

# a frame whose very first gradient is exactly zero: Adam does not move, |tau| = 0 < 1e-4 -> converged at iteration 0 (pose_utils.py:88)
cam2 = Camera(1, None, None, T0.clone(), proj, fx, fy, cx, cy, 0.0, 0.0, Hc, Wc, T0.clone(), device="cpu")
opt2 = torch.optim.Adam([{"params": [cam2.cam_rot_delta], "lr": 0.003}, {"params": [cam2.cam_trans_delta], "lr": 0.001}])
cam2.cam_rot_delta.grad, cam2.cam_trans_delta.grad = torch.zeros(3), torch.zeros(3)
with torch.no_grad():
    opt2.step()
    out["still_converged"] = np.array(bool(PU.update_pose(cam2)))
out["still_R"], out["still_T"] = cam2.R.numpy().copy(), cam2.T.numpy().copy()

np.savez_compressed(os.path.join(HERE, "slam_golden.npz"), **out)
print("wrote slam_golden.npz:", {k: np.asarray(v).shape for k, v in out.items()})
