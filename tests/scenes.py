"""Seeded synthetic workloads of the shapes BASELINE.json names.

Camera conventions follow the reference exactly:
  * getProjectionMatrix2      gaussian_splatting/utils/graphics_utils.py:72-93
  * world_view_transform = W2C^T, full_proj_transform = W2C^T @ P^T, camera_center
                              utils/camera_utils.py:96-109
  * tanfov = tan(FoV/2), FoV = 2 atan(dim / 2f)   gaussian_renderer/__init__.py:55-56
so the 4x4 tensors handed to GaussianRasterizationSettings are row-major transposes
(= column-major matrices), as the kernels expect.
Everything here is numpy on the host; `to_torch` moves a scene to a device.
"""
import math

import numpy as np

SH_C0 = 0.28209479177387814

# name -> (W, H, fx, fy, cx, cy, P, views)
CONFIGS = {
    # LDSC.py:1406-1410 camera, 15 synthetic Gaussians, SH degree 3
    "C0_script": dict(W=640, H=480, fx=577.5, fy=577.5, cx=319.5, cy=239.5, P=15, V=1, sh_degree=3),
    # configs/rgbd/tum/fr1_desk.yaml:6-17
    "C1_tum_tracking": dict(W=640, H=480, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, P=100_000, V=1, sh_degree=0),
    # configs/rgbd/replica/base_config.yaml:17-28
    "C2_replica_mapping": dict(W=1200, H=680, fx=600.0, fy=600.0, cx=599.5, cy=339.5, P=500_000, V=10, sh_degree=0),
    "C3_batched_tracking": dict(W=640, H=480, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, P=300_000, V=64, sh_degree=0),
    "C4_large": dict(W=1920, H=1080, fx=1000.0, fy=1000.0, cx=959.5, cy=539.5, P=3_000_000, V=32, sh_degree=0),
}


def projection_matrix2(znear, zfar, cx, cy, fx, fy, W, H):
    """graphics_utils.py:72-93 (float32 result like torch.zeros(4,4))."""
    left = ((2 * cx - W) / W - 1.0) * W / 2.0
    right = ((2 * cx - W) / W + 1.0) * W / 2.0
    top = ((2 * cy - H) / H + 1.0) * H / 2.0
    bottom = ((2 * cy - H) / H - 1.0) * H / 2.0
    left, right = znear / fx * left, znear / fx * right
    top, bottom = znear / fy * top, znear / fy * bottom
    Pm = np.zeros((4, 4), np.float32)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def so3_exp(theta):
    theta = np.asarray(theta, np.float64)
    a = np.linalg.norm(theta)
    Wm = np.array([[0, -theta[2], theta[1]], [theta[2], 0, -theta[0]], [-theta[1], theta[0], 0]])
    if a < 1e-5:
        return np.eye(3) + Wm + 0.5 * Wm @ Wm
    return np.eye(3) + math.sin(a) / a * Wm + (1 - math.cos(a)) / a**2 * Wm @ Wm


def se3_exp(tau):
    """utils/pose_utils.py:61-73; tau = [rho, theta]."""
    tau = np.asarray(tau, np.float64)
    rho, theta = tau[:3], tau[3:]
    a = np.linalg.norm(theta)
    Wm = np.array([[0, -theta[2], theta[1]], [theta[2], 0, -theta[0]], [-theta[1], theta[0], 0]])
    if a < 1e-5:
        V = np.eye(3) + 0.5 * Wm + Wm @ Wm / 6.0
    else:
        V = np.eye(3) + Wm * ((1 - math.cos(a)) / a**2) + Wm @ Wm * ((a - math.sin(a)) / a**3)
    T = np.eye(4)
    T[:3, :3] = so3_exp(theta)
    T[:3, 3] = V @ rho
    return T


def make_camera(W, H, fx, fy, cx, cy, w2c=None, znear=0.01, zfar=100.0):
    """Camera dict holding the GaussianRasterizationSettings matrix fields (float32)."""
    if w2c is None:
        w2c = np.eye(4)
    w2c = np.asarray(w2c, np.float64).astype(np.float32)
    view = np.ascontiguousarray(w2c.T)                              # world_view_transform
    proj_raw = np.ascontiguousarray(projection_matrix2(znear, zfar, cx, cy, fx, fy, W, H).T)
    full = (view @ proj_raw).astype(np.float32)                     # full_proj_transform (bmm in fp32)
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    fovx = 2 * math.atan(W / (2 * fx))
    fovy = 2 * math.atan(H / (2 * fy))
    return dict(image_width=int(W), image_height=int(H), tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5),
                viewmatrix=view, projmatrix=full, projmatrix_raw=proj_raw, campos=campos,
                w2c=w2c, fx=fx, fy=fy, cx=cx, cy=cy)


def base_pose():
    """A fixed non-trivial world-to-camera pose (rotation about a skew axis + translation)."""
    T = np.eye(4)
    T[:3, :3] = so3_exp([0.12, -0.31, 0.07])
    T[:3, 3] = [0.3, -0.2, 0.5]
    return T


def arc_poses(V, radius=0.5, seed=2, base=None):
    """V keyframe poses on a 0.5 m arc around the base pose (mapping window)."""
    base = base_pose() if base is None else base
    rng = np.random.default_rng(seed)
    out = []
    for i in range(V):
        ang = (i / max(V - 1, 1) - 0.5) * 0.6
        tau = np.array([radius * math.sin(ang), 0.02 * rng.standard_normal(), radius * (1 - math.cos(ang)),
                        0.0, -ang * 0.5, 0.0])
        out.append(se3_exp(tau) @ base)
    return out


def noisy_poses(V, sigma_rho=0.02, sigma_theta=math.radians(1.0), seed=2, base=None):
    """V candidate poses = Exp(noise) * base (batched tracking, C3)."""
    base = base_pose() if base is None else base
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(V):
        tau = np.concatenate([sigma_rho * rng.standard_normal(3), sigma_theta * rng.standard_normal(3)])
        out.append(se3_exp(tau) @ base)
    return out


def make_gaussians(P, cam, seed=0, sh_degree=0, w2c=None):
    """Seeded Gaussian cloud in front of camera `cam` (SURVEY §8(d) distribution)."""
    rng = np.random.default_rng(seed)
    W, H, fx, fy, cx, cy = cam["image_width"], cam["image_height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    z = rng.uniform(0.5, 6.0, P)
    near = rng.random(P) < 0.02
    z[near] = rng.uniform(0.02, 0.2, int(near.sum()))          # exercise the near-plane cull
    ndc = rng.uniform(-1.15, 1.15, (P, 2))                      # some off-screen
    px = (ndc[:, 0] + 1.0) * 0.5 * W - 0.5
    py = (ndc[:, 1] + 1.0) * 0.5 * H - 0.5
    xc = (px - cx) / fx * z
    yc = (py - cy) / fy * z
    pc = np.stack([xc, yc, z, np.ones(P)], 1)
    w2c = np.asarray(cam["w2c"] if w2c is None else w2c, np.float64)
    pw = (np.linalg.inv(w2c) @ pc.T).T[:, :3]
    scales = np.exp(np.log(0.01 * np.maximum(z, 0.05))[:, None] + 0.4 * rng.standard_normal((P, 3)))
    q = rng.standard_normal((P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = 1.0 / (1.0 + np.exp(-1.5 * rng.standard_normal((P, 1))))
    M = (sh_degree + 1) ** 2
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0, 1, (P, 3)) - 0.5) / SH_C0      # RGB2SH, sh_utils.py:121-122
    if M > 1:
        shs[:, 1:, :] = 0.15 * rng.standard_normal((P, M - 1, 3))
    f = np.float32
    return dict(means3D=pw.astype(f), scales=scales.astype(f), rotations=q.astype(f), opacities=opac.astype(f),
                shs=np.ascontiguousarray(shs.astype(f)), sh_degree=int(sh_degree))


def make_scene(name_or_cfg, seed=0, P=None, w2c=None, bg=(0.0, 0.0, 0.0)):
    """One view of a named config: camera + Gaussians + settings scalars -> scene dict."""
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    base = base_pose()
    cam_gen = make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], base)
    g = make_gaussians(cfg["P"] if P is None else P, cam_gen, seed=seed, sh_degree=cfg.get("sh_degree", 0))
    cam = cam_gen if w2c is None else make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], w2c)
    sc = dict(g)
    sc.update({k: cam[k] for k in ("image_width", "image_height", "tanfovx", "tanfovy", "viewmatrix", "projmatrix",
                                   "projmatrix_raw", "campos")})
    sc.update(bg=np.asarray(bg, np.float32), scale_modifier=1.0, prefiltered=False, debug=False)
    return sc


def with_camera(sc, cam):
    out = dict(sc)
    out.update({k: cam[k] for k in ("image_width", "image_height", "tanfovx", "tanfovy", "viewmatrix", "projmatrix",
                                    "projmatrix_raw", "campos")})
    return out


def make_pixel_grads(W, H, seed=1):
    """L1-mean-like upstream gradients (utils/slam_utils.py:70-71,87-88): sign noise / N."""
    rng = np.random.default_rng(seed)
    dc = np.sign(rng.standard_normal((3, H, W))).astype(np.float32) / np.float32(3 * H * W)
    dd = np.sign(rng.standard_normal((1, H, W))).astype(np.float32) / np.float32(H * W)
    return dc, dd


def to_torch(sc, device):
    import torch

    out = {}
    for k, v in sc.items():
        out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device) if isinstance(v, np.ndarray) else v
    return out
