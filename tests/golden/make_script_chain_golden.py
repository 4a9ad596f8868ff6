#!/usr/bin/env python
"""Generates tests/golden/script_chain_c0.npz by EXECUTING the reference's analytic-Jacobian script chain
(/root/reference/Loss_Derivative_script_compare.py) on the C0 configuration of BASELINE.json / SURVEY.md 8(d):
640x480, fx = fy = 577.5, cx = 319.5, cy = 239.5 (:1406-1410), pose = w2c_gt @ T_noise from Jacob_test_result/*.txt
(:1396-1425), 15 Gaussians, SH degree 3, the script's own loss gradient (sign of the L1 residual under a mask,
:1225-1233).  The reference's `optimized_params_small.pt` and NOCS images are not in the repository, so the 15 Gaussians and
the ground-truth images are synthetic (seeded); the rendered images the script takes from the CUDA rasterizer come from
the CPU oracle here (they only enter through sign(rendered - gt)).

The module cannot be imported (module-level imports of the CUDA extension, cv2, plotly ...), so -- like
make_kat_golden.py -- the source is SLICED and exec'd unchanged:
    functions   `def get_render_settings` (:54-137) and `def eval_sh` ... up to `if __name__` (:349-1352: SH forward /
                backward, pi / hat / GetAnalyticalJcobian / Get_dcovI_dJ :591-760, projection :764-971,
                dense dL/dmu_I, dL/dSigma_I :1173-1351)
    main chain  :1524-1695 (projection call, dense gradients, Jacobians of all Gaussians, chain rule incl. the depth
                :1627-1634 and SH :1636-1660 pose terms, np.save of the four Jacob_test_result/*.npy)
The only edit is the text replacement device='cuda' -> device='cpu' (no GPU in the build container).  The chain runs in a
temporary directory, so its np.save calls produce fresh Jacob_test_result/*.npy files whose SCHEMA (shapes, dtypes) is
checked against the files the reference ships.

Run (build container only):  python tests/golden/make_script_chain_golden.py
"""
import contextlib
import io
import os
import sys
import tempfile
import textwrap
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("GS_REFERENCE", "/root/reference")
REF = os.path.join(REF_ROOT, "Loss_Derivative_script_compare.py")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_ROOT)
from utils.camera_utils import Camera  # noqa: E402  (the reference's own Camera: full_proj_transform, camera_center)
from oracle.gs_oracle import Oracle  # noqa: E402

src = open(REF).read()
lines = src.split("\n")
fn_settings = src[src.index("def get_render_settings("):src.index("@dataclass\nclass pipeline_params")]
fn_block = src[src.index("def eval_sh(deg, sh, dirs):"):src.index('if __name__ == "__main__":')]
i0 = next(i for i, l in enumerate(lines) if l.strip().startswith("gaussians_sorted_by_depth = OrderGaussiansByDepth(xyz_cam)"))
i1 = next(i for i, l in enumerate(lines) if l.strip().startswith("np.save('./Jacob_test_result/dL_dtau.npy'"))
chain = textwrap.dedent("\n".join(lines[i0:i1 + 1]))
to_cpu = lambda s: s.replace("device='cuda'", "device='cpu'")

ns = {"np": np, "torch": torch, "math": __import__("math"), "Dict": dict, "Any": object}
exec(compile(to_cpu(fn_settings), REF, "exec"), ns)
exec(compile(to_cpu(fn_block), REF, "exec"), ns)

# ------------------------------------------------------------------ C0 inputs
w, h = 640, 480
# :1406-1410 (float32 there; float64 here because a CPU torch tensor refuses a numpy.float32 scalar in item assignment --
# the values are exactly representable in both)
cam_intrinsics = np.array([[577.5, 0, 319.5], [0, 577.5, 239.5], [0, 0, 1]], dtype=np.float64)
fx, fy, cx, cy = (cam_intrinsics[0, 0], cam_intrinsics[1, 1], cam_intrinsics[0, 2], cam_intrinsics[1, 2])
w2c_gt = np.loadtxt(os.path.join(REF_ROOT, "Jacob_test_result", "w2c_gt.txt"), dtype=np.float32)
T_noise = np.loadtxt(os.path.join(REF_ROOT, "Jacob_test_result", "T_noise.txt"), dtype=np.float32)
w2c = w2c_gt @ T_noise                                                                            # :1425
N = 15
rng = np.random.default_rng(15)
z = rng.uniform(0.8, 1.5, N)
xc = rng.uniform(-0.33, 0.33, N) * z
yc = rng.uniform(-0.24, 0.24, N) * z
xyz_cam_in = np.stack([xc, yc, z, np.ones(N)], 1)
xyz_world_np64 = (np.linalg.inv(w2c.astype(np.float64)) @ xyz_cam_in.T).T[:, :3]
s_lin = np.linalg.norm(w2c[:3, 0])                      # w2c_gt carries the object's scale (|column| = 0.46)
scales = np.exp(rng.normal(np.log(0.045), 0.35, (N, 3))) / s_lin
q = rng.normal(size=(N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
r, x, y, zq = q.T
Rq = np.stack([1 - 2 * (y * y + zq * zq), 2 * (x * y - r * zq), 2 * (x * zq + r * y),
               2 * (x * y + r * zq), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - r * x),
               2 * (x * zq - r * y), 2 * (y * zq + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(N, 3, 3)
L = Rq * scales[:, None, :]
Sig = L @ L.transpose(0, 2, 1)                           # build_covariance_from_scaling_rotation, gaussian_model.py:141-149
cov6 = np.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], 1).astype(np.float32)
opac = (1.0 / (1.0 + np.exp(-rng.normal(0.5, 1.0, (N, 1))))).astype(np.float32)
shs = np.zeros((N, 16, 3), np.float32)
shs[:, 0] = (rng.uniform(0.1, 0.9, (N, 3)) - 0.5) / 0.28209479177387814
shs[:, 1:] = rng.normal(0, 0.08, (N, 15, 3))
shs[3, 0, 1] = -2.5                                     # one clamped channel (colour + 0.5 < 0, forward.cu:65-72)
xyz_world = torch.from_numpy(xyz_world_np64.astype(np.float32))

with contextlib.redirect_stdout(io.StringIO()):
    w2c_ = torch.from_numpy(w2c).transpose(0, 1)                                                  # :1429-1431
    render_setting = ns["get_render_settings"](w, h, cam_intrinsics, w2c_)
projmatrix = render_setting["projmatrix"]
g = torch.Generator().manual_seed(4)
gt_img = torch.rand((3, h, w), generator=g)
gt_depth_t = torch.rand((1, h, w), generator=g) * 0.8 + 0.6
gt_depth_t[:, :, :40] = 0.0                                                                        # invalid depth
viewpoint = Camera(0, gt_img, gt_depth_t, torch.from_numpy(w2c_gt), projmatrix, fx, fy, cx, cy, 0.0, 0.0, h, w,
                   render_setting["viewmatrix"].T, device="cpu")                                  # :1452
mask_tensor = torch.zeros((h, w), dtype=torch.bool)
mask_tensor[30:450, 20:600] = True

# rendered images: the CPU oracle on exactly these inputs (the script reads them from the CUDA rasterizer, :1457-1470)
scene = dict(means3D=xyz_world.numpy(), opacities=opac, shs=shs, cov3D_precomp=cov6, image_height=h, image_width=w,
             tanfovx=float(render_setting["tanfovx"]), tanfovy=float(render_setting["tanfovy"]), bg=np.zeros(3, np.float32),
             scale_modifier=1.0, viewmatrix=render_setting["viewmatrix"].numpy().astype(np.float32),
             projmatrix=viewpoint.full_proj_transform.detach().numpy().astype(np.float32),
             projmatrix_raw=projmatrix.numpy().astype(np.float32), sh_degree=3,
             campos=viewpoint.camera_center.detach().numpy().astype(np.float32))
st = Oracle(np.float64).forward(scene)
render_image = torch.from_numpy(np.asarray(st["color"], np.float32).reshape(3, h, w))
render_depth = torch.from_numpy(np.asarray(st["depth"], np.float32).reshape(1, h, w))

# ------------------------------------------------------------------ the chain, :1474-1523 restated for the stubs, then :1524-1695 verbatim
xyz_world_homo = torch.cat([xyz_world, torch.ones(N, 1)], dim=1).numpy()
xyz_cam_homo = (w2c @ xyz_world_homo.T).T
gaussian_model = types.SimpleNamespace(get_features=torch.from_numpy(shs))
env = dict(ns)
env.update(xyz_world=xyz_world, xyz_world_homo=xyz_world_homo, xyz_cam=xyz_cam_homo[:, :3], gaussian_3D_covs=cov6, opacity=opac,
           gaussian_model=gaussian_model, viewpoint=viewpoint, render_setting=render_setting, fx=fx, fy=fy, cx=cx, cy=cy, w=w, h=h,
           image_size=(h, w), w2c=w2c, render_image=render_image, render_depth=render_depth, mask_tensor=mask_tensor)
cwd = os.getcwd()
with tempfile.TemporaryDirectory() as tmp:
    os.makedirs(os.path.join(tmp, "Jacob_test_result"))
    os.chdir(tmp)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(chain, REF, "exec"), env)
        saved = {f: np.load(os.path.join(tmp, "Jacob_test_result", f + ".npy"))
                 for f in ("grad_mu_I_pixel", "grad_Sigma_I_pixel", "grad_depth_per_gaussian", "dL_dtau")}
    finally:
        os.chdir(cwd)
for f, a in saved.items():          # same schema as the files the reference ships
    b = np.load(os.path.join(REF_ROOT, "Jacob_test_result", f + ".npy"))
    assert a.shape == b.shape and a.dtype == b.dtype, (f, a.shape, b.shape, a.dtype, b.dtype)

proj = env["image_projected_gaussians_sorted_by_depth"]
out = dict(
    w2c=w2c, intrinsics=cam_intrinsics, means3D=xyz_world.numpy(), cov3D=cov6, opacities=opac, shs=shs,
    viewmatrix=scene["viewmatrix"], projmatrix=scene["projmatrix"], projmatrix_raw=scene["projmatrix_raw"], campos=scene["campos"],
    tanfov=np.array([scene["tanfovx"], scene["tanfovy"]], np.float64),
    gt_color=gt_img.numpy(), gt_depth=gt_depth_t.numpy(), mask=mask_tensor.numpy(),
    rendered_color=render_image.numpy(), rendered_depth=render_depth.numpy(),
    # projection, :772-971 (depth-sorted order; `indices` maps a sorted position to the Gaussian)
    indices=np.asarray(env["indices"], np.int64),
    mean_2D=np.array([p["mean_2D"] for p in proj], np.float64), cov_2D=np.array([p["cov_2D"] for p in proj], np.float64),
    color=np.array([p["color"] for p in proj], np.float64), depth=np.array([p["depth"] for p in proj], np.float64),
    # dense gradients, :1173-1351 (the four Jacob_test_result/*.npy arrays, depth-sorted order) + per-Gaussian colour gradient
    grad_mu_I_pixel=saved["grad_mu_I_pixel"], grad_Sigma_I_pixel=saved["grad_Sigma_I_pixel"],
    grad_depth_per_gaussian=saved["grad_depth_per_gaussian"], grad_color_per_gaussian=np.asarray(env["grad_color_per_gaussian"]),
    # analytic Jacobians of every Gaussian, :633-760 (original order; rows scaled by the script to "pixel space")
    dmu_I_dT_all=np.asarray(env["dmu_I_dT_all"]), dcov_I_dT_all=np.asarray(env["dcov_I_dT_all"]),
    # chain rule, :1587-1695
    dL_dtau=saved["dL_dtau"], dL_dtau_mu=np.asarray(env["dL_dtau_mu_total"]), dL_dtau_cov=np.asarray(env["dL_dtau_cov_total"]),
    dL_dtau_depth=np.asarray(env["dL_dtau_depth_total"]), dL_dtau_sh=np.asarray(env["dL_dtau_sh_total"]),
)
np.savez_compressed(os.path.join(HERE, "script_chain_c0.npz"), **out)
print("wrote script_chain_c0.npz")
for k in ("dL_dtau", "dL_dtau_mu", "dL_dtau_cov", "dL_dtau_depth", "dL_dtau_sh"):
    print(k, out[k])
print("mean_2D", out["mean_2D"][:4], "\ncov_2D", out["cov_2D"][:2])
