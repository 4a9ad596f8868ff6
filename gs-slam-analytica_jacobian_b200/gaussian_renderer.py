"""render(): the per-view glue between a Gaussian model + camera and the rasterizer op -- the call every consumer of the
reference makes (gaussian_splatting/gaussian_renderer/__init__.py:24-164; callers: utils/slam_frontend.py:164-166,
utils/slam_backend.py:87-89,173-176,204-206, gui, eval).  Same signature, same argument meaning, same result dictionary, so
the application code keeps working when `gaussian_splatting.gaussian_renderer.render` is pointed here.

It is glue only: it reads the model's activated tensors and the camera's matrices, builds GaussianRasterizationSettings and
calls GaussianRasterizer.  Differences from the reference, all deliberate:
  * no debug printing of the matrices on every call (:70-77);
  * the `mask` branch returns all five outputs (the reference unpacks four of the five and then fails on `n_touched`,
    :125-138,156-164) and accepts models without SHs / scales in the masked call;
  * `override_color` is honoured (precomputed colours, as in upstream 3DGS); in the reference it is dead code: `colors_precomp` is
    set to None right in front of `if colors_precomp is None` (:106-121), so the argument is ignored there;
  * `pipe.convert_SHs_python` (SH -> RGB in torch, :108-117) is evaluated as one [P, M] basis matrix times the coefficients
    (`sh_to_rgb`; pinned to the reference's eval_sh, tests/test_render_glue.py).  The kernel path (default) computes the same
    colours and, unlike the torch path, also carries their camera-centre dependence into dL/dtau (backward.cu:139-143).
SLAM loops that call this V times per iteration should use engine.RasterEngine / window.KeyframeWindow instead (no per-call
allocations, no host synchronisation, CUDA graphs); this function is the compatible path, not the fast one.
"""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


# real spherical harmonics up to degree 3 in the 3DGS sign convention (gaussian_splatting/utils/sh_utils.py:24-52 holds the same constants)
_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
          -0.5900435899266435)


def sh_basis(degree, dirs):
    """[P, (degree + 1)^2] values of the SH basis functions at the unit directions dirs [P, 3]."""
    if not 0 <= degree <= 3:
        raise NotImplementedError("SH degree outside 0..3 (the rasterizer's range, forward.cu:22-76)")
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    cols = [torch.full_like(x, _SH_C0)]
    if degree > 0:
        cols += [-_SH_C1 * y, _SH_C1 * z, -_SH_C1 * x]
    if degree > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [_SH_C2[0] * xy, _SH_C2[1] * yz, _SH_C2[2] * (2.0 * zz - xx - yy), _SH_C2[3] * xz, _SH_C2[4] * (xx - yy)]
    if degree > 2:
        cols += [_SH_C3[0] * y * (3 * xx - yy), _SH_C3[1] * xy * z, _SH_C3[2] * y * (4 * zz - xx - yy),
                 _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _SH_C3[4] * x * (4 * zz - xx - yy), _SH_C3[5] * z * (xx - yy),
                 _SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(cols, dim=1)


def sh_to_rgb(features, degree, xyz, camera_center):
    """pipe.convert_SHs_python (reference :108-117): colours [P, 3] = max(0, SH(features [P, M, 3], direction camera -> Gaussian) + 0.5)."""
    d = xyz - camera_center.reshape(1, 3)
    basis = sh_basis(int(degree), d / d.norm(dim=1, keepdim=True))
    n = basis.shape[1]
    return torch.clamp_min((features[:, :n, :] * basis[:, :, None]).sum(dim=1) + 0.5, 0.0)


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, mask=None):
    """Returns None for an empty model, else {"render", "viewspace_points", "visibility_filter", "radii", "depth",
    "opacity", "n_touched"}.  `viewspace_points` is the zero tensor whose .grad receives dL/dmean2D (densification)."""
    xyz = pc.get_xyz
    if xyz.shape[0] == 0:
        return None
    screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0      # non-leaf with retained grad, like the reference
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        projmatrix_raw=viewpoint_camera.projection_matrix, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer(raster_settings=settings)

    scales = rotations = cov3D_precomp = None
    if getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales = pc.get_scaling
        if scales.shape[-1] == 1:            # isotropic models keep one scale per Gaussian
            scales = scales.repeat(1, 3)
        rotations = pc.get_rotation
    shs = colors_precomp = None
    if override_color is not None:
        colors_precomp = override_color
    elif getattr(pipe, "convert_SHs_python", False):
        colors_precomp = sh_to_rgb(pc.get_features, pc.active_sh_degree, xyz, viewpoint_camera.camera_center)
    else:
        shs = pc.get_features

    pick = (lambda t: t) if mask is None else (lambda t: None if t is None else t[mask])
    rendered_image, radii, depth, opacity, n_touched = rasterizer(
        means3D=pick(xyz), means2D=pick(screenspace_points), shs=pick(shs), colors_precomp=pick(colors_precomp),
        opacities=pick(pc.get_opacity), scales=pick(scales), rotations=pick(rotations), cov3D_precomp=pick(cov3D_precomp),
        theta=viewpoint_camera.cam_rot_delta, rho=viewpoint_camera.cam_trans_delta)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "opacity": opacity, "n_touched": n_touched}
