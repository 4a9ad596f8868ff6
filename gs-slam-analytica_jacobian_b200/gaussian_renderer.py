"""render(): the per-view glue between a Gaussian model + camera and the rasterizer op -- the call every consumer of the
reference makes (gaussian_splatting/gaussian_renderer/__init__.py:24-164; callers: utils/slam_frontend.py:164-166,
utils/slam_backend.py:87-89,173-176,204-206, gui, eval).  Same signature, same argument meaning, same result dictionary, so
the application code keeps working when `gaussian_splatting.gaussian_renderer.render` is pointed here.

It is glue only: it reads the model's activated tensors and the camera's matrices, builds GaussianRasterizationSettings and
calls GaussianRasterizer.  Differences from the reference, all deliberate:
  * no debug printing of the matrices on every call (:70-77);
  * the `mask` branch returns all five outputs (the reference unpacks four of the five and then fails on `n_touched`,
    :125-138,156-164) and accepts models without SHs / scales in the masked call;
  * `pipe.convert_SHs_python` (SH -> RGB in torch instead of in the kernel) is not offered: the kernel path computes the
    same colours and, unlike the torch path, carries the camera-centre dependence into dL/dtau (backward.cu:139-143).
SLAM loops that call this V times per iteration should use engine.RasterEngine / window.KeyframeWindow instead (no per-call
allocations, no host synchronisation, CUDA graphs); this function is the compatible path, not the fast one.
"""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


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
        raise NotImplementedError("convert_SHs_python: use the kernel's SH evaluation (it also carries the pose gradient)")
    else:
        shs = pc.get_features

    pick = (lambda t: t) if mask is None else (lambda t: None if t is None else t[mask])
    rendered_image, radii, depth, opacity, n_touched = rasterizer(
        means3D=pick(xyz), means2D=pick(screenspace_points), shs=pick(shs), colors_precomp=pick(colors_precomp),
        opacities=pick(pc.get_opacity), scales=pick(scales), rotations=pick(rotations), cov3D_precomp=pick(cov3D_precomp),
        theta=viewpoint_camera.cam_rot_delta, rho=viewpoint_camera.cam_trans_delta)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "opacity": opacity, "n_touched": n_touched}
