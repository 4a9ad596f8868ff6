"""Drop-in replacement of the reference Python operator API
(submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:186-259):

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

Same names, argument meaning, outputs (color, radii, depth, opacity, n_touched), gradient order and
error behaviour; the body calls the B200-native CUDA library through its C-ABI (include/gsr_b200.h)
instead of the reference's pybind module.  The reference's debug prints / snapshot dumps are not
reproduced, the arithmetic is.
"""
import ctypes as C
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _cabi
from ._cabi import GsrScene

_L = _cabi.load()   # fails loudly when the CUDA library is absent: there is no CPU fallback


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t, name):
    """contiguous fp32 CUDA tensor with a 16-byte aligned base, or None for an empty tensor."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


class _Call:
    """Marshals one (settings, tensors) call into a gsr_scene; keeps the tensors alive."""

    def __init__(self, raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp):
        rs = raster_settings
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")   # rasterize_points.cu:73-75
        self.device = means3D.device
        self.P = int(means3D.shape[0])
        self.W, self.H = int(rs.image_width), int(rs.image_height)
        self.keep = dict(
            means3D=_f32c(means3D, "means3D"), shs=_f32c(sh, "shs"), colors_precomp=_f32c(colors_precomp, "colors_precomp"),
            opacities=_f32c(opacities, "opacities"), scales=_f32c(scales, "scales"), rotations=_f32c(rotations, "rotations"),
            cov3D_precomp=_f32c(cov3Ds_precomp, "cov3D_precomp"), background=_f32c(rs.bg, "bg"),
            viewmatrix=_f32c(rs.viewmatrix, "viewmatrix"), projmatrix=_f32c(rs.projmatrix, "projmatrix"),
            projmatrix_raw=_f32c(rs.projmatrix_raw, "projmatrix_raw"), campos=_f32c(rs.campos, "campos"))
        k = self.keep
        self.M = int(sh.shape[1]) if k["shs"] is not None else 0
        s = GsrScene()
        s.P, s.D, s.M, s.W, s.H = self.P, int(rs.sh_degree), self.M, self.W, self.H
        for f in ("background", "means3D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp",
                  "viewmatrix", "projmatrix", "projmatrix_raw", "campos"):
            setattr(s, f, k[f].data_ptr() if k[f] is not None else None)
        s.scale_modifier = float(rs.scale_modifier)
        s.tan_fovx, s.tan_fovy = float(rs.tanfovx), float(rs.tanfovy)
        s.prefiltered, s.debug = int(bool(rs.prefiltered)), int(bool(rs.debug))
        self.scene = s

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)


def _forward_impl(call, capacity=None):
    """Runs the forward.  capacity=None: exact sizing after the (single) host read of num_rendered,
    like the reference.  capacity=int: no host synchronisation, the caller checks the overflow flag."""
    dev, P, W, H = call.device, call.P, call.W, call.H
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    u8 = dict(dtype=torch.uint8, device=dev)
    color = torch.empty((3, H, W), **f32)
    depth = torch.empty((1, H, W), **f32)
    opacity = torch.empty((1, H, W), **f32)
    radii = torch.empty((P,), **i32)
    n_touched = torch.empty((P,), **i32)
    if P == 0:   # the reference returns zero-filled outputs without launching (rasterize_points.cu:84-100)
        color.zero_(); depth.zero_(); opacity.zero_()
        e = torch.empty((0,), **u8)
        return 0, 0, color, radii, e, e, e, depth, opacity, n_touched
    geom_bytes = _L.gsr_geometry_bytes(P, W, H)
    img_bytes = _L.gsr_image_bytes(W, H)
    geom = torch.empty((geom_bytes,), **u8)
    img = torch.empty((img_bytes,), **u8)
    st = call.stream()
    with torch.cuda.device(dev):
        _cabi.check(_L.gsr_forward_plan(C.byref(call.scene), _ptr(geom), geom_bytes, _ptr(radii), _ptr(n_touched), st), "forward_plan")
        if capacity is None:
            R, mt = C.c_longlong(0), C.c_longlong(0)
            _cabi.check(_L.gsr_forward_num_rendered(_ptr(geom), st, C.byref(R), C.byref(mt)), "forward_num_rendered")
            num_rendered = cap = int(R.value)
            max_tile = int(mt.value)
        else:
            num_rendered, cap, max_tile = -1, int(capacity), 0
        bin_bytes = _L.gsr_binning_bytes(P, W, H, cap)
        binning = torch.empty((bin_bytes,), **u8)
        _cabi.check(_L.gsr_forward_render(C.byref(call.scene), _ptr(geom), _ptr(binning), bin_bytes, cap, num_rendered, max_tile,
                                          _ptr(img), img_bytes, _ptr(color), _ptr(depth), _ptr(opacity), _ptr(n_touched), st),
                    "forward_render")
    return num_rendered, cap, color, radii, geom, binning, img, depth, opacity, n_touched


def _backward_impl(call, radii, geom, binning, cap, img, grad_color, grad_depth):
    dev, P, M = call.device, call.P, call.M
    f32 = dict(dtype=torch.float32, device=dev)
    k = call.keep
    g_means3D = torch.empty((P, 3), **f32)
    g_means2D = torch.empty((P, 3), **f32)
    g_opac = torch.empty((P, 1), **f32)
    g_sh = torch.empty((P, M, 3), **f32) if k["shs"] is not None else None
    g_col = torch.empty((P, 3), **f32) if k["colors_precomp"] is not None else None
    g_scales = torch.empty((P, 3), **f32) if k["scales"] is not None else None
    g_rot = torch.empty((P, 4), **f32) if k["scales"] is not None else None
    g_cov = torch.empty((P, 6), **f32) if k["cov3D_precomp"] is not None else None
    g_tau = torch.empty((6,), **f32)
    if P == 0:
        return g_means3D, g_means2D, g_sh, g_col, g_opac, g_scales, g_rot, g_cov, g_tau.zero_()
    gc = _f32c(grad_color, "grad_out_color")
    gd = _f32c(grad_depth, "grad_out_depth")
    with torch.cuda.device(dev):
        _cabi.check(_L.gsr_rasterize_gaussians_backward(
            C.byref(call.scene), _ptr(radii), _ptr(geom), _ptr(binning), cap, _ptr(img), _ptr(gc), _ptr(gd),
            _ptr(g_means3D), _ptr(g_means2D), _ptr(g_sh), _ptr(g_col), _ptr(g_opac), _ptr(g_scales), _ptr(g_rot),
            _ptr(g_cov), _ptr(g_tau), call.stream()), "backward")
    return g_means3D, g_means2D, g_sh, g_col, g_opac, g_scales, g_rot, g_cov, g_tau


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     theta, rho, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    # reference: __init__.py:48-184.  means2D, theta and rho are gradient sinks only (never read).
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                raster_settings):
        call = _Call(raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        num_rendered, cap, color, radii, geom, binning, img, depth, opacity, n_touched = _forward_impl(call)
        ctx.call = call
        ctx.cap = cap
        ctx.num_rendered = num_rendered
        ctx.shapes = (opacities.shape, means2D.shape, theta.shape, rho.shape, sh.numel(), colors_precomp.numel(),
                      scales.numel(), rotations.numel(), cov3Ds_precomp.numel())
        ctx.save_for_backward(radii, geom, binning, img)
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, opacity, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_radii, grad_out_depth, grad_out_opacity, grad_n_touched):
        call = ctx.call
        radii, geom, binning, img = ctx.saved_tensors
        op_shape, m2d_shape, th_shape, rho_shape, n_sh, n_col, n_sc, n_rot, n_cov = ctx.shapes
        dev, H, W = call.device, call.H, call.W
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        if grad_out_depth is None:
            grad_out_depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
        # grad_out_opacity is dropped, exactly like the reference (__init__.py:114,139-140)
        g_means3D, g_means2D, g_sh, g_col, g_opac, g_scales, g_rot, g_cov, g_tau = _backward_impl(
            call, radii, geom, binning, ctx.cap, img, grad_out_color, grad_out_depth)
        grad_rho = g_tau[:3].view(1, -1) if rho_shape.numel() else None
        grad_theta = g_tau[3:].view(1, -1) if th_shape.numel() else None
        if grad_rho is not None and len(rho_shape) == 1:
            grad_rho = grad_rho.view(rho_shape)
        if grad_theta is not None and len(th_shape) == 1:
            grad_theta = grad_theta.view(th_shape)
        grads = (
            g_means3D,
            g_means2D.view(m2d_shape) if tuple(m2d_shape) == (call.P, 3) else None,
            g_sh if n_sh else None,
            g_col if n_col else None,
            g_opac.view(op_shape),
            g_scales if n_sc else None,
            g_rot if n_rot else None,
            g_cov if n_cov else None,
            grad_theta,
            grad_rho,
            None,
        )
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (near-plane test of the camera) with a boolean (reference :206-215)
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions, "positions")
            P = int(positions.shape[0])
            present = torch.zeros((P,), dtype=torch.bool, device=positions.device)
            if P:
                vm, pm = _f32c(rs.viewmatrix, "viewmatrix"), _f32c(rs.projmatrix, "projmatrix")
                with torch.cuda.device(positions.device):
                    st = C.c_void_p(torch.cuda.current_stream(positions.device).cuda_stream)
                    _cabi.check(_L.gsr_mark_visible(P, _ptr(pos), _ptr(vm), _ptr(pm), _ptr(present), st), "mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if theta is None:
            theta = torch.Tensor([])
        if rho is None:
            rho = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   theta, rho, raster_settings)
