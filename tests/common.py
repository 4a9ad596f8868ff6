"""Shared helpers of the test-suite (test infrastructure)."""
import ctypes as C
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgsref.so")


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def dev_view(ptr, nbytes, dtype, device="cuda"):
    """torch view (no copy) of raw device memory."""
    if nbytes == 0:
        return torch.empty((0,), dtype=dtype, device=device)
    return torch.as_tensor(_Raw(ptr, nbytes), device=device).view(dtype)


def rel_err(a, b):
    """max-norm relative error |a-b|_inf / max(|b|_inf, tiny)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if a.size else 0.0


def l2_err(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)) if a.size else 0.0


def settings_from_scene(sc_t):
    from diff_gaussian_rasterization import GaussianRasterizationSettings

    return GaussianRasterizationSettings(
        image_height=sc_t["image_height"], image_width=sc_t["image_width"], tanfovx=sc_t["tanfovx"], tanfovy=sc_t["tanfovy"],
        bg=sc_t["bg"], scale_modifier=sc_t["scale_modifier"], viewmatrix=sc_t["viewmatrix"], projmatrix=sc_t["projmatrix"],
        projmatrix_raw=sc_t["projmatrix_raw"], sh_degree=sc_t["sh_degree"], campos=sc_t["campos"],
        prefiltered=False, debug=bool(sc_t.get("debug", False)))


def run_ours(sc, dL_dcolor=None, dL_ddepth=None, device="cuda", capacity=None, on_demand=0, exact_exp=0, lean=False, spatial_order=None, depth_cut=None):
    """Run the product path through its C-ABI on `device`; returns outputs, internals and gradients
    as numpy arrays.  sc: numpy scene dict (tests/scenes.py).
    on_demand: per-call threshold (gsr_scene.sort_on_demand) -- 0 (default here): every per-tile list is sorted completely,
    so that the WHOLE point_list can be compared; > 0: lists longer than this are ordered only as far as they are read;
    None: the library default (on demand, 256).
    exact_exp: gsr_scene.exact_exp (0 = library default = exact, -1 = ex2.approx).
    lean: skip the host copies of the per-Gaussian records (large scenes)."""
    return _run_ours(sc, dL_dcolor, dL_ddepth, device, capacity, on_demand, exact_exp, lean, spatial_order, depth_cut)


def _run_ours(sc, dL_dcolor, dL_ddepth, device, capacity, on_demand=0, exact_exp=0, lean=False, spatial_order=None, depth_cut=None):
    import diff_gaussian_rasterization as dgr
    import scenes as S

    t = S.to_torch(sc, device)
    rs = settings_from_scene(t)
    e = torch.empty(0)
    call = dgr._Call(rs, t["means3D"], t.get("shs", e) if t.get("colors_precomp") is None else e,
                     t.get("colors_precomp", e) if t.get("colors_precomp") is not None else e, t["opacities"],
                     t.get("scales", e) if t.get("cov3D_precomp") is None else e,
                     t.get("rotations", e) if t.get("cov3D_precomp") is None else e,
                     t.get("cov3D_precomp", e) if t.get("cov3D_precomp") is not None else e)
    call.scene.sort_on_demand = 0 if on_demand is None else (-1 if int(on_demand) == 0 else int(on_demand))
    call.scene.exact_exp = int(exact_exp)
    call.scene.spatial_order = None if spatial_order is None else spatial_order.data_ptr()      # device int32 permutation
    call.scene.depth_cut = None if depth_cut is None else depth_cut.data_ptr()                  # device int32[tiles], in/out
    R, cap, color, radii, geom, binning, img, depth, opacity, n_touched = dgr._forward_impl(call, capacity)
    torch.cuda.synchronize()
    P, W, H = call.P, call.W, call.H
    out = dict(color=color, radii=radii, depth=depth, opacity=opacity, n_touched=n_touched)
    ptrs = (C.c_ulonglong * 8)()
    dgr._L.gsr_debug_pointers(P, W, H, C.c_void_p(geom.data_ptr()), C.c_void_p(binning.data_ptr()), cap,
                              C.c_void_p(img.data_ptr()), ptrs)
    hdr = dev_view(ptrs[7], 16, torch.int32, device).cpu().numpy()
    R_dev = int(hdr[0])
    out["num_rendered"] = R_dev
    out["overflow"] = int(hdr[1])
    rec = dev_view(ptrs[0], P * 48, torch.float32, device).view(P, 12)
    vis = radii.cpu().numpy() > 0
    out["rec_dev"] = rec
    if not lean:
        rec = rec.cpu().numpy()
        out["means2D"] = rec[:, 0:2].copy()
        out["conic_opacity"] = np.stack([rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5]], 1)
        out["depths"] = rec[:, 6].copy()
        out["rgb"] = np.stack([rec[:, 7], rec[:, 8], rec[:, 9]], 1)
    out["tiles_touched"] = dev_view(ptrs[1], P * 4, torch.int32, device).cpu().numpy().astype(np.uint32)
    out["clamped_bits"] = dev_view(ptrs[2], P, torch.uint8, device).cpu().numpy()
    n_list = min(R_dev, cap)
    out["point_list"] = dev_view(ptrs[3], n_list * 4, torch.int32, device).cpu().numpy().astype(np.uint32)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = dev_view(ptrs[4], tiles * 8, torch.int32, device).view(tiles, 2).cpu().numpy().astype(np.uint32)
    out["final_T"] = dev_view(ptrs[5], W * H * 4, torch.float32, device).view(H, W).cpu().numpy()
    out["n_contrib"] = dev_view(ptrs[6], W * H * 4, torch.int32, device).view(H, W).cpu().numpy().astype(np.uint32)
    out["visible"] = vis
    for k in ("color", "radii", "depth", "opacity", "n_touched"):
        out[k] = out[k].cpu().numpy()
    if dL_dcolor is not None:
        gc = torch.from_numpy(np.ascontiguousarray(dL_dcolor)).to(device)
        gd = torch.from_numpy(np.ascontiguousarray(dL_ddepth)).to(device)
        g = dgr._backward_impl(call, radii, geom, binning, cap, img, gc, gd)
        torch.cuda.synchronize()
        names = ("dL_dmeans3D", "dL_dmean2D", "dL_dsh", "dL_dcolor", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dcov3D", "dL_dtau")
        for n, v in zip(names, g):
            out[n] = None if v is None else v.cpu().numpy()
    return out


class RefLib:
    """The UNMODIFIED reference kernels (oracle/_ref/libgsref.so) driven through the shim."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(REF_LIB)
        self.L = C.CDLL(REF_LIB)
        self.L.gsref_create.restype = C.c_void_p
        self.L.gsref_last_error.restype = C.c_char_p
        self.h = C.c_void_p(self.L.gsref_create())

    def close(self):
        if self.h:
            self.L.gsref_destroy(self.h)
            self.h = None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def forward(self, sc, device="cuda"):
        import scenes as S

        t = S.to_torch(sc, device)
        P = int(t["means3D"].shape[0])
        W, H = sc["image_width"], sc["image_height"]
        use_pre_col = sc.get("colors_precomp") is not None
        use_pre_cov = sc.get("cov3D_precomp") is not None
        shs = None if use_pre_col else t["shs"]
        M = 0 if shs is None else int(shs.shape[1])
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        o = dict(color=torch.zeros((3, H, W), **f32), depth=torch.zeros((1, H, W), **f32), opacity=torch.zeros((1, H, W), **f32),
                 radii=torch.zeros((P,), **i32), n_touched=torch.zeros((P,), **i32))
        self.t, self.P, self.M, self.W, self.H, self.shs = t, P, M, W, H, shs
        self.use_pre_col, self.use_pre_cov = use_pre_col, use_pre_cov
        p = self._p
        R = self.L.gsref_forward(
            self.h, P, int(sc["sh_degree"]), M, p(t["bg"]), W, H, p(t["means3D"]), p(shs),
            p(t["colors_precomp"]) if use_pre_col else None, p(t["opacities"]),
            None if use_pre_cov else p(t["scales"]), C.c_float(sc["scale_modifier"]),
            None if use_pre_cov else p(t["rotations"]), p(t["cov3D_precomp"]) if use_pre_cov else None,
            p(t["viewmatrix"]), p(t["projmatrix"]), p(t["campos"]), C.c_float(sc["tanfovx"]), C.c_float(sc["tanfovy"]), 0,
            p(o["color"]), p(o["depth"]), p(o["opacity"]), p(o["radii"]), p(o["n_touched"]), 0)
        torch.cuda.synchronize()
        if R < 0:
            raise RuntimeError("reference forward failed")
        self.R, self.radii = R, o["radii"]
        ptrs = (C.c_ulonglong * 13)()
        self.L.gsref_state_ptrs(self.h, ptrs)
        out = {k: v.cpu().numpy() for k, v in o.items()}
        out["num_rendered"] = R
        out["depths"] = dev_view(ptrs[0], P * 4, torch.float32, device).cpu().numpy()
        out["clamped"] = dev_view(ptrs[1], P * 3, torch.uint8, device).view(P, 3).cpu().numpy()
        out["means2D"] = dev_view(ptrs[2], P * 8, torch.float32, device).view(P, 2).cpu().numpy()
        out["cov3D"] = dev_view(ptrs[3], P * 24, torch.float32, device).view(P, 6).cpu().numpy()
        out["conic_opacity"] = dev_view(ptrs[4], P * 16, torch.float32, device).view(P, 4).cpu().numpy()
        out["rgb"] = dev_view(ptrs[5], P * 12, torch.float32, device).view(P, 3).cpu().numpy()
        out["tiles_touched"] = dev_view(ptrs[6], P * 4, torch.int32, device).cpu().numpy().astype(np.uint32)
        out["point_list"] = dev_view(ptrs[8], R * 4, torch.int32, device).cpu().numpy().astype(np.uint32)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        out["final_T"] = dev_view(ptrs[10], W * H * 4, torch.float32, device).view(H, W).cpu().numpy()
        out["n_contrib"] = dev_view(ptrs[11], W * H * 4, torch.int32, device).view(H, W).cpu().numpy().astype(np.uint32)
        out["ranges"] = dev_view(ptrs[12], tiles * 8, torch.int32, device).view(tiles, 2).cpu().numpy().astype(np.uint32)
        out["visible"] = out["radii"] > 0
        return out

    def backward(self, sc, dL_dcolor, dL_ddepth, device="cuda"):
        t, P, M, W, H = self.t, self.P, self.M, self.W, self.H
        f32 = dict(dtype=torch.float32, device=device)
        gc = torch.from_numpy(np.ascontiguousarray(dL_dcolor)).to(device)
        gd = torch.from_numpy(np.ascontiguousarray(dL_ddepth)).to(device)
        g = dict(dL_dmean2D=torch.zeros((P, 3), **f32), dL_dconic=torch.zeros((P, 2, 2), **f32), dL_dopacity=torch.zeros((P, 1), **f32),
                 dL_dcolor=torch.zeros((P, 3), **f32), dL_ddepth=torch.zeros((P, 1), **f32), dL_dmeans3D=torch.zeros((P, 3), **f32),
                 dL_dcov3D=torch.zeros((P, 6), **f32), dL_dsh=torch.zeros((P, max(M, 1), 3), **f32),
                 dL_dscales=torch.zeros((P, 3), **f32), dL_drotations=torch.zeros((P, 4), **f32), dL_dtau_pg=torch.zeros((P, 6), **f32))
        p = self._p
        rc = self.L.gsref_backward(
            self.h, P, int(sc["sh_degree"]), M, self.R, p(t["bg"]), W, H, p(t["means3D"]), p(self.shs),
            p(t["colors_precomp"]) if self.use_pre_col else None, None if self.use_pre_cov else p(t["scales"]),
            C.c_float(sc["scale_modifier"]), None if self.use_pre_cov else p(t["rotations"]),
            p(t["cov3D_precomp"]) if self.use_pre_cov else None, p(t["viewmatrix"]), p(t["projmatrix"]), p(t["projmatrix_raw"]),
            p(t["campos"]), C.c_float(sc["tanfovx"]), C.c_float(sc["tanfovy"]), p(self.radii), p(gc), p(gd),
            p(g["dL_dmean2D"]), p(g["dL_dconic"]), p(g["dL_dopacity"]), p(g["dL_dcolor"]), p(g["dL_ddepth"]), p(g["dL_dmeans3D"]),
            p(g["dL_dcov3D"]), p(g["dL_dsh"]), p(g["dL_dscales"]), p(g["dL_drotations"]), p(g["dL_dtau_pg"]), 0)
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError("reference backward failed")
        out = {k: v.cpu().numpy() for k, v in g.items()}
        out["dL_dtau"] = g["dL_dtau_pg"].sum(0).cpu().numpy()     # __init__.py:162-164
        if M == 0:
            out["dL_dsh"] = None
        return out
