"""TEST INFRASTRUCTURE (oracle) -- NOT product code.

ctypes front-end of oracle/libgs_oracle.so, the CPU restatement of the reference rasterizer
(see gs_oracle.c for the reference file:line of every stage).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

A *scene* is a plain dict of numpy float32 arrays / scalars with the fields of the reference's
GaussianRasterizationSettings (py/__init__.py:186-199) plus the per-Gaussian inputs of
GaussianRasterizer.forward (py/__init__.py:217):
  means3D[P,3] opacities[P,1] and (shs[P,M,3] | colors_precomp[P,3]) and
  ((scales[P,3], rotations[P,4]) | cov3D_precomp[P,6]);
  image_height image_width tanfovx tanfovy bg[3] scale_modifier viewmatrix[4,4] projmatrix[4,4]
  projmatrix_raw[4,4] sh_degree campos[3]
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/libgs_oracle.so (gcc, seconds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libgs_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgs_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "gs_oracle.c")):
            build()
        _LIB = C.CDLL(path)
        for sfx in ("_f32", "_f64"):
            getattr(_LIB, "gso_preprocess" + sfx).restype = C.c_longlong
            getattr(_LIB, "gso_bin" + sfx).restype = C.c_int
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Oracle:
    """Stage-by-stage CPU oracle in working precision ``dtype`` (np.float32 or np.float64)."""

    def __init__(self, dtype=np.float64):
        self.dt = np.dtype(dtype)
        assert self.dt in (np.dtype(np.float32), np.dtype(np.float64))
        self.sfx = "_f32" if self.dt == np.float32 else "_f64"
        self.L = lib()

    def _fn(self, name):
        return getattr(self.L, name + self.sfx)

    # ---- stage 1 -------------------------------------------------------------------------
    def preprocess(self, sc):
        P = int(sc["means3D"].shape[0])
        shs = _f32(sc.get("shs"))
        M = 0 if shs is None or shs.size == 0 else int(shs.shape[1])
        if M == 0:
            shs = None
        st = dict(
            P=P, M=M, D=int(sc["sh_degree"]), W=int(sc["image_width"]), H=int(sc["image_height"]),
            radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), self.dt), depths=np.zeros(P, self.dt),
            cov3D=np.zeros((P, 6), self.dt), rgb=np.zeros((P, 3), self.dt), conic_opacity=np.zeros((P, 4), self.dt),
            clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32),
        )
        a = {k: _f32(sc.get(k)) for k in ("means3D", "scales", "rotations", "opacities", "cov3D_precomp",
                                          "colors_precomp", "viewmatrix", "projmatrix", "projmatrix_raw", "campos", "bg")}
        for k in ("scales", "rotations", "cov3D_precomp", "colors_precomp"):
            if a[k] is not None and a[k].size == 0:
                a[k] = None
        st["_in"] = a
        st["_shs"] = shs
        st["_sc"] = sc
        R = self._fn("gso_preprocess")(
            C.c_int(P), C.c_int(st["D"]), C.c_int(M), _p(a["means3D"]), _p(a["scales"]),
            C.c_float(sc["scale_modifier"]), _p(a["rotations"]), _p(a["opacities"]), _p(shs), _p(a["cov3D_precomp"]),
            _p(a["colors_precomp"]), _p(a["viewmatrix"]), _p(a["projmatrix"]), _p(a["campos"]),
            C.c_int(st["W"]), C.c_int(st["H"]), C.c_float(sc["tanfovx"]), C.c_float(sc["tanfovy"]),
            _p(st["radii"]), _p(st["means2D"]), _p(st["depths"]), _p(st["cov3D"]), _p(st["rgb"]),
            _p(st["conic_opacity"]), _p(st["clamped"]), _p(st["tiles_touched"]))
        st["num_rendered"] = int(R)
        if a["cov3D_precomp"] is not None:
            st["cov3D"] = a["cov3D_precomp"].astype(self.dt)
        st["features"] = a["colors_precomp"].astype(self.dt) if a["colors_precomp"] is not None else st["rgb"]
        return st

    # ---- stage 2 -------------------------------------------------------------------------
    def bin(self, st, with_keys=False):
        W, H = st["W"], st["H"]
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        R = int(np.sum(st["tiles_touched"], dtype=np.int64))
        st["num_rendered"] = R
        st["point_list"] = np.zeros(max(R, 1), np.uint32)[:R]
        st["ranges"] = np.zeros((tiles, 2), np.uint32)
        keys = np.zeros(R, np.uint64) if with_keys else None
        rc = self._fn("gso_bin")(C.c_int(st["P"]), C.c_int(W), C.c_int(H), _p(st["radii"]),
                                 _p(np.ascontiguousarray(st["means2D"], self.dt)), _p(np.ascontiguousarray(st["depths"], self.dt)),
                                 C.c_longlong(R), _p(st["point_list"]) if R else None, _p(st["ranges"]), _p(keys))
        if rc != 0:
            raise RuntimeError("gso_bin failed rc=%d" % rc)
        if with_keys:
            st["keys"] = keys
        return st

    # ---- stage 3 -------------------------------------------------------------------------
    def render(self, st):
        W, H, P = st["W"], st["H"], st["P"]
        st["color"] = np.zeros((3, H, W), self.dt)
        st["depth"] = np.zeros((1, H, W), self.dt)
        st["opacity"] = np.zeros((1, H, W), self.dt)
        st["final_T"] = np.zeros((H, W), self.dt)
        st["n_contrib"] = np.zeros((H, W), np.uint32)
        st["n_touched"] = np.zeros(P, np.int32)
        self._fn("gso_render")(
            C.c_int(W), C.c_int(H), _p(st["ranges"]), _p(st["point_list"]), _p(np.ascontiguousarray(st["means2D"], self.dt)),
            _p(np.ascontiguousarray(st["features"], self.dt)), _p(np.ascontiguousarray(st["conic_opacity"], self.dt)),
            _p(np.ascontiguousarray(st["depths"], self.dt)), _p(st["_in"]["bg"]), _p(st["color"]), _p(st["depth"]),
            _p(st["opacity"]), _p(st["final_T"]), _p(st["n_contrib"]), _p(st["n_touched"]))
        return st

    def forward(self, sc):
        return self.render(self.bin(self.preprocess(sc)))

    # ---- stage 4 -------------------------------------------------------------------------
    def render_bwd(self, st, dL_dcolor, dL_ddepth):
        W, H, P = st["W"], st["H"], st["P"]
        g = dict(dL_dmean2D=np.zeros((P, 3), self.dt), dL_dconic=np.zeros((P, 2, 2), self.dt),
                 dL_dopacity=np.zeros((P, 1), self.dt), dL_dcolor=np.zeros((P, 3), self.dt),
                 dL_ddepth=np.zeros((P, 1), self.dt))
        dpc = np.ascontiguousarray(np.asarray(dL_dcolor).reshape(3, H, W), self.dt)
        dpd = np.ascontiguousarray(np.asarray(dL_ddepth).reshape(H, W), self.dt)
        self._fn("gso_render_bwd")(
            C.c_int(W), C.c_int(H), _p(st["ranges"]), _p(st["point_list"]), _p(np.ascontiguousarray(st["means2D"], self.dt)),
            _p(np.ascontiguousarray(st["features"], self.dt)), _p(np.ascontiguousarray(st["conic_opacity"], self.dt)),
            _p(np.ascontiguousarray(st["depths"], self.dt)), _p(st["_in"]["bg"]),
            _p(np.ascontiguousarray(st["final_T"], self.dt)), _p(st["n_contrib"]), _p(dpc), _p(dpd),
            _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolor"]), _p(g["dL_ddepth"]))
        return g

    # ---- stage 5 -------------------------------------------------------------------------
    def preprocess_bwd(self, st, g):
        P, M = st["P"], st["M"]
        a, sc = st["_in"], st["_sc"]
        o = dict(dL_dmeans3D=np.zeros((P, 3), self.dt), dL_dcov3D=np.zeros((P, 6), self.dt),
                 dL_dsh=np.zeros((P, M, 3), self.dt), dL_dscales=np.zeros((P, 3), self.dt),
                 dL_drotations=np.zeros((P, 4), self.dt), dL_dtau_per_gaussian=np.zeros((P, 6), self.dt))
        shs = st["_shs"] if a["colors_precomp"] is None else None
        self._fn("gso_preprocess_bwd")(
            C.c_int(P), C.c_int(st["D"]), C.c_int(M), _p(a["means3D"]), _p(st["radii"]), _p(shs), _p(st["clamped"]),
            _p(a["scales"]), _p(a["rotations"]), C.c_float(sc["scale_modifier"]),
            _p(np.ascontiguousarray(st["cov3D"], self.dt)), _p(a["viewmatrix"]), _p(a["projmatrix"]),
            _p(a["projmatrix_raw"]), _p(a["campos"]), C.c_int(st["W"]), C.c_int(st["H"]),
            C.c_float(sc["tanfovx"]), C.c_float(sc["tanfovy"]),
            _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolor"]), _p(g["dL_ddepth"]),
            _p(o["dL_dmeans3D"]), _p(o["dL_dcov3D"]), _p(o["dL_dsh"]), _p(o["dL_dscales"]), _p(o["dL_drotations"]),
            _p(o["dL_dtau_per_gaussian"]))
        o["dL_dtau"] = o["dL_dtau_per_gaussian"].sum(axis=0)  # py/__init__.py:162-164
        o.update(g)
        return o

    def backward(self, st, dL_dcolor, dL_ddepth):
        return self.preprocess_bwd(st, self.render_bwd(st, dL_dcolor, dL_ddepth))

    # ---- KAT hooks -----------------------------------------------------------------------
    def cov2d_pose_jacobian(self, mean, fx, fy, tanx, tany, cov3D6, viewmatrix):
        jac = np.zeros((3, 6), np.float64)
        self._fn("gso_cov2d_pose_jacobian")(_p(np.ascontiguousarray(mean, np.float64)), C.c_double(fx), C.c_double(fy),
                                            C.c_double(tanx), C.c_double(tany), _p(np.ascontiguousarray(cov3D6, np.float64)),
                                            _p(_f32(viewmatrix)), _p(jac))
        return jac

    def mean2d_pose_jacobian(self, mean, viewmatrix, projmatrix, projmatrix_raw):
        jac = np.zeros((2, 6), np.float64)
        self._fn("gso_mean2d_pose_jacobian")(_p(np.ascontiguousarray(mean, np.float64)), _p(_f32(viewmatrix)),
                                             _p(_f32(projmatrix)), _p(_f32(projmatrix_raw)), _p(jac))
        return jac

    def mark_visible(self, means3D, viewmatrix):
        m = _f32(means3D)
        out = np.zeros(m.shape[0], np.uint8)
        self._fn("gso_mark_visible")(C.c_int(m.shape[0]), _p(m), _p(_f32(viewmatrix)), _p(out))
        return out.astype(bool)
