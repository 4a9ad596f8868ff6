"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/gsr_b200.h
declares, sizes are sane and argument errors mirror the reference (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gsr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsr_[a-z_0-9]+)\s*\(", src)) - {"gsr_alloc_fn"})


def test_library_exports_every_declared_symbol():
    from diff_gaussian_rasterization import _cabi

    lib = _cabi.load()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_cabi.EXPORTS) == names
    assert lib.gsr_version() >= 100


def test_workspace_sizes_monotone():
    from diff_gaussian_rasterization import _cabi

    lib = _cabi.load()
    g1, g2 = lib.gsr_geometry_bytes(1000, 640, 480), lib.gsr_geometry_bytes(100000, 640, 480)
    assert 0 < g1 < g2 and g2 >= 100000 * (48 + 64 + 4 + 1) + 1200 * 16
    assert lib.gsr_image_bytes(640, 480) >= 640 * 480 * 8
    b1, b2 = lib.gsr_binning_bytes(1000, 640, 480, 10000), lib.gsr_binning_bytes(1000, 640, 480, 1000000)
    assert b1 < b2 and b2 >= 1000000 * 12
    assert lib.gsr_geometry_bytes(0, 16, 16) > 0


def test_argument_errors_match_reference_messages():
    from diff_gaussian_rasterization import _cabi
    from diff_gaussian_rasterization._cabi import GsrScene

    lib = _cabi.load()
    s = GsrScene()
    s.P, s.W, s.H = 10, 64, 64
    fake = 0x1000  # never dereferenced: validation fails first
    for f in ("means3D", "opacities", "viewmatrix", "projmatrix", "background"):
        setattr(s, f, fake)
    rc = lib.gsr_forward_plan(C.byref(s), None, 0, None, None, None)
    assert rc == _cabi.GSR_ERR_ARG
    assert b"excatly one of either SHs or precomputed colors" in lib.gsr_error_string()
    s.colors_precomp = fake
    rc = lib.gsr_forward_plan(C.byref(s), None, 0, None, None, None)
    assert rc == _cabi.GSR_ERR_ARG
    assert b"scale/rotation pair or precomputed 3D covariance" in lib.gsr_error_string()
    s.cov3D_precomp = fake
    rc = lib.gsr_forward_plan(C.byref(s), None, 0, None, None, None)
    assert rc == _cabi.GSR_ERR_WORKSPACE
    with pytest.raises(Exception):
        _cabi.check(_cabi.GSR_ERR_ARG, "x")


def test_python_api_surface_matches_reference():
    import inspect

    import diff_gaussian_rasterization as d

    assert d.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "projmatrix_raw", "sh_degree", "campos", "prefiltered", "debug")
    sig = inspect.signature(d.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                    "rotations", "cov3D_precomp", "theta", "rho"]
    assert hasattr(d.GaussianRasterizer, "markVisible") and callable(d.rasterize_gaussians)
    r = d.GaussianRasterizer(None)
    import torch

    x = torch.zeros(1, 3)
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r.forward(x, x, x)
    with pytest.raises(Exception, match="exactly one of either scale/rotation"):
        r.forward(x, x, x, shs=x)


def test_argument_errors_of_the_round_2_entry_points():
    """Validation that returns before any CUDA work: bands of tile rows, SH degree / coefficient count, image limits, the
    backward's binning capacity, the switch reduction's rank / alignment / signal-pad checks."""
    from diff_gaussian_rasterization import _cabi
    from diff_gaussian_rasterization._cabi import GsrScene

    lib = _cabi.load()
    fake = 0x1000      # never dereferenced
    err = lambda: lib.gsr_error_string().decode()

    def scene(**kw):
        s = GsrScene()
        s.P, s.W, s.H, s.M, s.D = 10, 64, 48, 1, 0
        for f in ("means3D", "opacities", "viewmatrix", "projmatrix", "projmatrix_raw", "background", "campos", "shs", "scales", "rotations"):
            setattr(s, f, fake)
        for k, v in kw.items():
            setattr(s, k, v)
        return s

    plan = lambda s: lib.gsr_forward_plan(C.byref(s), None, 0, None, None, None)
    assert plan(scene()) == _cabi.GSR_ERR_WORKSPACE                                  # a valid scene gets as far as the workspace check
    assert plan(scene(tile_row_begin=2, tile_row_end=2)) == _cabi.GSR_ERR_ARG and "tile_row_begin" in err()
    assert plan(scene(tile_row_begin=0, tile_row_end=4)) == _cabi.GSR_ERR_ARG        # 48 rows of pixels = 3 tile rows
    assert plan(scene(tile_row_begin=1, tile_row_end=3)) == _cabi.GSR_ERR_WORKSPACE
    assert plan(scene(tile_row_begin=1, tile_row_end=3, densify_denom=fake)) == _cabi.GSR_ERR_ARG and "per view" in err()
    assert plan(scene(D=2, M=4)) == _cabi.GSR_ERR_ARG and "SH degree" in err()       # degree 2 needs 9 coefficients
    assert plan(scene(D=3, M=16)) == _cabi.GSR_ERR_WORKSPACE
    assert plan(scene(M=17)) == _cabi.GSR_ERR_ARG
    assert plan(scene(W=16 * 65535 + 1)) == _cabi.GSR_ERR_ARG and "too large" in err()
    assert plan(scene(W=0)) == _cabi.GSR_ERR_ARG and plan(scene(P=-1)) == _cabi.GSR_ERR_ARG
    assert plan(scene(rotations=fake + 4)) == _cabi.GSR_ERR_ARG and "16-byte" in err()
    vp = C.c_void_p
    bwd = lambda cap: lib.gsr_rasterize_gaussians_backward(C.byref(scene()), vp(fake), vp(fake), vp(fake), cap, vp(fake), vp(fake), vp(fake),
                                                           *([vp(fake)] * 9), None)
    assert bwd(-1) == _cabi.GSR_ERR_ARG and "capacity out of range" in err()
    assert bwd(1 << 31) == _cabi.GSR_ERR_ARG
    red = lambda mc, rank, world, n, pad: lib.gsr_window_allreduce(vp(mc), vp(fake), rank, world, n, 64, pad, vp(fake), None)
    assert red(0x10000, 0, 1, 1024, 4096) == _cabi.GSR_ERR_ARG and "world size" in err()
    assert red(0x10000, 8, 8, 1024, 4096) == _cabi.GSR_ERR_ARG
    assert red(0x10004, 0, 8, 1024, 4096) == _cabi.GSR_ERR_ARG and "16-byte aligned" in err()
    assert red(0x10000, 0, 8, 1022, 4096) == _cabi.GSR_ERR_ARG
    assert red(0x10000, 0, 8, 1024, 16) == _cabi.GSR_ERR_WORKSPACE and "signal pad" in err()
    assert lib.gsr_mark_visible(-1, None, None, None, None, None) == _cabi.GSR_ERR_ARG
    assert lib.gsr_sort_on_demand(-1) == lib.gsr_sort_on_demand(-1) > 0              # query leaves the threshold alone
