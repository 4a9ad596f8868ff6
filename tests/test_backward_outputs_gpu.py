"""GPU tests of how the per-Gaussian backward delivers its outputs (csrc/preprocess_backward.cu): the [P,3] arrays leave
through shared-memory row slices as 16-byte stores when the buffers are 16-byte aligned and as scalar stores otherwise,
a partial last slice takes the tail path, and dL_dcolors can be asked for together with dL_dsh.  Everything through the
C-ABI (gsr_rasterize_gaussians_backward), compared bit for bit with the aligned call on the same forward state."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import settings_from_scene

pytestmark = pytest.mark.gpu


def _scene(P, sh_degree=0, seed=11):
    import scenes as S

    cfg = dict(W=176, H=144, fx=160.0, fy=158.0, cx=88.0, cy=72.0, P=P, sh_degree=sh_degree)
    sc = S.make_scene(cfg, seed=seed)
    sc["scales"] = sc["scales"] * 2.0
    return sc


def _forward(sc):
    import diff_gaussian_rasterization as dgr
    import scenes as S

    t = S.to_torch(sc, "cuda")
    e = torch.empty(0)
    call = dgr._Call(settings_from_scene(t), t["means3D"], t["shs"], e, t["opacities"], t["scales"], t["rotations"], e)
    R, cap, color, radii, geom, binning, img, depth, opacity, n_touched = dgr._forward_impl(call)
    return dgr, call, (radii, geom, binning, cap, img)


def _backward_raw(dgr, call, state, gc, gd, outs, accumulate=False):
    """outs: dict name -> tensor (or None); called straight through the C-ABI."""
    radii, geom, binning, cap, img = state
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    call.scene.accumulate_grads = 1 if accumulate else 0
    try:
        with torch.cuda.device(call.device):
            rc = dgr._L.gsr_rasterize_gaussians_backward(
                C.byref(call.scene), p(radii), p(geom), p(binning), cap, p(img), p(gc), p(gd), p(outs["means3D"]), p(outs["means2D"]),
                p(outs["sh"]), p(outs.get("colors")), p(outs["opacity"]), p(outs["scales"]), p(outs["rot"]), None, p(outs["tau"]),
                call.stream())
        assert rc == 0
        torch.cuda.synchronize()
    finally:
        call.scene.accumulate_grads = 0


def _alloc(P, M, offset_floats=0, fill=None, colors=False):
    """Output buffers whose data pointers sit `offset_floats` floats behind a 16-byte boundary."""
    def buf(*shape):
        n = int(np.prod(shape))
        base = torch.empty(n + 8, dtype=torch.float32, device="cuda")
        v = base[offset_floats:offset_floats + n].view(*shape)
        if fill is not None:
            v.fill_(fill)
        assert v.data_ptr() % 16 == (4 * offset_floats) % 16
        return v
    o = dict(means3D=buf(P, 3), means2D=buf(P, 3), sh=buf(P, M, 3), opacity=buf(P, 1), scales=buf(P, 3), rot=torch.empty((P, 4), device="cuda"),
             tau=torch.empty(6, device="cuda"))
    if fill is not None:
        o["rot"].fill_(fill)
    if colors:
        o["colors"] = buf(P, 3)
    return o


@pytest.mark.parametrize("P", [2501, 2560, 31])      # partial last slice (197 rows: 591 floats, 3-float tail), full slices, one warp
def test_misaligned_outputs_equal_aligned_outputs(P):
    import scenes as S

    sc = _scene(P)
    dc, dd = S.make_pixel_grads(176, 144)
    gc, gd = torch.from_numpy(dc).cuda(), torch.from_numpy(dd).cuda()
    res = []
    for off in (0, 1, 2):
        dgr, call, state = _forward(sc)       # the backward consumes the accumulators: one forward per backward
        o = _alloc(P, 1, off)
        _backward_raw(dgr, call, state, gc, gd, o)
        res.append({k: v.clone() for k, v in o.items()})
    assert float(res[0]["means3D"].abs().max()) > 0 and float(res[0]["sh"].abs().max()) > 0
    for r in res[1:]:
        for k in ("means2D", "opacity", "rot", "scales", "sh", "means3D"):
            # fp32 atomics in the compositing backward: run-to-run summation order differs, the staging itself is exact
            a, b = r[k].cpu().numpy(), res[0][k].cpu().numpy()
            assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-30), k
    # rows of culled Gaussians are written (zeros), whichever store path is taken
    radii = state[0].cpu().numpy()
    for r in res:
        assert np.all(r["means3D"].cpu().numpy()[radii <= 0] == 0) and np.all(r["sh"].cpu().numpy()[radii <= 0] == 0)


def test_accumulation_adds_rows_and_leaves_culled_rows_alone():
    import scenes as S

    P = 2501
    sc = _scene(P)
    dc, dd = S.make_pixel_grads(176, 144)
    gc, gd = torch.from_numpy(dc).cuda(), torch.from_numpy(dd).cuda()
    dgr, call, state = _forward(sc)
    plain = _alloc(P, 1, 0)
    _backward_raw(dgr, call, state, gc, gd, plain)
    for off in (0, 1):
        dgr, call, state = _forward(sc)
        acc = _alloc(P, 1, off, fill=0.25)
        _backward_raw(dgr, call, state, gc, gd, acc, accumulate=True)
        for k in ("means3D", "scales", "sh", "opacity", "rot"):
            a, b = acc[k].cpu().numpy() - 0.25, plain[k].cpu().numpy()
            assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-30) + 1e-7, k
        radii = state[0].cpu().numpy()
        assert np.all(acc["means3D"].cpu().numpy()[radii <= 0] == 0.25)


def test_colour_gradient_together_with_sh_gradient():
    import scenes as S

    P = 1800
    sc = _scene(P)
    dc, dd = S.make_pixel_grads(176, 144)
    gc, gd = torch.from_numpy(dc).cuda(), torch.from_numpy(dd).cuda()
    dgr, call, state = _forward(sc)
    o = _alloc(P, 1, 0, colors=True)
    _backward_raw(dgr, call, state, gc, gd, o)
    col, sh = o["colors"].cpu().numpy(), o["sh"].cpu().numpy()[:, 0, :]
    radii = state[0].cpu().numpy()
    assert np.all(col[radii <= 0] == 0)
    want = np.float32(0.28209479177387814) * col          # backward.cu:21-145 at degree 0, clamped channels -> 0
    live = sh != 0
    assert live.any() and np.array_equal(sh[live], want[live])
