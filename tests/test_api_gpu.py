"""GPU tests of the drop-in Python operator API (GaussianRasterizationSettings / GaussianRasterizer):
autograd wiring, gradient order and shapes, error behaviour, markVisible, P == 0, and the C-ABI
one-call entry point with the allocator callback.  Numerical truth = the CPU oracle (float64)."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import rel_err, settings_from_scene

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _scene(P=2500, sh_degree=0, seed=4):
    import scenes as S

    cfg = dict(W=176, H=144, fx=160.0, fy=158.0, cx=88.0, cy=72.0, P=P, sh_degree=sh_degree)
    sc = S.make_scene(cfg, seed=seed)
    sc["scales"] = sc["scales"] * 2.0
    sc["bg"] = np.array([0.1, 0.3, 0.2], np.float32)
    return sc


def _oracle_grads(sc, dc, dd):
    from oracle.gs_oracle import Oracle

    o = Oracle(np.float64)
    st = o.forward(sc)
    return st, o.backward(st, dc, dd)


@pytest.mark.parametrize("sh_degree", [0, 2])
def test_autograd_end_to_end(sh_degree):
    from diff_gaussian_rasterization import GaussianRasterizer
    import scenes as S

    sc = _scene(sh_degree=sh_degree)
    t = S.to_torch(sc, "cuda")
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    theta = torch.zeros(3, device="cuda", requires_grad=True)       # cam_rot_delta, utils/camera_utils.py:50-55
    rho = torch.zeros(3, device="cuda", requires_grad=True)
    rast = GaussianRasterizer(settings_from_scene(t))
    color, radii, depth, opacity, n_touched = rast(
        means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=leaves["shs"],
        scales=leaves["scales"], rotations=leaves["rotations"], theta=theta, rho=rho)
    assert color.shape == (3, 144, 176) and depth.shape == (1, 144, 176) and opacity.shape == (1, 144, 176)
    assert radii.dtype == torch.int32 and n_touched.dtype == torch.int32 and not radii.requires_grad
    dc, dd = S.make_pixel_grads(176, 144, seed=3)
    gdc, gdd = torch.from_numpy(dc).cuda(), torch.from_numpy(dd).cuda()
    # the opacity image gets a gradient too: the operator must DROP it like the reference (B11)
    loss = (color * gdc).sum() + (depth * gdd).sum() + 7.0 * opacity.sum()
    loss.backward()
    st, g = _oracle_grads(sc, dc, dd)
    assert rel_err(color.detach().cpu().numpy(), st["color"]) <= TOL
    assert rel_err(depth.detach().cpu().numpy(), st["depth"]) <= TOL
    assert rel_err(opacity.detach().cpu().numpy(), st["opacity"]) <= TOL
    assert np.mean(radii.cpu().numpy() != st["radii"]) <= 2e-3
    for k, gk in (("means3D", "dL_dmeans3D"), ("shs", "dL_dsh"), ("opacities", "dL_dopacity"), ("scales", "dL_dscales"),
                  ("rotations", "dL_drotations")):
        a = leaves[k].grad.cpu().numpy()
        assert a.shape == tuple(leaves[k].shape)
        assert rel_err(a, np.asarray(g[gk]).reshape(a.shape)) <= TOL, k
    assert means2D.grad.shape == (2500, 3)
    assert rel_err(means2D.grad.cpu().numpy(), g["dL_dmean2D"]) <= TOL
    assert theta.grad.shape == (3,) and rho.grad.shape == (3,)
    tau = np.concatenate([rho.grad.cpu().numpy(), theta.grad.cpu().numpy()])      # tau = [rho, theta]
    assert rel_err(tau, g["dL_dtau"]) <= TOL


def test_backward_twice_and_determinism_of_integer_outputs():
    from diff_gaussian_rasterization import GaussianRasterizer
    import scenes as S

    sc = _scene(P=1500)
    t = S.to_torch(sc, "cuda")
    m = t["means3D"].clone().requires_grad_(True)
    rast = GaussianRasterizer(settings_from_scene(t))
    outs = []
    for _ in range(2):
        color, radii, depth, opacity, n_touched = rast(means3D=m, means2D=torch.zeros_like(m), opacities=t["opacities"],
                                                        shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        color.sum().backward(retain_graph=True)
        g1 = m.grad.clone()
        m.grad = None
        color.sum().backward()          # accumulators were cleared by the first backward: same answer
        g2 = m.grad.clone()
        m.grad = None
        # fp32 atomics: run-to-run summation order differs (reference B22), compare in max-norm
        assert rel_err(g1.cpu().numpy(), g2.cpu().numpy()) <= 1e-5
        assert float(g1.abs().max()) > 0
        outs.append((radii.clone(), n_touched.clone(), color.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2])      # forward is bit-reproducible


def test_mark_visible_and_empty_input():
    from diff_gaussian_rasterization import GaussianRasterizer
    import scenes as S
    from oracle.gs_oracle import Oracle

    sc = _scene(P=4000)
    t = S.to_torch(sc, "cuda")
    rast = GaussianRasterizer(settings_from_scene(t))
    vis = rast.markVisible(t["means3D"])
    assert vis.dtype == torch.bool
    np.testing.assert_array_equal(vis.cpu().numpy(), Oracle(np.float32).mark_visible(sc["means3D"], sc["viewmatrix"]))
    e = torch.zeros((0, 3), device="cuda")
    color, radii, depth, opacity, n_touched = rast(means3D=e, means2D=e, opacities=torch.zeros((0, 1), device="cuda"),
                                                    shs=torch.zeros((0, 1, 3), device="cuda"),
                                                    scales=e, rotations=torch.zeros((0, 4), device="cuda"))
    assert radii.numel() == 0 and float(color.abs().max()) == 0.0     # reference: zero-filled, no launch


def test_error_behaviour():
    from diff_gaussian_rasterization import GaussianRasterizer
    import scenes as S

    sc = _scene(P=100)
    t = S.to_torch(sc, "cuda")
    rast = GaussianRasterizer(settings_from_scene(t))
    with pytest.raises(Exception, match="excatly one of either SHs"):
        rast(means3D=t["means3D"], means2D=t["means3D"], opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair"):
        rast(means3D=t["means3D"], means2D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"])
    with pytest.raises(RuntimeError, match="num_points, 3"):
        rast(means3D=t["means3D"][:, :2], means2D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
             rotations=t["rotations"])


def test_cabi_one_call_with_allocator_callback():
    """gsr_rasterize_gaussians: the reference-shaped single entry point (allocator callback for the
    data-dependent binning buffer, returns num_rendered)."""
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import _cabi
    import scenes as S

    sc = _scene(P=3000)
    t = S.to_torch(sc, "cuda")
    e = torch.empty(0)
    call = dgr._Call(settings_from_scene(t), t["means3D"], t["shs"], e, t["opacities"], t["scales"], t["rotations"], e)
    L = _cabi.load()
    P, W, H = call.P, call.W, call.H
    f32 = dict(dtype=torch.float32, device="cuda")
    color, depth, opacity = torch.empty((3, H, W), **f32), torch.empty((1, H, W), **f32), torch.empty((1, H, W), **f32)
    radii = torch.empty(P, dtype=torch.int32, device="cuda")
    n_touched = torch.empty(P, dtype=torch.int32, device="cuda")
    gb, ib = L.gsr_geometry_bytes(P, W, H), L.gsr_image_bytes(W, H)
    geom = torch.empty(gb, dtype=torch.uint8, device="cuda")
    img = torch.empty(ib, dtype=torch.uint8, device="cuda")
    keep = []

    def alloc(user, nbytes):
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
        keep.append(buf)
        return buf.data_ptr()

    cb = _cabi.ALLOC_FN(alloc)
    binning, R = C.c_void_p(0), C.c_longlong(-1)
    p = lambda x: C.c_void_p(x.data_ptr())
    rc = L.gsr_rasterize_gaussians(C.byref(call.scene), p(geom), gb, p(img), ib, cb, None, C.byref(binning), C.byref(R),
                                   p(color), p(depth), p(opacity), p(radii), p(n_touched), call.stream())
    assert rc == 0, L.gsr_error_string()
    torch.cuda.synchronize()
    ref = dgr._forward_impl(call)
    assert R.value == ref[0] and binning.value == keep[0].data_ptr()
    assert torch.equal(color, ref[2]) and torch.equal(radii, ref[3]) and torch.equal(n_touched, ref[9])
