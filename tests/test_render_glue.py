"""render() glue (gaussian_renderer.py) as integration harness of the op.

CPU (build container, where /root/reference exists): the reference's own render()
(gaussian_splatting/gaussian_renderer/__init__.py:24-164) is parsed and every keyword it passes to
GaussianRasterizationSettings(...) and rasterizer(...) must be accepted by this package's classes, in the reference's field
order; our render() must have the reference's signature and result keys.
GPU: render() driven with GaussianModel-like / Camera-like stand-ins; autograd through the returned images must deliver the
gradients the C-ABI path returns (viewspace_points.grad = dL/dmean2D, cam_rot_delta.grad / cam_trans_delta.grad = dL/dtau)."""
import ast
import inspect
import os
import types

import numpy as np
import pytest
import torch

REF_RENDER = "/root/reference/gaussian_splatting/gaussian_renderer/__init__.py"


def _calls(tree, name):
    out = []
    for n in ast.walk(tree):
        if isinstance(n, ast.Call):
            f = n.func
            if (isinstance(f, ast.Name) and f.id == name) or (isinstance(f, ast.Attribute) and f.attr == name):
                out.append(n)
    return out


@pytest.mark.skipif(not os.path.exists(REF_RENDER), reason="reference tree not present (GPU box)")
def test_reference_render_call_sites_fit_this_package():
    import diff_gaussian_rasterization as dgr
    import gaussian_renderer as gr

    tree = ast.parse(open(REF_RENDER).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "render")
    # settings: same keywords, same order as the NamedTuple fields
    (call,) = _calls(fn, "GaussianRasterizationSettings")
    assert [k.arg for k in call.keywords] == list(dgr.GaussianRasterizationSettings._fields)
    # rasterizer(...) keywords of both call sites are parameters of GaussianRasterizer.forward
    params = list(inspect.signature(dgr.GaussianRasterizer.forward).parameters)[1:]
    sites = _calls(fn, "rasterizer")
    assert len(sites) == 2
    for c in sites:
        assert not c.args and all(k.arg in params for k in c.keywords)
    assert sorted(k.arg for k in sites[1].keywords) == sorted(params)    # the unmasked call names every parameter
    # our glue: the reference's signature (names + defaults) and result keys
    ref_args = [a.arg for a in fn.args.args]
    ours = inspect.signature(gr.render)
    assert list(ours.parameters) == ref_args
    assert [p.default for p in ours.parameters.values() if p.default is not inspect._empty] == [ast.literal_eval(d) for d in fn.args.defaults]
    ret = [n for n in ast.walk(fn) if isinstance(n, ast.Return) and isinstance(n.value, ast.Dict)][0]
    ref_keys = [k.value for k in ret.value.keys]
    src = inspect.getsource(gr.render)
    assert all('"%s"' % k in src for k in ref_keys)


def test_sh_to_rgb_matches_the_reference_eval_sh_golden():
    """pipe.convert_SHs_python: sh_to_rgb against colours computed by the reference's own eval_sh + render() expression
    (tests/golden/make_sh_golden.py imports gaussian_splatting/utils/sh_utils.py), every degree, float64; plus autograd."""
    import gaussian_renderer as gr

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sh_eval_golden.npz"))
    f, xyz, c = (torch.from_numpy(g[k]) for k in ("features", "xyz", "camera_center"))
    for deg in range(4):
        got = gr.sh_to_rgb(f, deg, xyz, c).numpy()
        assert got.shape == g["rgb_deg%d" % deg].shape
        np.testing.assert_allclose(got, g["rgb_deg%d" % deg], rtol=0, atol=1e-13)
        assert (got == 0).any() and (got > 0).any()      # the clamp at zero is exercised
    with pytest.raises(NotImplementedError):
        gr.sh_basis(4, xyz)
    f8, x8 = f[:8].clone().requires_grad_(True), xyz[:8].clone().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, b: gr.sh_to_rgb(a, 3, b, c), (f8, x8), eps=1e-6, atol=1e-6)


def _stand_ins(sc, device):
    import scenes as S

    t = S.to_torch(sc, device)
    leaf = lambda x: x.clone().requires_grad_(True)
    pc = types.SimpleNamespace(get_xyz=leaf(t["means3D"]), get_opacity=leaf(t["opacities"]), get_scaling=leaf(t["scales"]),
                               get_rotation=leaf(t["rotations"]), get_features=leaf(t["shs"]), active_sh_degree=sc["sh_degree"],
                               max_sh_degree=3)
    fov = lambda tan: 2.0 * np.arctan(tan)
    cam = types.SimpleNamespace(image_height=sc["image_height"], image_width=sc["image_width"], FoVx=fov(sc["tanfovx"]), FoVy=fov(sc["tanfovy"]),
                                world_view_transform=t["viewmatrix"], full_proj_transform=t["projmatrix"], projection_matrix=t["projmatrix_raw"],
                                camera_center=t["campos"], cam_rot_delta=torch.zeros(3, device=device, requires_grad=True),
                                cam_trans_delta=torch.zeros(3, device=device, requires_grad=True))
    return t, pc, cam


@pytest.mark.gpu
@pytest.mark.parametrize("masked", [False, True])
def test_render_glue_outputs_and_gradients(masked):
    import gaussian_renderer as gr
    import scenes as S
    from common import rel_err, run_ours

    cfg = dict(W=200, H=136, fx=180.0, fy=182.0, cx=100.0, cy=68.0, P=3000, sh_degree=2)
    sc = S.make_scene(cfg, seed=4)
    sc["scales"] = sc["scales"] * 2.0
    t, pc, cam = _stand_ins(sc, "cuda")
    mask = None
    if masked:
        mask = torch.rand(cfg["P"], generator=torch.Generator().manual_seed(1)).cuda() > 0.3
    pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False)
    pkg = gr.render(cam, pc, pipe, t["bg"], mask=mask)
    assert set(pkg) == {"render", "viewspace_points", "visibility_filter", "radii", "depth", "opacity", "n_touched"}
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"])
    loss = (pkg["render"] * torch.from_numpy(dc).cuda()).sum() + (pkg["depth"] * torch.from_numpy(dd).cuda()).sum()
    loss.backward()
    sub = sc if mask is None else dict(sc, **{k: sc[k][mask.cpu().numpy()] for k in ("means3D", "opacities", "scales", "rotations", "shs")})
    o = run_ours(sub, dc, dd)
    sel = slice(None) if mask is None else mask
    np.testing.assert_array_equal(pkg["radii"].cpu().numpy(), o["radii"])
    np.testing.assert_array_equal(pkg["n_touched"].cpu().numpy(), o["n_touched"])
    np.testing.assert_array_equal(pkg["visibility_filter"].cpu().numpy(), o["radii"] > 0)
    assert rel_err(pkg["render"].detach().cpu().numpy(), o["color"]) <= 1e-6
    assert rel_err(pkg["viewspace_points"].grad[sel].cpu().numpy(), o["dL_dmean2D"]) <= 1e-5
    assert rel_err(pc.get_xyz.grad[sel].cpu().numpy(), o["dL_dmeans3D"]) <= 1e-5
    assert rel_err(pc.get_features.grad[sel].cpu().numpy(), o["dL_dsh"]) <= 1e-5
    assert rel_err(pc.get_rotation.grad[sel].cpu().numpy(), o["dL_drotations"]) <= 1e-5
    assert rel_err(cam.cam_trans_delta.grad.cpu().numpy(), o["dL_dtau"][:3]) <= 1e-5        # rho = tau[:3]
    assert rel_err(cam.cam_rot_delta.grad.cpu().numpy(), o["dL_dtau"][3:]) <= 1e-5          # theta = tau[3:]
    if masked:
        assert float(pc.get_xyz.grad[~mask].abs().max()) == 0.0
    # an empty model renders nothing (reference :38-39)
    empty = types.SimpleNamespace(get_xyz=torch.zeros((0, 3), device="cuda"))
    assert gr.render(cam, empty, pipe, t["bg"]) is None


@pytest.mark.gpu
def test_render_glue_with_python_sh_conversion_matches_the_kernel_sh_path():
    """pipe.convert_SHs_python=True (colours from sh_to_rgb, handed over as colors_precomp) renders what the kernel's own SH
    evaluation renders; SH coefficients and positions receive the same gradients (autograd through sh_to_rgb against
    backward.cu:21-145).  The pose gradient differs on purpose: only the kernel path carries the colours' camera-centre term."""
    import gaussian_renderer as gr
    import scenes as S
    from common import rel_err

    cfg = dict(W=200, H=136, fx=180.0, fy=182.0, cx=100.0, cy=68.0, P=3000, sh_degree=3)
    sc = S.make_scene(cfg, seed=6)
    sc["scales"] = sc["scales"] * 2.0
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"])
    res = []
    for python_sh in (False, True):
        t, pc, cam = _stand_ins(sc, "cuda")
        pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=python_sh)
        pkg = gr.render(cam, pc, pipe, t["bg"])
        ((pkg["render"] * torch.from_numpy(dc).cuda()).sum() + (pkg["depth"] * torch.from_numpy(dd).cuda()).sum()).backward()
        res.append((pkg, pc))
    (a, pa), (b, pb) = res
    np.testing.assert_array_equal(a["radii"].cpu().numpy(), b["radii"].cpu().numpy())
    np.testing.assert_array_equal(a["n_touched"].cpu().numpy(), b["n_touched"].cpu().numpy())
    assert rel_err(b["render"].detach().cpu().numpy(), a["render"].detach().cpu().numpy()) <= 1e-5
    assert rel_err(pb.get_features.grad.cpu().numpy(), pa.get_features.grad.cpu().numpy()) <= 1e-4
    assert rel_err(pb.get_xyz.grad.cpu().numpy(), pa.get_xyz.grad.cpu().numpy()) <= 1e-4
    assert rel_err(pb.get_opacity.grad.cpu().numpy(), pa.get_opacity.grad.cpu().numpy()) <= 1e-4
