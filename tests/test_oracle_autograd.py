"""CPU oracle (analytic gradients restating the reference kernels) vs torch.autograd through an
independent dense float64 re-statement of the forward with the pose entering as Exp(tau)*T_cw
(the check VerifyJacobian.ipynb does for one Gaussian, here for whole small scenes and every input)."""
import numpy as np
import pytest
import torch

from autograd_ref import render_autograd
import scenes as S
from oracle.gs_oracle import Oracle


def _setup(P, seed, W=96, H=64, bg=(0.1, 0.2, 0.3), scale_mul=3.0):
    # centred principal point: the reference's mean2D pose Jacobian omits the principal-point terms (B17)
    cfg = dict(W=W, H=H, fx=90.0, fy=85.0, cx=W / 2, cy=H / 2, P=P, sh_degree=0)
    sc = S.make_scene(cfg, seed=seed)
    sc["scales"] = sc["scales"] * scale_mul
    sc["bg"] = np.asarray(bg, np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    dc = np.stack([np.sin(xx / 40.0) + 0.3, np.cos(yy / 35.0), 0.5 + 0 * xx]).astype(np.float32)
    dd = (0.2 + np.sin((xx + yy) / 50.0))[None].astype(np.float32)
    Pm = S.projection_matrix2(0.01, 100.0, cfg["cx"], cfg["cy"], cfg["fx"], cfg["fy"], W, H).astype(np.float64)
    return cfg, sc, dc, dd, Pm


@pytest.mark.parametrize("P,seed", [(120, 3), (200, 11)])
def test_all_gradients_match_autograd(P, seed):
    cfg, sc, dc, dd, Pm = _setup(P, seed)
    o = Oracle(np.float64)
    st = o.forward(sc)
    g = o.backward(st, dc, dd)
    w2c = sc["viewmatrix"].T.astype(np.float64)
    tau = torch.zeros(6, dtype=torch.float64, requires_grad=True)
    params = {k: torch.tensor(sc[k], dtype=torch.float64, requires_grad=True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    c, d = render_autograd(sc, w2c, Pm, st, params, tau)
    assert np.abs(c.detach().numpy() - st["color"]).max() < 5e-6
    assert np.abs(d.detach().numpy() - st["depth"][0]).max() < 5e-5
    L = (c * torch.tensor(dc, dtype=torch.float64)).sum() + (d * torch.tensor(dd[0], dtype=torch.float64)).sum()
    L.backward()
    ref_tau = tau.grad.numpy()
    assert np.abs(ref_tau - g["dL_dtau"]).max() / np.abs(ref_tau).max() < 1e-5
    for k, gk in (("means3D", "dL_dmeans3D"), ("scales", "dL_dscales"), ("rotations", "dL_drotations"),
                  ("opacities", "dL_dopacity"), ("shs", "dL_dsh")):
        a = params[k].grad.numpy()
        b = np.asarray(g[gk]).reshape(a.shape)
        assert np.abs(a - b).max() / max(np.abs(a).max(), 1e-30) < 1e-5, k


def test_f32_oracle_tracks_f64():
    cfg, sc, dc, dd, Pm = _setup(150, 5)
    g64 = Oracle(np.float64)
    g32 = Oracle(np.float32)
    s64, s32 = g64.forward(sc), g32.forward(sc)
    assert np.mean(s64["radii"] != s32["radii"]) < 0.02
    assert np.abs(s64["color"] - s32["color"]).max() < 2e-4
