"""The numpy restatement of the reference's CPU script (oracle/loss_derivative_2d.py) against the golden vectors the
reference module itself produced (tests/golden/loss_derivative_2d_kat.json)."""
import json
import os

import numpy as np

from oracle import loss_derivative_2d as LD

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    d = json.load(open(os.path.join(HERE, "golden", "loss_derivative_2d_kat.json")))
    gs = [dict(mu_I=np.array(g["mu_I"]), Sigma_I=np.array(g["Sigma_I"]), opacity=g["opacity"], color=np.array(g["color"]),
               depth=g["depth"]) for g in d["gaussians"]]
    return d, gs


def test_gradients_match_reference_script():
    d, gs = _load()
    gm, gS = LD.compute_gradients_2D(gs, np.array(d["rendered_color"]), np.array(d["rendered_depth"]), np.array(d["gt_color"]),
                                     np.array(d["gt_depth"]), (d["H"], d["W"]))
    np.testing.assert_allclose(np.asarray(gm), np.asarray(d["grad_mu_I"]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.asarray(gS), np.asarray(d["grad_Sigma_I"]), rtol=1e-10, atol=1e-12)


def test_alpha_matches_reference_script():
    d, gs = _load()
    H, W = d["H"], d["W"]
    mu = np.array([g["mu_I"] for g in gs]); S = np.array([g["Sigma_I"] for g in gs]); o = np.array([g["opacity"] for g in gs])
    px = np.array([[u, v] for v in (0, H - 1) for u in (0, 5, W - 1)], np.float64)
    a, _, _ = LD.alpha_at_pixels(mu, S, o, px)                # [6, N]
    got = np.array([[a[vi * 3 + ui, gi] for ui in range(3)] for gi in range(len(gs)) for vi in range(2)])
    np.testing.assert_allclose(got, np.asarray(d["alpha_samples"]), rtol=1e-12, atol=0)
