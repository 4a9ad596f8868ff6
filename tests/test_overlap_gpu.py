"""Kernels that start inside their predecessor's tail: the compositing backward launched as a programmatic
dependent of the compositing forward with per-tile release/acquire flags (gsr_scene.overlap_forward), the upstream-gradient
word (gsr_scene.upstream_ready) and the per-Gaussian backward behind the compositing backward.  Ordering bugs here would be
timing dependent, so the same steps are repeated many times, eagerly and as CUDA graphs, and every result is compared
with the plainly ordered path: forward products bit-identical, gradients within fp32 summation noise."""
import numpy as np
import pytest
import torch

from common import rel_err

pytestmark = pytest.mark.gpu


def _engine(P=60000, W=320, H=240, seed=0, overlap=True):
    import scenes as S
    from diff_gaussian_rasterization.engine import RasterEngine

    cfg = dict(S.CONFIGS["C1_tum_tracking"], W=W, H=H, P=P, fx=260.0, fy=260.0, cx=W / 2 - 0.5, cy=H / 2 - 0.5)
    sc = S.make_scene(cfg, seed=seed)
    t = S.to_torch(sc, "cuda")
    eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                       W, H, sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)
    eng.overlap = overlap
    poses = S.noisy_poses(12, seed=5)
    cams = []
    for w2c in poses:
        cam = S.make_camera(W, H, cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], w2c)
        cams.append(eng.pack_camera(*(torch.from_numpy(cam[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda())
    dc, dd = S.make_pixel_grads(W, H, seed=3)
    eng.dL_dcolor.copy_(torch.from_numpy(dc))
    eng.dL_ddepth.copy_(torch.from_numpy(dd))
    for c in cams:
        eng.set_camera(c)
        eng.calibrate()
    return eng, cams, (dc, dd)


def _snapshot(eng):
    torch.cuda.synchronize()
    return dict(color=eng.color.cpu().numpy().copy(), depth=eng.depth.cpu().numpy().copy(), n_touched=eng.n_touched.cpu().numpy().copy(),
                tau=eng.g_tau.cpu().numpy().copy(), means3D=eng.g_means3D.cpu().numpy().copy(), opacity=eng.g_opacity.cpu().numpy().copy(),
                rot=eng.g_rot.cpu().numpy().copy())


def _same(a, b):
    for k in ("color", "depth", "n_touched"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("tau", "means3D", "opacity", "rot"):
        assert rel_err(a[k], b[k]) <= 2e-5, k


@pytest.mark.parametrize("use_graph", [False, True])
def test_overlapped_step_matches_plain_order(use_graph):
    plain, cams, _ = _engine(overlap=False)
    ref = []
    for c in cams:
        plain.set_camera(c)
        plain.step(use_graph=False)
        ref.append(_snapshot(plain))
    eng, cams, _ = _engine(overlap=True)
    for rep in range(6):                     # back to back, no host sync between steps inside a repetition
        for i, c in enumerate(cams):
            eng.set_camera(c)
            eng.step(use_graph=use_graph)
            if (i + rep) % 3 == 0:           # sample some steps, let the others chase each other on the stream
                _same(_snapshot(eng), ref[i])
        assert eng.header()[1] is False


def test_host_step_with_upstream_word_matches():
    plain, cams, (dc, dd) = _engine(overlap=False)
    ref = []
    for c in cams:
        plain.set_camera(c)
        plain.step(use_graph=False)
        ref.append(_snapshot(plain))
    eng, cams, (dc, dd) = _engine(overlap=True)
    cam_pin = torch.empty(52, dtype=torch.float32).pin_memory()
    dc_pin, dd_pin = torch.from_numpy(dc).pin_memory(), torch.from_numpy(dd).pin_memory()
    eng.dL_dcolor.zero_()                    # the graph has to bring the gradients in itself
    eng.dL_ddepth.zero_()
    cam_pin.copy_(cams[0].cpu())
    eng.capture_host_step(cam_pin, dc_pin, dd_pin)
    assert getattr(eng, "up_flag", None) is not None
    for rep in range(5):
        for i, c in enumerate(cams):
            cam_pin.copy_(c.cpu())
            tau, hdr = eng.step_host()
            assert int(hdr[1]) == 0
            assert rel_err(tau.numpy(), ref[i]["tau"]) <= 2e-5
            _same(_snapshot(eng), ref[i])


def test_window_views_overlap_their_own_forward():
    from diff_gaussian_rasterization.window import KeyframeWindow

    plain, cams, _ = _engine(overlap=False)
    eng, _, _ = _engine(overlap=True)
    packed = torch.stack(cams[:5])
    gc = torch.stack([plain.dL_dcolor] * 5)
    gd = torch.stack([plain.dL_ddepth] * 5)
    w0 = KeyframeWindow(plain, packed)
    w1 = KeyframeWindow(eng, packed)
    f0 = w0.iteration((gc, gd)).clone()
    for _ in range(4):
        f1 = w1.iteration((gc, gd), upstream_precomputed=True)
        torch.cuda.synchronize()
        assert rel_err(f1.cpu().numpy(), f0.cpu().numpy()) <= 2e-5
        assert rel_err(w1.tau.cpu().numpy(), w0.tau.cpu().numpy()) <= 2e-5


_CHAIN_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "gs-slam-analytica_jacobian_b200")); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch
from test_overlap_gpu import _engine, _snapshot
eng, cams, _ = _engine(P=40000, overlap=True)
out = {}
for rep in range(3):
    for i, c in enumerate(cams[:6]):
        eng.set_camera(c)
        eng.step(use_graph=bool(rep))          # eager first, then graph replays chasing each other
        if rep == 2:
            for k, v in _snapshot(eng).items():
                out["%s_%d" % (k, i)] = v
assert eng.header()[1] is False
np.savez(sys.argv[2], **out)
"""


def test_programmatic_launch_chain_matches_serialised_launches(tmp_path):
    """The whole chain of programmatic dependent launches (preprocess -> compositing forward -> compositing backward ->
    per-Gaussian backward with its arithmetic in front of the dependency wait) against the same library with every launch
    fully serialised (GSR_NO_PDL=1, read when the library is loaded: two fresh processes)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for name, env_extra in (("chain", {}), ("serial", {"GSR_NO_PDL": "1"})):
        env = dict(os.environ)
        env.pop("GSR_NO_PDL", None)
        env.update(env_extra)
        path = str(tmp_path / (name + ".npz"))
        r = subprocess.run([sys.executable, "-c", _CHAIN_SCRIPT, root, path], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[name] = dict(np.load(path))
    assert set(res["chain"]) == set(res["serial"]) and len(res["chain"]) == 7 * 6
    for k, a in res["chain"].items():
        b = res["serial"][k]
        if k.split("_")[0] in ("color", "depth", "n"):      # forward products: bit-identical
            assert np.array_equal(a, b), k
        else:
            assert rel_err(a, b) <= 2e-5, k
