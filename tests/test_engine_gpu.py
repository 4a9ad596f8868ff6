"""GPU tests of RasterEngine (persistent workspaces, no host sync, CUDA graphs) and KeyframeWindow
(per-view accumulation into one flat gradient buffer) against the plain per-call path."""
import os

import numpy as np
import pytest
import torch

from common import rel_err, run_ours

pytestmark = pytest.mark.gpu


def _scene_and_cams(P=6000, V=4):
    import scenes as S

    cfg = dict(W=208, H=160, fx=190.0, fy=188.0, cx=104.0, cy=80.0, P=P, sh_degree=0)
    sc = S.make_scene(cfg, seed=8)
    sc["scales"] = sc["scales"] * 2.0
    cams = [S.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], w2c)
            for w2c in S.arc_poses(V, radius=0.3)]
    return cfg, sc, cams


def _engine(sc, cfg, **kw):
    import scenes as S
    from diff_gaussian_rasterization.engine import RasterEngine

    t = S.to_torch(sc, "cuda")
    return RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
                             rotations=t["rotations"]), cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"],
                        sh_degree=cfg["sh_degree"], **kw)


def _pack(cam):
    from diff_gaussian_rasterization.engine import RasterEngine

    return RasterEngine.pack_camera(*(torch.from_numpy(cam[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda()


def test_engine_graph_matches_per_call_path():
    import scenes as S

    cfg, sc, cams = _scene_and_cams(V=3)
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"], seed=2)
    eng = _engine(sc, cfg)
    eng.dL_dcolor.copy_(torch.from_numpy(dc))
    eng.dL_ddepth.copy_(torch.from_numpy(dd))
    for cam in cams:
        eng.set_camera(_pack(cam))
        eng.calibrate()
    eng.capture()
    for cam in cams:                      # same graph replayed at different poses
        eng.set_camera(_pack(cam))
        R = eng.step_checked(use_graph=True)
        ref = run_ours(S.with_camera(sc, cam), dc, dd)
        assert R == ref["num_rendered"]
        assert np.array_equal(eng.radii.cpu().numpy(), ref["radii"])
        assert np.array_equal(eng.n_touched.cpu().numpy(), ref["n_touched"])
        assert np.array_equal(eng.color.cpu().numpy(), ref["color"])        # forward is deterministic
        assert rel_err(eng.g_tau.cpu().numpy(), ref["dL_dtau"]) <= 1e-5
        assert rel_err(eng.g_means3D.cpu().numpy(), ref["dL_dmeans3D"]) <= 1e-5
        assert rel_err(eng.g_rot.cpu().numpy(), ref["dL_drotations"]) <= 1e-5
        assert rel_err(eng.g_sh.cpu().numpy(), ref["dL_dsh"]) <= 1e-5


def test_engine_overflow_is_detected_and_recovered():
    import scenes as S

    cfg, sc, cams = _scene_and_cams(V=1)
    eng = _engine(sc, cfg)
    eng.set_camera(_pack(cams[0]))
    R = eng.calibrate()
    # shrink the workspace behind the engine's back: the next step must flag and recover
    eng.capacity = 0
    eng.ensure_capacity(R // 4)
    eng.capacity_before = eng.capacity
    got = eng.step_checked(use_graph=False)
    assert got == R and eng.capacity > eng.capacity_before
    ref = run_ours(S.with_camera(sc, cams[0]))
    assert np.array_equal(eng.color.cpu().numpy(), ref["color"])


def test_window_accumulates_views():
    import scenes as S
    from diff_gaussian_rasterization.window import KeyframeWindow

    V = 4
    cfg, sc, cams = _scene_and_cams(V=V)
    eng = _engine(sc, cfg)
    packed = torch.stack([_pack(c) for c in cams])
    grads = [S.make_pixel_grads(cfg["W"], cfg["H"], seed=10 + v) for v in range(V)]
    gc = torch.stack([torch.from_numpy(g[0]) for g in grads]).cuda()
    gd = torch.stack([torch.from_numpy(g[1]) for g in grads]).cuda()
    win = KeyframeWindow(eng, packed)
    win.calibrate()
    win.iteration((gc, gd))
    torch.cuda.synchronize()
    refs = [run_ours(S.with_camera(sc, cams[v]), grads[v][0], grads[v][1]) for v in range(V)]
    for name, mine in (("dL_dmeans3D", eng.g_means3D), ("dL_dsh", eng.g_sh), ("dL_dopacity", eng.g_opacity),
                       ("dL_dscales", eng.g_scales), ("dL_drotations", eng.g_rot)):
        expect = sum(r[name].astype(np.float64) for r in refs)
        assert rel_err(mine.cpu().numpy(), expect) <= 1e-5, name
    for i in range(V):
        assert rel_err(win.tau[i].cpu().numpy(), refs[i]["dL_dtau"]) <= 1e-5
    # a second iteration must not see leftovers of the first (first local view overwrites)
    win.iteration((gc, gd))
    expect = sum(r["dL_dmeans3D"].astype(np.float64) for r in refs)
    assert rel_err(eng.g_means3D.cpu().numpy(), expect) <= 1e-5


def test_densification_statistics_in_backward_epilogue():
    """xyz_gradient_accum / denom / max_radii2D updated by the backward kernel == the reference's masked torch ops after
    each view (gaussian_splatting/scene/gaussian_model.py:767-771, utils/slam_backend.py:115-121)."""
    import scenes as S
    from diff_gaussian_rasterization.window import KeyframeWindow

    V = 3
    cfg, sc, cams = _scene_and_cams(V=V)
    eng = _engine(sc, cfg)
    P = eng.P
    accum = torch.zeros((P, 1), dtype=torch.float32, device="cuda")
    denom = torch.zeros((P, 1), dtype=torch.float32, device="cuda")
    maxr = torch.zeros((P,), dtype=torch.float32, device="cuda")
    eng.attach_densification_stats(accum, denom, maxr)
    packed = torch.stack([_pack(c) for c in cams])
    grads = [S.make_pixel_grads(cfg["W"], cfg["H"], seed=20 + v) for v in range(V)]
    gc = torch.stack([torch.from_numpy(g[0]) for g in grads]).cuda()
    gd = torch.stack([torch.from_numpy(g[1]) for g in grads]).cuda()
    win = KeyframeWindow(eng, packed)
    win.calibrate()
    e_acc, e_den, e_max = np.zeros((P, 1)), np.zeros((P, 1)), np.zeros(P)

    def on_view(i, v):      # the reference's per-view statistics, from the per-view outputs
        nonlocal e_max
        vis = eng.radii.cpu().numpy() > 0                                         # visibility_filter = radii > 0
        g2 = eng.g_means2D.cpu().numpy().astype(np.float64)
        e_acc[vis] += np.linalg.norm(g2[vis, :2], axis=-1, keepdims=True)          # add_densification_stats
        e_den[vis] += 1
        e_max[vis] = np.maximum(e_max[vis], eng.radii.cpu().numpy()[vis])

    win.iteration((gc, gd), on_view=on_view)
    torch.cuda.synchronize()
    assert rel_err(accum.cpu().numpy(), e_acc) <= 1e-6
    np.testing.assert_array_equal(denom.cpu().numpy(), e_den)
    np.testing.assert_array_equal(maxr.cpu().numpy(), e_max)
    eng.attach_densification_stats()      # detach: a further backward must leave the statistics alone
    win.iteration((gc, gd))
    np.testing.assert_array_equal(denom.cpu().numpy(), e_den)


def test_host_driven_step_graph_matches_resident_step():
    """RasterEngine.capture_host_step / step_host (pinned camera block + upstream gradients copied inside the graph,
    results copied back) == the device-resident step at the same pose."""
    import scenes as S

    cfg, sc, cams = _scene_and_cams(V=3)
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"], seed=4)
    eng = _engine(sc, cfg)
    for cam in cams:
        eng.set_camera(_pack(cam))
        eng.calibrate()
    cam_pin = torch.empty(52, dtype=torch.float32).pin_memory()
    dc_pin, dd_pin = torch.from_numpy(dc).pin_memory(), torch.from_numpy(dd).pin_memory()
    cam_pin.copy_(_pack(cams[0]).cpu())
    eng.capture_host_step(cam_pin, dc_pin, dd_pin)
    for cam in cams:
        cam_pin.copy_(_pack(cam).cpu())
        tau, hdr = eng.step_host()
        ref = run_ours(S.with_camera(sc, cam), dc, dd)
        assert int(hdr[0]) == ref["num_rendered"] and int(hdr[1]) == 0
        assert rel_err(tau.numpy(), ref["dL_dtau"]) <= 1e-5
        assert np.array_equal(eng.color.cpu().numpy(), ref["color"])


def test_fused_sort_survives_lists_longer_than_the_hint():
    """The engine decides from its calibrated longest-list hint whether the forward kernel sorts its own tiles (lists up to one
    shared-memory chunk).  If a later pose produces longer lists than the hint promised, the fused kernel must still sort them
    (general path) -- slower, never wrong."""
    import scenes as S

    cfg = dict(W=160, H=128, fx=120.0, fy=120.0, cx=79.5, cy=63.5, P=15000, sh_degree=0)
    sc = S.make_scene(cfg, seed=6)
    sc["scales"] = sc["scales"] * 6.0                       # lists of 2-4 k entries
    cam = S.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], S.base_pose())
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"], seed=3)
    eng = _engine(sc, cfg)
    eng.set_camera(_pack(cam))
    eng.dL_dcolor.copy_(torch.from_numpy(dc)); eng.dL_ddepth.copy_(torch.from_numpy(dd))
    eng.calibrate()
    assert eng.max_tile_hint > 2048
    eng.max_tile_hint = 512                                  # a stale, far too small hint -> fused path
    eng.graph_fwd = eng.graph_bwd = eng.graph_all = None
    eng.step(use_graph=False)
    R, ov = eng.header()
    ref = run_ours(S.with_camera(sc, cam), dc, dd)
    assert R == ref["num_rendered"] and not ov
    assert np.array_equal(eng.color.cpu().numpy(), ref["color"])
    assert rel_err(eng.g_tau.cpu().numpy(), ref["dL_dtau"]) <= 1e-5
    assert rel_err(eng.g_means3D.cpu().numpy(), ref["dL_dmeans3D"]) <= 1e-5


def test_window_with_two_engines_on_two_streams_matches_one_engine():
    import scenes as S
    from diff_gaussian_rasterization.window import KeyframeWindow

    V = 5
    cfg, sc, cams = _scene_and_cams(V=V)
    packed = torch.stack([_pack(c) for c in cams])
    grads = [S.make_pixel_grads(cfg["W"], cfg["H"], seed=30 + v) for v in range(V)]
    gc = torch.stack([torch.from_numpy(g[0]) for g in grads]).cuda()
    gd = torch.stack([torch.from_numpy(g[1]) for g in grads]).cuda()
    e1 = _engine(sc, cfg)
    w1 = KeyframeWindow(e1, packed)
    w1.calibrate()
    flat1 = w1.iteration((gc, gd)).clone()
    tau1 = w1.tau.clone()
    ea = _engine(sc, cfg)
    eb = _engine(sc, cfg, grad_flat=ea.grad_flat)      # one gradient buffer, REDs from both streams
    w2 = KeyframeWindow(ea, packed, extra_engines=[eb])
    w2.calibrate()
    for _ in range(2):      # twice: the second iteration must not see leftovers of the first
        flat2 = w2.iteration((gc, gd))
        torch.cuda.synchronize()
        assert rel_err(flat2.cpu().numpy(), flat1.cpu().numpy()) <= 1e-5
        assert rel_err(w2.tau.cpu().numpy(), tau1.cpu().numpy()) <= 1e-5
    # the same iteration as ONE CUDA graph (both engine streams forked and joined inside the capture); new poses at replay
    graph = w2.capture((gc, gd))
    for shift in (0, 1):
        packed_new = torch.roll(packed, shift, 0)
        w1.cameras.copy_(packed_new); w2.cameras.copy_(packed_new)
        ref = w1.iteration((gc, gd)).clone()
        ref_tau = w1.tau.clone()
        w2.engine.grad_flat.fill_(7.0)
        graph.replay()
        torch.cuda.synchronize()
        assert rel_err(w2.engine.grad_flat.cpu().numpy(), ref.cpu().numpy()) <= 1e-5
        assert rel_err(w2.tau.cpu().numpy(), ref_tau.cpu().numpy()) <= 1e-5


def test_several_bands_of_one_view_on_one_rank_sum_their_pose_gradient():
    """plan_units(whole_bands=2) puts both bands of every view on the SAME rank: the backward kernel stores its dL/dtau, so the
    second band must not overwrite the first (window.tau_rows: spare rows merged before the reduction)."""
    import scenes as S
    from diff_gaussian_rasterization.window import KeyframeWindow

    V = 3
    cfg, sc, cams = _scene_and_cams(V=V)
    packed = torch.stack([_pack(c) for c in cams])
    grads = [S.make_pixel_grads(cfg["W"], cfg["H"], seed=40 + v) for v in range(V)]
    gc = torch.stack([torch.from_numpy(g[0]) for g in grads]).cuda()
    gd = torch.stack([torch.from_numpy(g[1]) for g in grads]).cuda()
    w1 = KeyframeWindow(_engine(sc, cfg), packed)
    w1.calibrate()
    flat1 = w1.iteration((gc, gd)).clone()
    tau1 = w1.tau_all.clone()
    n = flat1.numel() - 8 * w1.engine.tau_slots
    ea = _engine(sc, cfg)
    for extra in ([], [_engine(sc, cfg, grad_flat=ea.grad_flat)]):      # one stream, then two engines on two streams
        w2 = KeyframeWindow(ea, packed, extra_engines=extra, whole_bands=2)
        assert len(w2.units) == 2 * V and len(w2.tau_merges) == V
        w2.calibrate()
        for _ in range(2):
            flat2 = w2.iteration((gc, gd))
            torch.cuda.synchronize()
            assert rel_err(flat2[:n].cpu().numpy(), flat1[:n].cpu().numpy()) <= 1e-5
            assert rel_err(w2.tau_all.cpu().numpy(), tau1.cpu().numpy()) <= 1e-5


def test_bands_of_a_view_sum_to_the_view():
    """gsr_scene.tile_row_begin / tile_row_end: the bands of a view partition its pixels, and per-Gaussian gradients, dL/dtau
    and n_touched are linear in the pixels (window.py splits left-over keyframes over the ranks this way)."""
    import scenes as S
    from diff_gaussian_rasterization.window import band_rows

    cfg, sc, cams = _scene_and_cams(V=1)
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"], seed=41)
    eng = _engine(sc, cfg)
    eng.set_camera(_pack(cams[0]).cuda())
    eng.dL_dcolor.copy_(torch.from_numpy(dc)); eng.dL_ddepth.copy_(torch.from_numpy(dd))
    eng.calibrate()
    eng.step(use_graph=False)
    torch.cuda.synchronize()
    full = dict(color=eng.color.clone(), depth=eng.depth.clone(), opacity=eng.opacity.clone(), n_touched=eng.n_touched.clone(),
                radii=eng.radii.clone(), flat=eng.grad_flat.clone(), tau=eng.g_tau.clone(), m2d=eng.g_means2D.clone())
    grid_y = (cfg["H"] + 15) // 16
    for parts in (2, 3):
        color, depth, opacity = torch.zeros_like(eng.color), torch.zeros_like(eng.depth), torch.zeros_like(eng.opacity)
        n_touched, radii = torch.zeros_like(eng.n_touched), torch.zeros_like(eng.radii)
        tau, m2d = torch.zeros_like(eng.g_tau), torch.zeros_like(eng.g_means2D)
        for k, (y0, y1) in enumerate(band_rows(grid_y, parts)):
            eng.set_band(y0, y1)
            eng.calibrate()
            eng.launch_forward()
            eng.launch_backward(accumulate=(k > 0))
            torch.cuda.synchronize()
            R, ov = eng.header()
            assert not ov
            r0, r1 = 16 * y0, min(16 * y1, cfg["H"])
            color[:, r0:r1] = eng.color[:, r0:r1]; depth[:, r0:r1] = eng.depth[:, r0:r1]; opacity[:, r0:r1] = eng.opacity[:, r0:r1]
            n_touched += eng.n_touched
            radii = torch.maximum(radii, eng.radii)
            tau += eng.g_tau
            m2d += eng.g_means2D
        flat = eng.grad_flat.clone()
        # the last band once more as a captured graph (forward + overlapped backward): same per-band results
        tau_band, nt_band = eng.g_tau.clone(), eng.n_touched.clone()
        eng.step(use_graph=True)
        torch.cuda.synchronize()
        assert rel_err(eng.g_tau.cpu().numpy(), tau_band.cpu().numpy()) <= 2e-5 and torch.equal(eng.n_touched, nt_band)
        assert torch.equal(color, full["color"]) and torch.equal(depth, full["depth"]) and torch.equal(opacity, full["opacity"])
        assert torch.equal(n_touched, full["n_touched"])
        # a Gaussian's radius is reported by every band its rectangle reaches; one that no band renders has no tiles at all
        assert torch.equal(radii, full["radii"])
        assert rel_err(flat.cpu().numpy(), full["flat"].cpu().numpy()) <= 2e-5
        assert rel_err(tau.cpu().numpy(), full["tau"].cpu().numpy()) <= 2e-5
        assert rel_err(m2d.cpu().numpy(), full["m2d"].cpu().numpy()) <= 2e-5
    eng.set_band(0, 0)


def test_window_plan_splits_leftover_views_and_matches_unsplit_window():
    """10 views on 8 ranks (simulated one rank at a time on this GPU): every rank holds one whole view + one band; the sum of
    the ranks' buffers is the unsplit window's gradient and every view's dL/dtau."""
    import scenes as S
    from diff_gaussian_rasterization.window import KeyframeWindow, plan_units

    V, world = 5, 4
    cfg, sc, cams = _scene_and_cams(V=V)
    packed = torch.stack([_pack(c) for c in cams]).cuda()
    grads = [S.make_pixel_grads(cfg["W"], cfg["H"], seed=60 + v) for v in range(V)]
    gc = torch.stack([torch.from_numpy(g[0]) for g in grads]).cuda()
    gd = torch.stack([torch.from_numpy(g[1]) for g in grads]).cuda()
    e = _engine(sc, cfg)
    w = KeyframeWindow(e, packed)
    w.calibrate()
    ref = w.iteration((gc, gd)).clone()
    grid_y = (cfg["H"] + 15) // 16
    plan = plan_units(V, world, grid_y)
    assert [len(u) for u in plan] == [2, 2, 2, 2] and all(u[1][2] > u[1][1] for u in plan)      # whole view + band of view 4
    total = torch.zeros_like(ref)
    for r in range(world):
        er = _engine(sc, cfg)
        wr = KeyframeWindow(er, packed, rank=r, world_size=world)
        wr.calibrate()
        total += wr.iteration((gc, gd), reduce=False)
        torch.cuda.synchronize()
    assert rel_err(total.cpu().numpy(), ref.cpu().numpy()) <= 2e-5
    assert rel_err(total[-8 * e.tau_slots:].cpu().numpy(), ref[-8 * e.tau_slots:].cpu().numpy()) <= 2e-5     # every view's dL/dtau


def test_bands_and_spatial_orders_on_the_two_kernel_path_in_a_fresh_process():
    """The window of three ranks emulated on one GPU with the cooperative preprocess + scatter kernel switched off
    (GSR_NO_FUSED_SCATTER=1, read when the library loads): every unit -- whole views and bands of tile rows -- then runs the
    two-kernel path with ITS spatial order (RasterEngine.use_order / gsr_spatial_order).  The ranks' gradient buffers must sum
    to the unsplit, unordered window."""
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path[:0] = [%r, %r, %r]\n"
        "import scenes as S\n"
        "from common import rel_err\n"
        "from diff_gaussian_rasterization.engine import RasterEngine\n"
        "from diff_gaussian_rasterization.window import KeyframeWindow\n"
        "cfg = dict(W=320, H=240, fx=290.0, fy=290.0, cx=159.5, cy=119.5, P=20000, sh_degree=0)\n"
        "sc = S.make_scene(cfg, seed=5); sc['scales'] = sc['scales'] * 1.5\n"
        "t = S.to_torch(sc, 'cuda')\n"
        "mk = lambda **kw: RasterEngine(dict(means3D=t['means3D'], opacities=t['opacities'], shs=t['shs'], scales=t['scales'], rotations=t['rotations']),\n"
        "                               cfg['W'], cfg['H'], sc['tanfovx'], sc['tanfovy'], sc['bg'], sh_degree=0, **kw)\n"
        "V = 4\n"
        "cams = [S.make_camera(cfg['W'], cfg['H'], cfg['fx'], cfg['fy'], cfg['cx'], cfg['cy'], w2c) for w2c in S.arc_poses(V, radius=0.3)]\n"
        "pack = lambda c: RasterEngine.pack_camera(*(torch.from_numpy(c[k]) for k in ('viewmatrix', 'projmatrix', 'projmatrix_raw', 'campos'))).cuda()\n"
        "packed = torch.stack([pack(c) for c in cams])\n"
        "grads = [S.make_pixel_grads(cfg['W'], cfg['H'], seed=30 + v) for v in range(V)]\n"
        "gc = torch.stack([torch.from_numpy(g[0]) for g in grads]).cuda(); gd = torch.stack([torch.from_numpy(g[1]) for g in grads]).cuda()\n"
        "e0 = mk(); e0.use_spatial_order = False\n"
        "w0 = KeyframeWindow(e0, packed); w0.calibrate()\n"
        "ref = w0.iteration((gc, gd)).clone(); ref_tau = w0.tau_all.clone()\n"
        "total = torch.zeros_like(ref); orders = 0\n"
        "for r in range(3):\n"
        "    ea = mk(); eb = mk(grad_flat=ea.grad_flat)\n"
        "    assert ea.use_spatial_order      # two-kernel path: the engine builds orders\n"
        "    w = KeyframeWindow(ea, packed, rank=r, world_size=3, extra_engines=[eb]); w.calibrate()\n"
        "    orders += len(ea.spatial_orders) + len(eb.spatial_orders)\n"
        "    assert any(u[2] > 0 for u in w.units)      # a band among the rank's units\n"
        "    total += w.iteration((gc, gd), reduce=False)\n"
        "    for e in (ea, eb):\n"
        "        assert not e.header()[1]\n"
        "assert orders >= 4\n"
        "n = ref.numel() - 8 * e0.tau_slots\n"
        "assert rel_err(total[:n].cpu().numpy(), ref[:n].cpu().numpy()) <= 1e-5\n"
        "assert rel_err(total[n:].view(-1, 8)[:V, :6].cpu().numpy(), ref_tau.cpu().numpy()) <= 1e-5\n"
        "print('ok')\n"
    ) % (here, os.path.join(root, "gs-slam-analytica_jacobian_b200"), root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, GSR_NO_FUSED_SCATTER="1"))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]
