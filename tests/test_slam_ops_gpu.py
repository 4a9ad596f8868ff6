"""GPU parity of the fused callers of the rasterizer (slam_ops: loss + gradients, Adam + pose update + camera tensors)
against a plain-torch restatement of the reference (oracle/slam_ref.py, pinned to the reference modules by tests/test_slam_golden.py) with torch autograd / torch.optim.Adam, and an
end-to-end check of the graph-captured tracking loop."""
import numpy as np
import pytest
import torch

from oracle import slam_ref as R
from common import rel_err

pytestmark = pytest.mark.gpu


def _images(W, H, seed):
    g = torch.Generator().manual_seed(seed)
    color = torch.rand((3, H, W), generator=g)
    depth = torch.rand((1, H, W), generator=g) * 4
    opacity = torch.rand((1, H, W), generator=g) * 0.2 + 0.85          # straddles the 0.95 opacity mask
    gt = torch.rand((3, H, W), generator=g)
    gt[:, : H // 8] = 0.001                                              # below the rgb boundary threshold
    gtd = torch.rand((1, H, W), generator=g) * 4
    gtd[:, :, : W // 10] = 0.0                                           # invalid depth
    gmask = torch.rand((1, H, W), generator=g) > 0.3
    return color, depth, opacity, gt, gtd, gmask


@pytest.mark.parametrize("mode", ["track_rgbd", "track_mono", "map_rgbd", "map_mono", "map_init"])
def test_loss_and_gradients_match_autograd(mode):
    from diff_gaussian_rasterization import slam_ops as S

    W, H = 203, 117
    color, depth, opacity, gt, gtd, gmask = _images(W, H, 3)
    a, b = torch.tensor(0.07), torch.tensor(-0.03)
    thr, alpha = 0.01, 0.9
    c64, d64 = color.double().requires_grad_(True), depth.double().requires_grad_(True)
    a64, b64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
    mono = mode.endswith("mono")
    if mode.startswith("track"):
        loss = R.loss_tracking(c64, d64, opacity.double(), gt.double(), gtd.double(), gmask, a64, b64, thr, alpha, mono)
    else:
        loss = R.loss_mapping(c64, d64, gt.double(), gtd.double(), a64, b64, thr, alpha, mono or mode == "map_init",
                              initialization=(mode == "map_init"))
    loss.backward()
    ws = S.LossWorkspace(W, H)
    cu = lambda t: t.cuda().contiguous()
    sums = S.slam_loss(ws, cu(color), cu(depth), cu(opacity), cu(gt), None if (mono or mode == "map_init") else cu(gtd),
                       cu(gmask.to(torch.uint8)) if mode.startswith("track") else None,
                       None if mode == "map_init" else torch.stack([a, b]).cuda(), thr, alpha, tracking=mode.startswith("track"))
    sums = sums.cpu().numpy()
    assert abs(sums[0] - loss.item()) <= 1e-5 * abs(loss.item())
    assert rel_err(ws.dL_dcolor.cpu().numpy(), c64.grad.numpy()) <= 1e-6
    gd = d64.grad.numpy() if d64.grad is not None else np.zeros((1, H, W))
    assert rel_err(ws.dL_ddepth.cpu().numpy(), gd) <= 1e-6
    if mode != "map_init":
        assert abs(sums[1] - a64.grad.item()) <= 1e-4 * abs(a64.grad.item()) + 1e-9
        assert abs(sums[2] - b64.grad.item()) <= 1e-4 * abs(b64.grad.item()) + 1e-9
    # second call on the same workspace (ticket reset, no stale sums)
    sums2 = S.slam_loss(ws, cu(color), cu(depth), cu(opacity), cu(gt), None if (mono or mode == "map_init") else cu(gtd),
                        cu(gmask.to(torch.uint8)) if mode.startswith("track") else None,
                        None if mode == "map_init" else torch.stack([a, b]).cuda(), thr, alpha, tracking=mode.startswith("track"))
    np.testing.assert_array_equal(sums2.cpu().numpy(), sums)


def test_tracking_step_matches_adam_and_update_pose():
    import scenes as SC
    from diff_gaussian_rasterization import slam_ops as S

    cam = SC.make_camera(640, 480, 517.3, 516.5, 318.6, 255.3, SC.base_pose())
    w2c = torch.from_numpy(np.asarray(cam["w2c"], np.float64))
    Rm, Tm = w2c[:3, :3].clone(), w2c[:3, 3].clone()
    proj_raw = torch.from_numpy(cam["projmatrix_raw"]).double()
    rot = torch.zeros(3, dtype=torch.float64, requires_grad=True)
    trans = torch.zeros(3, dtype=torch.float64, requires_grad=True)
    ea = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    eb = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([{"params": [rot], "lr": 0.003}, {"params": [trans], "lr": 0.001}, {"params": [ea], "lr": 0.01},
                            {"params": [eb], "lr": 0.01}])                     # utils/slam_frontend.py:132-162
    pose = S.PoseState(Rm, Tm, cam["projmatrix_raw"])
    block = torch.zeros(52, dtype=torch.float32, device="cuda")
    g = torch.Generator().manual_seed(11)
    for it in range(6):
        scale = 1e-2 if it < 5 else 0.0                                         # last step: zero gradients
        gtau = (torch.randn(6, generator=g) * scale).double()
        gexp = (torch.randn(2, generator=g) * scale).double()
        rot.grad, trans.grad = gtau[3:].clone(), gtau[:3].clone()               # theta = tau[3:], rho = tau[:3]
        ea.grad, eb.grad = gexp[:1].clone(), gexp[1:].clone()
        opt.step()
        with torch.no_grad():
            Rm, Tm, conv = R.update_pose(Rm, Tm, trans.detach(), rot.detach())
            rot.zero_(); trans.zero_()                                          # pose_utils.py:91-92
        sums = torch.tensor([0.0, gexp[0], gexp[1], 0.0], dtype=torch.float32, device="cuda")
        S.tracking_step(pose, gtau.float().cuda(), sums, block)
        rt = pose.RT.cpu().double()
        assert rel_err(rt[:9].reshape(3, 3).numpy(), Rm.numpy()) <= 1e-5
        assert rel_err(rt[9:].numpy(), Tm.numpy()) <= 1e-5
        assert abs(pose.exposure[0].item() - ea.item()) <= 2e-6 and abs(pose.exposure[1].item() - eb.item()) <= 2e-6
        wvt, full, center = R.camera_tensors(Rm, Tm, proj_raw)
        blk = block.cpu().double().numpy()
        assert rel_err(blk[0:16].reshape(4, 4), wvt.numpy()) <= 1e-5
        assert rel_err(blk[16:32].reshape(4, 4), full.numpy()) <= 1e-5
        assert rel_err(blk[32:48].reshape(4, 4), proj_raw.numpy()) == 0.0
        assert rel_err(blk[48:51], center.numpy()) <= 1e-5
        st = pose.status.cpu().numpy()
        assert st[1] == it + 1
    # Adam with zero gradients still moves (momentum), so convergence must come from the threshold test, like the reference
    assert st[0] == int(conv)


def test_tracking_loop_graph_converges_towards_the_target_pose():
    import scenes as SC
    from diff_gaussian_rasterization import slam_ops as S
    from diff_gaussian_rasterization.engine import RasterEngine

    cfg = dict(W=320, H=240, fx=290.0, fy=290.0, cx=159.5, cy=119.5, P=20000, sh_degree=0)
    sc = SC.make_scene(cfg, seed=5)
    sc["scales"] = sc["scales"] * 1.5
    t = SC.to_torch(sc, "cuda")

    def engine():
        return RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                            cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)

    base = SC.base_pose()
    target = SC.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], base)
    eng = engine()
    eng.set_camera(RasterEngine.pack_camera(*(torch.from_numpy(target[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda())
    eng.calibrate()
    eng.launch_forward()
    gt_color, gt_depth = eng.color.clone(), eng.depth.clone()
    start = SC.se3_exp([0.01, -0.008, 0.012, 0.004, -0.003, 0.002]) @ base
    gmask = torch.ones((cfg["H"], cfg["W"]), dtype=torch.uint8, device="cuda")

    def run(use_graph):
        e = engine()
        pose = S.PoseState(start[:3, :3], start[:3, 3], target["projmatrix_raw"])
        loop = S.TrackingLoop(e, pose, gt_color, gt_depth, gmask, alpha=0.9)
        e.calibrate()
        loss0 = None
        n, first, overflow = loop.run(max_iters=60, check_every=60, use_graph=use_graph)
        assert n == 60 and not overflow
        return pose.RT.cpu().numpy().astype(np.float64), loop.ws.sums.cpu().numpy()

    rt_g, sums_g = run(True)
    rt_e, sums_e = run(False)
    err0 = np.linalg.norm(start[:3, 3] - base[:3, 3])
    err1 = np.linalg.norm(rt_g[9:] - base[:3, 3])
    assert err1 < 0.35 * err0, (err0, err1)                                 # the pose moved most of the way to the target
    assert rel_err(rt_g, rt_e) <= 1e-3                                       # graph replay == eager launches (fp32 atomics aside)


@pytest.mark.parametrize("fused", [True, False])
def test_mapping_window_with_fused_loss_matches_per_view_autograd(fused):
    """MappingWindow (render -> fused mapping loss -> backward, accumulated over the window) == per-view rasterizer calls fed
    with the autograd gradients of the restated get_loss_mapping (utils/slam_utils.py:92-128), summed like autograd does."""
    from common import run_ours
    import scenes as SC
    from diff_gaussian_rasterization import slam_ops as S
    from diff_gaussian_rasterization.engine import RasterEngine
    from diff_gaussian_rasterization.window import KeyframeWindow

    V = 3
    cfg = dict(W=208, H=160, fx=190.0, fy=188.0, cx=104.0, cy=80.0, P=6000, sh_degree=0)
    sc = SC.make_scene(cfg, seed=8)
    sc["scales"] = sc["scales"] * 2.0
    cams = [SC.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], w2c) for w2c in SC.arc_poses(V, radius=0.3)]
    t = SC.to_torch(sc, "cuda")
    eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                       cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)
    pack = lambda c: RasterEngine.pack_camera(*(torch.from_numpy(c[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda()
    g = torch.Generator().manual_seed(21)
    gt_c = torch.rand((V, 3, cfg["H"], cfg["W"]), generator=g).cuda()
    gt_d = (torch.rand((V, 1, cfg["H"], cfg["W"]), generator=g) * 3).cuda()
    expo = (torch.randn((V, 2), generator=g) * 0.05).cuda()
    win = KeyframeWindow(eng, torch.stack([pack(c) for c in cams]))
    win.calibrate()
    mw = S.MappingWindow(win, gt_c, gt_d, expo, alpha=0.9, fused=fused)      # loss in the forward's epilogue / stand-alone kernel
    flat, sums, tau = mw.iteration()
    torch.cuda.synchronize()
    expect = {k: 0.0 for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations")}
    for v in range(V):
        fwd = run_ours(SC.with_camera(sc, cams[v]))
        img = torch.from_numpy(fwd["color"]).double().requires_grad_(True)
        dep = torch.from_numpy(fwd["depth"]).double().requires_grad_(True)
        a, b = expo[v, 0].cpu().double().requires_grad_(True), expo[v, 1].cpu().double().requires_grad_(True)
        loss = R.loss_mapping(img, dep, gt_c[v].cpu().double(), gt_d[v].cpu().double(), a, b, 0.01, 0.9, False)
        loss.backward()
        assert abs(sums[v, 0].item() - loss.item()) <= 1e-5 * abs(loss.item())
        assert abs(sums[v, 1].item() - a.grad.item()) <= 1e-4 * abs(a.grad.item()) + 1e-9
        o = run_ours(SC.with_camera(sc, cams[v]), img.grad.float().numpy(), dep.grad.float().numpy())
        for k in expect:
            expect[k] = expect[k] + o[k].astype(np.float64)
        assert rel_err(tau[v].cpu().numpy(), o["dL_dtau"]) <= 1e-4
    for k, mine in (("dL_dmeans3D", eng.g_means3D), ("dL_dsh", eng.g_sh), ("dL_dopacity", eng.g_opacity), ("dL_dscales", eng.g_scales),
                    ("dL_drotations", eng.g_rot)):
        assert rel_err(mine.cpu().numpy(), expect[k]) <= 1e-4, k


@pytest.mark.parametrize("mode", ["track_rgbd", "track_mono", "map_rgbd"])
def test_loss_fused_into_the_forward_epilogue_matches_the_loss_kernel(mode):
    """gsr_fused_loss: dL/dcolor, dL/ddepth and {loss, dL/da, dL/db} written by the forward compositing kernel's epilogue ==
    the stand-alone loss kernel run on the images that forward produced; with and without lists ordered on demand."""
    import scenes as SC
    from diff_gaussian_rasterization import slam_ops as S
    from diff_gaussian_rasterization.engine import RasterEngine
    import diff_gaussian_rasterization as dgr

    cfg = dict(W=328, H=250, fx=290.0, fy=290.0, cx=163.5, cy=124.5, P=20000, sh_degree=0)      # ragged: 328 = 20.5 tiles
    sc = SC.make_scene(cfg, seed=5)
    sc["scales"] = sc["scales"] * 1.5
    t = SC.to_torch(sc, "cuda")
    g = torch.Generator().manual_seed(3)
    gt_c = torch.rand((3, cfg["H"], cfg["W"]), generator=g).cuda()
    gt_c[:, :20] = 0.0                                                    # below the rgb boundary threshold
    gt_d = (torch.rand((1, cfg["H"], cfg["W"]), generator=g) * 3).cuda()
    gt_d[:, -30:] = 0.0                                                   # invalid depth
    gmask = (torch.rand((cfg["H"], cfg["W"]), generator=g) > 0.3).to(torch.uint8).cuda()
    expo = torch.tensor([0.03, -0.02], dtype=torch.float32, device="cuda")
    tracking = mode.startswith("track")
    depth = None if mode.endswith("mono") else gt_d
    for lazy_min in (0, 64):
        prev = dgr._L.gsr_sort_on_demand(lazy_min)
        try:
            eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                               cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)
            eng.set_camera(RasterEngine.pack_camera(*(torch.from_numpy(sc[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda())
            eng.calibrate()
            ws_f, ws_k = S.LossWorkspace(cfg["W"], cfg["H"]), S.LossWorkspace(cfg["W"], cfg["H"])
            fl = S.FusedLoss(ws_f, gt_c, depth, gmask if tracking else None, expo, alpha=0.9, tracking=tracking)
            for rep in range(3):                                          # the ticket must re-arm itself
                ws_f.dL_dcolor.fill_(7.0); ws_f.dL_ddepth.fill_(7.0); ws_f.sums.fill_(7.0)
                eng.launch_forward(fused_loss=fl.struct)
                S.slam_loss(ws_k, eng.color, eng.depth, eng.opacity, gt_c, depth, gmask if tracking else None, expo, alpha=0.9,
                            tracking=tracking)
                torch.cuda.synchronize()
                assert rel_err(ws_f.dL_dcolor.cpu().numpy(), ws_k.dL_dcolor.cpu().numpy()) <= 1e-6
                assert rel_err(ws_f.dL_ddepth.cpu().numpy(), ws_k.dL_ddepth.cpu().numpy()) <= 1e-6
                assert rel_err(ws_f.sums.cpu().numpy()[:3], ws_k.sums.cpu().numpy()[:3]) <= 1e-5
        finally:
            dgr._L.gsr_sort_on_demand(prev)


def test_tracking_loop_with_fused_loss_follows_the_unfused_loop():
    import scenes as SC
    from diff_gaussian_rasterization import slam_ops as S
    from diff_gaussian_rasterization.engine import RasterEngine

    cfg = dict(W=320, H=240, fx=290.0, fy=290.0, cx=159.5, cy=119.5, P=20000, sh_degree=0)
    sc = SC.make_scene(cfg, seed=5)
    sc["scales"] = sc["scales"] * 1.5
    t = SC.to_torch(sc, "cuda")

    def engine():
        return RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                            cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)

    base = SC.base_pose()
    target = SC.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], base)
    eng = engine()
    eng.set_camera(RasterEngine.pack_camera(*(torch.from_numpy(target[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda())
    eng.calibrate()
    eng.launch_forward()
    gt_color, gt_depth = eng.color.clone(), eng.depth.clone()
    start = SC.se3_exp([0.01, -0.008, 0.012, 0.004, -0.003, 0.002]) @ base
    gmask = torch.ones((cfg["H"], cfg["W"]), dtype=torch.uint8, device="cuda")

    def run(fused, use_graph):
        e = engine()
        pose = S.PoseState(start[:3, :3], start[:3, 3], target["projmatrix_raw"])
        loop = S.TrackingLoop(e, pose, gt_color, gt_depth, gmask, alpha=0.9, fused=fused)
        e.calibrate()
        n, first, overflow = loop.run(max_iters=40, check_every=40, use_graph=use_graph)
        assert n == 40 and not overflow
        return pose.RT.cpu().numpy().astype(np.float64), loop.ws.sums.cpu().numpy(), pose.exposure.cpu().numpy()

    rt_u, sums_u, ex_u = run(False, True)
    for use_graph in (True, False):
        rt_f, sums_f, ex_f = run(True, use_graph)
        assert rel_err(rt_f, rt_u) <= 1e-3
        assert rel_err(sums_f[:3], sums_u[:3]) <= 5e-2           # the loss after 40 chaotic Adam steps: same ballpark
        assert rel_err(ex_f, ex_u) <= 5e-2
