#!/usr/bin/env python
"""One small forward + backward through the public operator and through the engine (graph-free), for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from common import run_ours
import scenes as S
from diff_gaussian_rasterization import slam_ops as SO
from diff_gaussian_rasterization.engine import RasterEngine

for cfg, mult in ((dict(W=80, H=64, fx=75.0, fy=75.0, cx=40.0, cy=32.0, P=1500, sh_degree=0), 2.0),
                  (dict(W=64, H=48, fx=60.0, fy=60.0, cx=32.0, cy=24.0, P=3000, sh_degree=1), 12.0)):     # second: lists > 2048
    sc = S.make_scene(cfg, seed=1)
    sc["scales"] = sc["scales"] * mult
    dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"])
    o = run_ours(sc, dc, dd)
    print("public path R", o["num_rendered"], "max list", int((o["ranges"][:, 1] - o["ranges"][:, 0]).max()), "tau", o["dL_dtau"][:3])
    if cfg["sh_degree"] == 0:
        t = S.to_torch(sc, "cuda")
        eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                           cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"])
        cam = RasterEngine.pack_camera(*(torch.from_numpy(sc[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos")))
        eng.set_camera(cam.cuda())
        eng.dL_dcolor.copy_(torch.from_numpy(dc)); eng.dL_ddepth.copy_(torch.from_numpy(dd))
        eng.calibrate()
        eng.step(use_graph=False)
        ws = SO.LossWorkspace(cfg["W"], cfg["H"])
        SO.slam_loss(ws, eng.color, eng.depth, eng.opacity, eng.color * 0.9, eng.depth * 1.1, None, None, tracking=False)
        torch.cuda.synchronize()
        print("engine R", eng.header(), "loss", float(ws.sums[0]))
