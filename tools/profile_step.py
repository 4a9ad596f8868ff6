#!/usr/bin/env python
"""Runs a few un-graphed forward+backward steps of a named workload (for ncu / compute-sanitizer).
    python tools/profile_step.py [workload] [steps] [P]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import scenes as S
from diff_gaussian_rasterization.engine import RasterEngine

name = sys.argv[1] if len(sys.argv) > 1 else "C1_tum_tracking"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = int(sys.argv[3]) if len(sys.argv) > 3 else None
cfg = S.CONFIGS[name]
sc = S.make_scene(name, seed=0, P=P)
t = S.to_torch(sc, "cuda")
eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                   cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=cfg["sh_degree"])
cam = RasterEngine.pack_camera(*(torch.from_numpy(sc[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos")))
eng.set_camera(cam.cuda())
dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"])
eng.dL_dcolor.copy_(torch.from_numpy(dc))
eng.dL_ddepth.copy_(torch.from_numpy(dd))
R = eng.calibrate()
for _ in range(steps):
    eng.step(use_graph=False)
torch.cuda.synchronize()
R2, ov = eng.header()
print("workload", name, "P", eng.P, "R", R, R2, "overflow", ov, "dL_dtau", eng.g_tau.cpu().numpy())
