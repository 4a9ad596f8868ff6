#!/usr/bin/env python
"""Phase timeline of the three per-Gaussian kernels (needs a library built with GSR_PHASE_PROBE=1):
    GSR_PHASE_PROBE=1 python gs-slam-analytica_jacobian_b200/build.py --force && python tools/phase_probe.py
Prints, per kernel, when (us after the kernel's first CTA started) the CTAs reached each phase boundary."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from diff_gaussian_rasterization import _cabi
from diff_gaussian_rasterization import scenes as S
from diff_gaussian_rasterization.engine import RasterEngine

name = sys.argv[1] if len(sys.argv) > 1 else "C1_tum_tracking"
cfg = S.CONFIGS[name]
sc = S.make_scene(name, seed=0)
t = S.to_torch(sc, "cuda")
eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                   cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=cfg["sh_degree"])
cam = RasterEngine.pack_camera(*(torch.from_numpy(sc[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos")))
eng.set_camera(cam.cuda())
eng.calibrate()
eng.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    flush.fill_(1)
    eng.step()
torch.cuda.synchronize()
L = _cabi.load()
buf = np.zeros((3, 4096, 8), np.uint64)
_cabi.check(L.gsr_debug_probe(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), buf.nbytes), "probe")
nblk = min(4096, (cfg["P"] + 255) // 256)
for k, (kname, nph) in enumerate((("preprocess_forward", 7), ("scatter", 5), ("preprocess_backward", 4))):
    tb = buf[k, :nblk, :nph].astype(np.int64)
    t0 = tb[:, 0].min()
    print("== %s: %d CTAs; columns = phase boundary, rows = min / median / max over CTAs [us since first CTA start]" % (kname, nblk))
    for ph in range(nph):
        col = tb[:, ph]
        col = col[col > 0] - t0
        if col.size:
            print("   phase %d: n=%4d  min %7.2f  med %7.2f  max %7.2f" % (ph, col.size, col.min() / 1e3, np.median(col) / 1e3, col.max() / 1e3))
