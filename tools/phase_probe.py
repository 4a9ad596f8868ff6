#!/usr/bin/env python
"""Phase timeline of the three per-Gaussian kernels (needs a library built with GSR_PHASE_PROBE=1):
    GSR_PHASE_PROBE=1 python gs-slam-analytica_jacobian_b200/build.py --force && python tools/phase_probe.py
Prints, per kernel, when (us after the kernel's first CTA started) the CTAs reached each phase boundary."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from diff_gaussian_rasterization import _cabi
import scenes as S
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
buf = np.zeros((5, 4096, 8), np.uint64)
_cabi.check(L.gsr_debug_probe(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), buf.nbytes), "probe")
tiles = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)
for k, (kname, nph) in enumerate((("preprocess_forward", 8), ("scatter", 5), ("preprocess_backward", 4), ("render_forward", 5))):
    nblk = min(4096, tiles if k == 3 else (cfg["P"] + 255) // 256)
    tb = buf[k, :nblk, :nph].astype(np.int64)
    t0 = tb[:, 0].min()
    print("== %s: %d CTAs; columns = phase boundary, rows = min / median / max over CTAs [us since first CTA start]" % (kname, nblk))
    for ph in range(nph):
        col = tb[:, ph]
        col = col[col > 0] - t0
        if col.size:
            print("   phase %d: n=%4d  min %7.2f  med %7.2f  max %7.2f" % (ph, col.size, col.min() / 1e3, np.median(col) / 1e3, col.max() / 1e3))

# render forward: per-CTA durations of its phases (0 start, 1 list ready [sort / first slab], 2 end of the first run of
# batches, 3 a further slab ordered, 4 compositing finished)
tb = buf[3, :min(4096, tiles), :5].astype(np.int64)
ok = tb[:, 0] > 0
d01 = (tb[ok, 1] - tb[ok, 0]) / 1e3
d12 = (tb[ok, 2] - tb[ok, 1]) / 1e3
d04 = (tb[ok, 4] - tb[ok, 0]) / 1e3
more = (tb[ok, 3] > tb[ok, 0]).mean()
print("render_forward per CTA [us]: list ready med %.2f p90 %.2f | first run of batches med %.2f p90 %.2f | whole CTA med %.2f p90 %.2f | CTAs that ordered a further slab: %.1f %%"
      % (np.median(d01), np.percentile(d01, 90), np.median(d12), np.percentile(d12, 90), np.median(d04), np.percentile(d04, 90), 100 * more))

# render backward: 0 start, 1 tile flag acquired, 2 per-pixel state loaded + first batch of records staged (the prologue a
# compositing CTA that ran the backward right behind its own forward would not have), 3 end
tbb = buf[4, :min(4096, tiles), :4].astype(np.int64)
okb = tbb[:, 0] > 0
if okb.any():
    w01 = (tbb[okb, 1] - tbb[okb, 0]) / 1e3
    w12 = (tbb[okb, 2] - tbb[okb, 1]) / 1e3
    w03 = (tbb[okb, 3] - tbb[okb, 0]) / 1e3
    print("render_backward per CTA [us]: flag wait med %.2f p90 %.2f | loads + first gather med %.2f p90 %.2f | whole CTA med %.2f p90 %.2f | prologue share of the CTA time: %.1f %%"
          % (np.median(w01), np.percentile(w01, 90), np.median(w12), np.percentile(w12, 90), np.median(w03), np.percentile(w03, 90),
             100 * (w01.sum() + w12.sum()) / w03.sum()))

# scheduling headroom of the compositing kernel: greedy list scheduling of the measured CTA durations on the resident slots
import heapq


def makespan(durs, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    for d in durs:
        t = heapq.heappop(h)
        heapq.heappush(h, t + d)
    return max(h)


dur = d04
slots = 148 * 4
print("render_forward scheduling: %d CTAs, sum/slots = %.1f us, max CTA %.1f us, makespan in launch order %.1f us, longest first %.1f us, measured span %.1f us"
      % (dur.size, dur.sum() / slots, dur.max(), makespan(list(dur), slots), makespan(sorted(dur, reverse=True), slots),
         (tb[ok, 4].max() - tb[ok, 0].min()) / 1e3))
