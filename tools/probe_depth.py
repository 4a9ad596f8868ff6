#!/usr/bin/env python
"""How deep do the tiles go?  Per workload: list length per tile vs the deepest contributor of the tile
(positions behind it are never composited, forward or backward).
    python tools/probe_depth.py C2_replica_mapping C4_large_mapping
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import scenes as S
from common import run_ours

for name in sys.argv[1:]:
    cfg = S.CONFIGS[name]
    sc = S.make_scene(name, seed=0)
    o = run_ours(sc)
    W, H = cfg["W"], cfg["H"]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nc = np.zeros((gy * 16, gx * 16), np.int64)
    nc[:H, :W] = o["n_contrib"]
    top = nc.reshape(gy, 16, gx, 16).max(axis=(1, 3)).ravel()
    n = (o["ranges"][:, 1].astype(np.int64) - o["ranges"][:, 0].astype(np.int64))
    R = int(n.sum())
    print(name, "R", R, "tiles", n.size, "mean list", n.mean(), "max list", n.max(), "p50/p90/p99", np.percentile(n, [50, 90, 99]))
    print("  sum(top)/R = %.3f   mean top %.1f  p50/p90/p99 top %s  mean top/n %.3f" % (top.sum() / R, top.mean(), np.percentile(top, [50, 90, 99]), np.mean(top / np.maximum(n, 1))))
    for cap in (2048, 4096, 8192):
        print("  lists > %d: %d tiles holding %.1f%% of R" % (cap, (n > cap).sum(), 100.0 * n[n > cap].sum() / R))
