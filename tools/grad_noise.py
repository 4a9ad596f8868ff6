#!/usr/bin/env python
"""Where do the last 1e-4 of the gradient comparison come from?  For one view of a named config: the reference's own
run-to-run spread (its fp32 atomics are unordered), this build's spread, and this build against the reference with the
compositing backward in its fast (ex2.approx + rcp.approx) and exact (expf + division) forms.
    python tools/grad_noise.py C3_batched_tracking"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import scenes as S
from common import RefLib, l2_err, rel_err, run_ours

name = sys.argv[1] if len(sys.argv) > 1 else "C3_batched_tracking"
cfg = S.CONFIGS[name]
sc = S.make_scene(name, seed=0)
dc, dd = S.make_pixel_grads(cfg["W"], cfg["H"], seed=1)
ref = RefLib()
ref.forward(sc)
r1 = ref.backward(sc, dc, dd)
r2 = ref.backward(sc, dc, dd)
o1 = run_ours(sc, dc, dd, on_demand=None, exact_exp=1, lean=True)
o1b = run_ours(sc, dc, dd, on_demand=None, exact_exp=1, lean=True)
o2 = run_ours(sc, dc, dd, on_demand=None, exact_exp=2, lean=True)
o0 = run_ours(sc, dc, dd, on_demand=None, exact_exp=-1, lean=True)
keys = ("dL_dmeans3D", "dL_dmean2D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dsh", "dL_dtau")
print("%s: max-norm / L2 relative differences" % name)
print("%-14s %-19s %-19s %-19s %-19s %-19s" % ("", "ref vs ref", "ours vs ours", "fwd exact, bwd fast", "fwd+bwd exact", "all fast"))
for k in keys:
    b = np.asarray(r1[k])
    f = lambda a, bb: "%.2e / %.2e" % (rel_err(a, np.asarray(bb).reshape(np.asarray(a).shape)), l2_err(a, np.asarray(bb).reshape(np.asarray(a).shape)))
    print("%-14s %-19s %-19s %-19s %-19s %-19s" % (k, f(np.asarray(r2[k]), b), f(o1b[k], o1[k]), f(o1[k], b), f(o2[k], b), f(o0[k], b)))
