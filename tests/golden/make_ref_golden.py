"""Generates tests/golden/ref_small_*.npz by running the UNMODIFIED reference kernels
(oracle/_ref/libgsref.so, built from /root/reference by oracle/build_ref.sh) on a GPU:

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out'      # then copy the .npz into tests/golden/

Each fixture stores the seeded scene parameters (not the arrays: they are regenerated from the seed by
tests/scenes.py), the upstream gradients' seed, and every reference output:
radii, tiles_touched, depth keys, means2D, conic/opacity, rgb, per-tile ranges, the sorted point list,
colour/depth/opacity images, final_T, n_contrib, n_touched and all gradients incl. dL/dtau.
The CPU tests pin the oracle against these files; the GPU tests pin the CUDA product against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

CASES = {
    "ref_small_sh0": dict(cfg=dict(W=96, H=80, fx=90.0, fy=88.0, cx=48.0, cy=40.0, P=400, sh_degree=0), seed=21,
                          scale_mul=2.5, bg=(0.0, 0.0, 0.0), grad_seed=5),
    "ref_small_sh3": dict(cfg=dict(W=112, H=64, fx=100.0, fy=100.0, cx=55.5, cy=31.5, P=300, sh_degree=3), seed=22,
                          scale_mul=3.0, bg=(0.2, 0.4, 0.1), grad_seed=6),
}


def build_case(c):
    import scenes as S

    sc = S.make_scene(c["cfg"], seed=c["seed"])
    sc["scales"] = sc["scales"] * np.float32(c["scale_mul"])
    sc["bg"] = np.asarray(c["bg"], np.float32)
    dc, dd = S.make_pixel_grads(c["cfg"]["W"], c["cfg"]["H"], seed=c["grad_seed"])
    # smooth, non-trivial upstream gradients (sign noise alone hides ordering mistakes)
    yy, xx = np.mgrid[0:c["cfg"]["H"], 0:c["cfg"]["W"]]
    dc = (dc * (1.0 + 0.5 * np.sin(xx / 9.0))[None]).astype(np.float32)
    dd = (dd * (1.0 + 0.5 * np.cos(yy / 7.0))[None]).astype(np.float32)
    return sc, dc, dd


def main(outdir):
    from common import RefLib

    os.makedirs(outdir, exist_ok=True)
    ref = RefLib()
    for name, c in CASES.items():
        sc, dc, dd = build_case(c)
        r = ref.forward(sc)
        g = ref.backward(sc, dc, dd)
        out = {("fwd_" + k): v for k, v in r.items() if isinstance(v, np.ndarray)}
        out.update({("bwd_" + k): v for k, v in g.items() if isinstance(v, np.ndarray)})
        out["num_rendered"] = np.int64(r["num_rendered"])
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print(name, "R =", r["num_rendered"], "visible =", int(r["visible"].sum()), "dL_dtau =", g["dL_dtau"])
    ref.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
