"""dL/dtau against the reference's analytic-Jacobian SCRIPT chain on identical inputs (north_star; VERDICT r1 "missing" #4).

tests/golden/script_chain_c0.npz holds what /root/reference/Loss_Derivative_script_compare.py computes on the C0
configuration (640x480, fx = fy = 577.5, pose w2c_gt @ T_noise from Jacob_test_result/*.txt, 15 Gaussians, SH degree 3):
projection (:772-971), dense dL/dmu_I, dL/dSigma_I, dL/ddepth_i, dL/dcolour_i (:1173-1351) -- the arrays the script saves as
Jacob_test_result/{grad_mu_I_pixel,grad_Sigma_I_pixel,grad_depth_per_gaussian}.npy --, the analytic Jacobians of every
Gaussian (:633-760) and the chain rule (:1587-1695, dL_dtau.npy).  Generator: tests/golden/make_script_chain_golden.py
(slices and executes the reference source unchanged).

Where the script and the rasterizer differ BY CONSTRUCTION, and how each difference is handled here:
  (1) the script is the dense "math version" of compositing: alpha = clip(o G, 0, 1) for every (pixel, Gaussian) pair -- no
      3-sigma tile rectangle, no alpha < 1/255 skip, no 0.99 clamp, no T < 1e-4 stop (forward.cu:481-507).  The CPU oracle
      has a switch for exactly that formulation (gso_set_math_mode) and must reproduce the script's arrays to float32
      round-off (1e-5); with the cut-offs on (the rasterizer's semantics) the same quantities move by 0.5-4 % on this
      scene, which bounds the direct CUDA-vs-script comparison below (stated per term);
  (2) the script's mean term multiplies a PIXEL-space gradient dL/dmu_I with a Jacobian scaled to NDC
      (diag(2fx/W, 2fy/H) . dmu_n/dtau, :743-758) instead of diag(fx, fy): its dL_dtau.npy is therefore not the
      rasterizer's grad_tau (the script's own closing comment records the mismatch, :1697-1705).  With the mean term
      rescaled to pixel units the script chain equals the oracle's dL/dtau (and hence autograd's, test_oracle_autograd.py);
      covariance, depth and SH terms are taken as the script computes them;
  (3) the analytic Jacobians know neither the 1.3 tan(fov) clamp of t (B18) nor the +0.3 low-pass (no tau dependence):
      inactive / irrelevant on this scene.
"""
import os

import numpy as np
import pytest

from common import rel_err
from oracle.gs_oracle import Oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "script_chain_c0.npz"))
W, H = 640, 480
N = int(G["means3D"].shape[0])
IDX = G["indices"]          # depth-sorted position -> Gaussian


def _scene():
    return dict(means3D=G["means3D"], opacities=G["opacities"], shs=G["shs"], cov3D_precomp=G["cov3D"], image_height=H, image_width=W,
                tanfovx=float(G["tanfov"][0]), tanfovy=float(G["tanfov"][1]), bg=np.zeros(3, np.float32), scale_modifier=1.0,
                viewmatrix=G["viewmatrix"], projmatrix=G["projmatrix"], projmatrix_raw=G["projmatrix_raw"], sh_degree=3, campos=G["campos"])


def _upstream():
    """the script's loss gradient (:1225-1233): sign of the L1 residuals under the mask, un-normalised"""
    mask = G["mask"]
    gc = np.sign(G["rendered_color"] - G["gt_color"]) * mask[None]
    gd = np.sign(G["rendered_depth"] - G["gt_depth"]) * ((G["gt_depth"] > 0) & mask[None])
    return gc.astype(np.float32), gd.astype(np.float32)


def _script_chain_pixel_units():
    """The script's chain rule (:1587-1695) with its mean term in consistent units: pixel-space dL/dmu_I times
    diag(fx, fy) . (normalised-coordinate Jacobian).  Everything comes from the arrays the reference code produced."""
    fx, fy = float(G["intrinsics"][0, 0]), float(G["intrinsics"][1, 1])
    dmu_norm = G["dmu_I_dT_all"] / np.array([2 * fx / W, 2 * fy / H])[None, :, None]
    mu = sum(G["grad_mu_I_pixel"][i].astype(np.float64) @ (np.diag([fx, fy]) @ dmu_norm[IDX[i]]) for i in range(N))
    return mu + G["dL_dtau_cov"] + G["dL_dtau_depth"] + G["dL_dtau_sh"]


def _conic_grad_to_cov_grad(conic_opacity, dL_dconic):
    """dL/dSigma (2x2, both off-diagonals filled like the script's) from the rasterizer's dL/dconic (B14: xy stored once)."""
    co = conic_opacity[IDX]
    Cm = np.stack([np.stack([co[:, 0], co[:, 1]], 1), np.stack([co[:, 1], co[:, 2]], 1)], 1)
    M = dL_dconic[IDX]
    M = np.stack([np.stack([M[:, 0, 0], M[:, 0, 1]], 1), np.stack([M[:, 0, 1], M[:, 1, 1]], 1)], 1)
    return -Cm @ M @ Cm


def test_script_arrays_have_the_schema_of_the_shipped_npy_files():
    assert G["grad_mu_I_pixel"].shape == (15, 2) and G["grad_mu_I_pixel"].dtype == np.float32
    assert G["grad_Sigma_I_pixel"].shape == (15, 2, 2) and G["grad_Sigma_I_pixel"].dtype == np.float32
    assert G["grad_depth_per_gaussian"].shape == (15,) and G["grad_depth_per_gaussian"].dtype == np.float32
    assert G["dL_dtau"].shape == (6,) and G["dL_dtau"].dtype == np.float64


def test_oracle_projection_matches_the_script():
    st = Oracle(np.float64).preprocess(_scene())
    assert (st["radii"] > 0).all()
    assert np.array_equal(np.argsort(st["depths"], kind="stable"), IDX)          # OrderGaussiansByDepth, :764-770
    assert rel_err(st["means2D"][IDX], G["mean_2D"]) <= 1e-9
    co = st["conic_opacity"][IDX]
    cov = np.linalg.inv(np.stack([np.stack([co[:, 0], co[:, 1]], 1), np.stack([co[:, 1], co[:, 2]], 1)], 1))
    assert rel_err(cov, G["cov_2D"]) <= 1e-7
    assert rel_err(st["rgb"][IDX], G["color"]) <= 1e-6                            # SH degree 3 incl. one clamped channel
    assert rel_err(st["depths"][IDX], G["depth"]) <= 1e-9


def _oracle_dense(math_mode):
    """Oracle backward over DENSE lists (every tile sees all Gaussians in depth order, like the script's per-pixel loop
    over all Gaussians), with or without the rasterizer's cut-offs."""
    o = Oracle(np.float64)
    st = o.preprocess(_scene())
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    order = np.argsort(st["depths"], kind="stable").astype(np.uint32)
    st["point_list"] = np.tile(order, tiles).astype(np.uint32)
    st["ranges"] = np.stack([np.arange(tiles) * N, (np.arange(tiles) + 1) * N], 1).astype(np.uint32)
    gc, gd = _upstream()
    o.L.gso_set_math_mode_f64(1 if math_mode else 0)
    try:
        o.render(st)
        g = o.backward(st, gc, gd)
    finally:
        o.L.gso_set_math_mode_f64(0)
    return st, g


def test_oracle_in_math_mode_reproduces_the_script_chain():
    st, g = _oracle_dense(math_mode=True)
    gm = g["dL_dmean2D"][IDX][:, :2] / np.array([0.5 * W, 0.5 * H])               # NDC units (backward.cu:837-838) -> pixels
    assert rel_err(gm, G["grad_mu_I_pixel"]) <= 1e-5
    assert rel_err(_conic_grad_to_cov_grad(st["conic_opacity"], g["dL_dconic"]), G["grad_Sigma_I_pixel"]) <= 1e-5
    assert rel_err(g["dL_ddepth"][IDX, 0], G["grad_depth_per_gaussian"]) <= 1e-5
    assert rel_err(g["dL_dcolor"][IDX], G["grad_color_per_gaussian"]) <= 1e-5
    chain = _script_chain_pixel_units()
    assert rel_err(g["dL_dtau"], chain) <= 5e-4       # float32 accumulation of the script's dense gradients over 307 200 pixels
    # the script's literal dL_dtau (mean term in mixed units) is NOT that gradient -- difference (2) of the header
    assert rel_err(G["dL_dtau"], chain) > 0.1


def test_cutoffs_move_the_script_quantities_by_a_few_percent():
    """Bounds difference (1): the same dense lists with the rasterizer's cut-offs on."""
    st, g = _oracle_dense(math_mode=False)
    gm = g["dL_dmean2D"][IDX][:, :2] / np.array([0.5 * W, 0.5 * H])
    assert 1e-4 < rel_err(gm, G["grad_mu_I_pixel"]) <= 3e-2
    assert rel_err(g["dL_ddepth"][IDX, 0], G["grad_depth_per_gaussian"]) <= 1.5e-2
    assert rel_err(g["dL_dtau"], _script_chain_pixel_units()) <= 6e-2


@pytest.mark.gpu
def test_cuda_dL_dtau_and_per_gaussian_gradients_against_the_script_chain():
    """The CUDA op on the script's inputs: tight against the oracle with the rasterizer's semantics (1e-4, the north star's
    fp32 tolerance), and within the cut-off bounds of difference (1) against the script's own arrays."""
    from common import run_ours

    sc = _scene()
    gc, gd = _upstream()
    o = run_ours(sc, gc, gd)
    orc = Oracle(np.float64)
    st = orc.forward(sc)
    g = orc.backward(st, gc, gd)
    assert np.array_equal(o["radii"], st["radii"])
    assert rel_err(o["color"], st["color"]) <= 1e-4 and rel_err(o["depth"], st["depth"]) <= 1e-4
    assert rel_err(o["dL_dtau"], g["dL_dtau"]) <= 1e-4
    assert rel_err(o["dL_dmean2D"][:, :2], g["dL_dmean2D"][:, :2]) <= 1e-4
    assert rel_err(o["dL_dcov3D"], g["dL_dcov3D"]) <= 1e-4
    assert rel_err(o["dL_dsh"], g["dL_dsh"]) <= 1e-4
    # ... and against what the reference's script computed (math version, bounds from the CPU test above)
    gm = o["dL_dmean2D"][IDX][:, :2] / np.array([0.5 * W, 0.5 * H])
    assert rel_err(gm, G["grad_mu_I_pixel"]) <= 3e-2
    assert rel_err(o["dL_dtau"], _script_chain_pixel_units()) <= 6e-2
    assert np.all(np.sign(o["dL_dtau"][[0, 2, 3, 4]]) == np.sign(_script_chain_pixel_units()[[0, 2, 3, 4]]))
