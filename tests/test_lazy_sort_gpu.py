"""Per-tile lists ordered on demand (include/gsr_b200.h gsr_sort_on_demand; render.cu): the forward compositing kernel
selects and sorts one depth slab of a tile's list at a time and stops as soon as every pixel of the tile is opaque.
Bar: every output, n_contrib and final_T BIT-identical to the run that sorts every list completely (same order, same
arithmetic); point_list bit-exact against the UNMODIFIED reference kernels on every position up to the tile's deepest
contributor (positions behind it are never read: forward.cu:497-502, backward.cu:763); gradients within 1e-4."""
import os

import numpy as np
import pytest

from common import REF_LIB, RefLib, rel_err, run_ours
from test_binning_gpu import _identity_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libgsref.so not built")
    r = RefLib()
    yield r
    r.close()


def _tile_top(o, W, H):
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nc = np.zeros((gy * 16, gx * 16), np.int64)
    nc[:H, :W] = o["n_contrib"]
    return nc.reshape(gy, 16, gx, 16).max(axis=(1, 3)).ravel()


def _check_on_demand(sc, ref, threshold, grads=True, expect_partial=True):
    import scenes as S

    W, H = int(sc["image_width"]), int(sc["image_height"])
    dc, dd = S.make_pixel_grads(W, H, seed=11)
    full = run_ours(sc, dc, dd)                              # complete lists
    lazy = run_ours(sc, dc, dd, on_demand=threshold)
    r = ref.forward(sc)
    assert lazy["overflow"] == 0 and lazy["num_rendered"] == r["num_rendered"]
    np.testing.assert_array_equal(lazy["ranges"], r["ranges"])
    for k in ("color", "depth", "opacity", "final_T", "n_contrib", "n_touched", "radii"):
        np.testing.assert_array_equal(lazy[k], full[k], err_msg=k)
    top = _tile_top(lazy, W, H)
    start = r["ranges"][:, 0].astype(np.int64)
    n = r["ranges"][:, 1].astype(np.int64) - start
    assert (top <= n).all()
    bad = [t for t in range(n.size)
           if not np.array_equal(lazy["point_list"][start[t]:start[t] + top[t]], r["point_list"][start[t]:start[t] + top[t]])]
    assert not bad, "tiles whose consumed list prefix differs from the reference: %s" % bad[:8]
    if expect_partial:      # the point of the exercise: some tiles were NOT sorted to the end
        assert (top[n > threshold] < n[n > threshold]).any()
    if grads:
        for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dtau", "dL_dmean2D"):
            assert rel_err(lazy[k], full[k]) <= 1e-5, k
        rb = ref.backward(sc, dc, dd)
        for k in ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dtau"):
            assert rel_err(lazy[k], rb[k]) <= 1e-4, k
    return lazy, full, r


@pytest.mark.parametrize("threshold", [64, 300, 1024])
def test_on_demand_matches_complete_sort(ref, threshold):
    """Opaque scene, lists of 1-3 k entries of which a few hundred are read."""
    sc = _identity_scene(160, 128, 15000, seed=6, f=120.0)
    sc["scales"] = (sc["scales"] * 5.0).astype(np.float32)
    sc["opacities"] = np.maximum(sc["opacities"], np.float32(0.6))
    _check_on_demand(sc, ref, threshold)


def test_on_demand_transparent_scene_reads_everything(ref):
    """Low opacities: no pixel ever saturates, every slab of every list is selected and sorted in turn."""
    sc = _identity_scene(96, 80, 6000, seed=8, f=120.0)
    sc["scales"] = (sc["scales"] * 4.0).astype(np.float32)
    sc["opacities"] = np.full_like(sc["opacities"], 0.004)
    lazy, full, r = _check_on_demand(sc, ref, 64, expect_partial=False)
    np.testing.assert_array_equal(lazy["point_list"], r["point_list"])      # everything was read, so everything is ordered


@pytest.mark.parametrize("levels", [3, 40])
def test_on_demand_with_depth_ties(ref, levels):
    """Masses of exactly equal depths: whole slabs share one key (the transposition finisher gives up inside a slab) and
    single histogram bins outgrow a shared-memory chunk (the CTA falls back to the complete sort mid-way)."""
    sc = _identity_scene(160, 128, 6000 if levels == 40 else 20000, seed=3)
    if levels == 3:
        sc["scales"] = (sc["scales"] * 4.0).astype(np.float32)
    z = sc["means3D"][:, 2]
    q = np.float32(5.5 / levels)
    sc["means3D"][:, 2] = np.where(z > 0.3, np.round(z / q) * q, z).astype(np.float32)
    _check_on_demand(sc, ref, 64, expect_partial=False)


def test_on_demand_depth_cluster(ref):
    """A surface at 2 m +- 1 cm plus outliers from 0.5 to 6 m: most of the list falls into a handful of histogram bins."""
    P = 12000
    sc = _identity_scene(160, 128, P, seed=4)
    rng = np.random.default_rng(5)
    z = (2.0 + 0.01 * rng.standard_normal(P)).astype(np.float32)
    out = rng.random(P) < 0.01
    z[out] = rng.uniform(0.5, 6.0, int(out.sum())).astype(np.float32)
    scale = z / sc["means3D"][:, 2]
    sc["means3D"] = (sc["means3D"] * scale[:, None]).astype(np.float32)
    sc["scales"] = (sc["scales"] * 3.0).astype(np.float32)
    _check_on_demand(sc, ref, 64, expect_partial=False)


def test_on_demand_very_long_lists(ref):
    """Every tile holds ~all 30 k Gaussians (lists far beyond the shared-memory capacity) and reads a few hundred."""
    sc = _identity_scene(64, 48, 30000, seed=6, f=120.0)
    sc["scales"] = (sc["scales"] * 40.0).astype(np.float32)
    sc["opacities"] = np.maximum(sc["opacities"], np.float32(0.5))
    lazy, full, r = _check_on_demand(sc, ref, 1024)
    n = r["ranges"][:, 1].astype(np.int64) - r["ranges"][:, 0]
    assert n.max() > 12288


def test_engine_default_is_on_demand_and_matches():
    """The engine path (no host sync, CUDA graph) with the library default threshold on a scene of long lists."""
    import torch
    import diff_gaussian_rasterization as dgr
    import scenes as S
    from diff_gaussian_rasterization.engine import RasterEngine

    assert dgr._L.gsr_sort_on_demand(-1) > 0
    sc = _identity_scene(160, 128, 15000, seed=6, f=120.0)
    sc["scales"] = (sc["scales"] * 6.0).astype(np.float32)
    sc["opacities"] = np.maximum(sc["opacities"], np.float32(0.6))
    dc, dd = S.make_pixel_grads(160, 128, seed=11)
    full = run_ours(sc, dc, dd)
    t = S.to_torch(sc, "cuda")
    eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                       160, 128, sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=0)
    eng.set_camera(RasterEngine.pack_camera(*(torch.from_numpy(sc[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).cuda())
    eng.dL_dcolor.copy_(torch.from_numpy(dc)); eng.dL_ddepth.copy_(torch.from_numpy(dd))
    eng.calibrate()
    for use_graph in (False, True, True):
        eng.step(use_graph=use_graph)
        torch.cuda.synchronize()
        assert np.array_equal(eng.color.cpu().numpy(), full["color"])
        assert np.array_equal(eng.n_touched.cpu().numpy(), full["n_touched"])
        assert rel_err(eng.g_tau.cpu().numpy(), full["dL_dtau"]) <= 1e-5
        assert rel_err(eng.g_means3D.cpu().numpy(), full["dL_dmeans3D"]) <= 1e-5


# ---- depth partition of the segments (gsr_scene.depth_cut; binning.cu scatter_kernel<true>, render.cu two-part lists) ----
def _check_partition(sc, ref, threshold, cuts):
    """Scatter in spatial order with per-tile depth hints `cuts` (device int32[tiles], in/out): every output, n_contrib,
    final_T, n_touched bit-identical to the run without hints, point_list bit-exact against the reference on the consumed prefix,
    gradients within 1e-5 of the run without hints.  Returns the updated hints."""
    import scenes as S
    from test_binning_gpu import _order_for

    W, H = int(sc["image_width"]), int(sc["image_height"])
    dc, dd = S.make_pixel_grads(W, H, seed=11)
    order, _ = _order_for(sc)
    plain = run_ours(sc, dc, dd, on_demand=threshold)
    part = run_ours(sc, dc, dd, on_demand=threshold, spatial_order=order, depth_cut=cuts)
    r = ref.forward(sc)
    assert part["overflow"] == 0 and part["num_rendered"] == r["num_rendered"]
    np.testing.assert_array_equal(part["ranges"], r["ranges"])
    for k in ("color", "depth", "opacity", "final_T", "n_contrib", "n_touched", "radii"):
        np.testing.assert_array_equal(part[k], plain[k], err_msg=k)
    top = _tile_top(part, W, H)
    start = r["ranges"][:, 0].astype(np.int64)
    bad = [t for t in range(start.size)
           if not np.array_equal(part["point_list"][start[t]:start[t] + top[t]], r["point_list"][start[t]:start[t] + top[t]])]
    assert not bad, "tiles whose consumed list prefix differs from the reference: %s" % bad[:8]
    for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dtau", "dL_dmean2D"):
        assert rel_err(part[k], plain[k]) <= 1e-5, k
    return cuts


@pytest.mark.parametrize("kind", ["inf", "random", "zero", "iterated"])
def test_depth_partition_gives_the_same_results_for_any_hint(ref, kind):
    """The hint only moves pairs between the two ends of a tile's segment: +inf (everything in front: the plain path), random
    depths (front parts of every size, many tiles must go on into their back part), 0 (everything behind), and the hints the
    forward itself writes, fed back three times (the steady state of a SLAM loop)."""
    import torch

    sc = _identity_scene(160, 128, 15000, seed=6, f=120.0)
    sc["scales"] = (sc["scales"] * 5.0).astype(np.float32)
    sc["opacities"] = np.maximum(sc["opacities"], np.float32(0.6))
    tiles = 10 * 8
    if kind == "inf":
        cuts = torch.full((tiles,), 0x7f800000, dtype=torch.int32, device="cuda")
    elif kind == "zero":
        cuts = torch.zeros((tiles,), dtype=torch.int32, device="cuda")
    else:
        g = torch.Generator().manual_seed(3)
        depth = (torch.rand(tiles, generator=g) * 6.0).float()
        cuts = depth.view(torch.int32).cuda() if kind == "random" else torch.full((tiles,), 0x7f800000, dtype=torch.int32, device="cuda")
    before = cuts.clone()
    for rep in range(3 if kind == "iterated" else 1):
        _check_partition(sc, ref, 300, cuts)
    after = cuts.cpu().numpy()
    # the forward wrote next iteration's hints: finite depths for tiles that closed, +inf for tiles with an open pixel
    assert (after != before.cpu().numpy()).any() or kind == "inf"
    assert ((after == 0x7f800000) | ((after.view(np.float32) > 0.2) & (after.view(np.float32) < 50.0))).all()
    if kind == "iterated":
        assert (after != 0x7f800000).mean() > 0.5      # an opaque scene: most tiles close and carry a cut


def test_depth_partition_transparent_scene_and_depth_ties(ref):
    """No pixel ever closes (every tile reads front AND back part to the end); and exact depth ties across the cut."""
    import torch

    sc = _identity_scene(96, 80, 6000, seed=8, f=120.0)
    sc["scales"] = (sc["scales"] * 4.0).astype(np.float32)
    sc["opacities"] = np.minimum(sc["opacities"], np.float32(0.02))
    tiles = 6 * 5
    cuts = torch.tensor(np.full(tiles, 2.5, np.float32)).view(torch.int32).cuda()
    _check_partition(sc, ref, 64, cuts)
    assert (cuts.cpu().numpy() == 0x7f800000).all()          # open tiles: no cut for the next iteration
    sc = _identity_scene(160, 128, 6000, seed=3)
    z = sc["means3D"][:, 2]
    q = np.float32(5.5 / 40)
    sc["means3D"][:, 2] = np.where(z > 0.3, np.round(z / q) * q, z).astype(np.float32)
    sc["scales"] = (sc["scales"] * 3.0).astype(np.float32)
    cuts = torch.tensor(np.full(10 * 8, float(q * 12), np.float32)).view(torch.int32).cuda()      # exactly ON a depth level
    _check_partition(sc, ref, 64, cuts)
