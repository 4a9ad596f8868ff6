"""GPU parity of the tile-binning stage on the inputs that stress its sort: exact depth ties (order must fall
back to the Gaussian id, reference key = tile<<32 | depth bits with a stable sort over ascending ids,
rasterizer_impl.cu:98-108,353-358), depth clusters with outliers (wide key range, dense buckets), tile lists
longer than one shared-memory chunk and longer than the shared-memory capacity, and single-entry / empty tiles.
Bar: ranges and per-tile sorted lists bit-exact against the UNMODIFIED reference kernels."""
import os

import numpy as np
import pytest

from common import REF_LIB, RefLib, rel_err, run_ours

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libgsref.so not built")
    r = RefLib()
    yield r
    r.close()


def _identity_scene(W, H, P, seed, f=150.0):
    """Camera at the identity pose: camera-space depth == world z exactly, so depth bits can be planted."""
    import scenes as S

    cam = S.make_camera(W, H, f, f, W / 2 - 0.5, H / 2 - 0.5, np.eye(4))
    g = S.make_gaussians(P, cam, seed=seed)
    sc = dict(g)
    sc.update({k: cam[k] for k in ("image_width", "image_height", "tanfovx", "tanfovy", "viewmatrix", "projmatrix",
                                   "projmatrix_raw", "campos")})
    sc.update(bg=np.zeros(3, np.float32), scale_modifier=1.0, prefiltered=False, debug=False)
    return sc


def _check(o, r):
    assert o["overflow"] == 0
    assert o["num_rendered"] == r["num_rendered"]
    np.testing.assert_array_equal(o["radii"], r["radii"])
    np.testing.assert_array_equal(o["ranges"], r["ranges"])
    np.testing.assert_array_equal(o["point_list"], r["point_list"])
    for k in ("color", "depth", "opacity"):
        assert rel_err(o[k], r[k]) <= 1e-4, k


@pytest.mark.parametrize("levels", [3, 40, 4000])
def test_exact_depth_ties(ref, levels):
    """levels=3: hundreds of equal keys per tile (the transposition finisher gives up -> id-first radix path);
    levels=40: runs of equal keys; levels=4000: occasional pairs."""
    sc = _identity_scene(160, 128, 6000, seed=3)
    z = sc["means3D"][:, 2]
    q = np.float32(5.5 / levels)
    sc["means3D"][:, 2] = np.where(z > 0.3, np.round(z / q) * q, z).astype(np.float32)
    o, r = run_ours(sc), ref.forward(sc)
    v = r["visible"]
    assert len(np.unique(r["depths"][v])) <= levels + 2
    _check(o, r)


def test_depth_cluster_with_outliers(ref):
    """A surface at 2 m +- 1 cm plus 1 % outliers between 0.5 and 6 m: the key range spans 25 bits while most keys
    share their leading bits."""
    sc = _identity_scene(160, 128, 8000, seed=4)
    rng = np.random.default_rng(5)
    z = (2.0 + 0.01 * rng.standard_normal(8000)).astype(np.float32)
    out = rng.random(8000) < 0.01
    z[out] = rng.uniform(0.5, 6.0, int(out.sum())).astype(np.float32)
    scale = z / sc["means3D"][:, 2]
    sc["means3D"] = (sc["means3D"] * scale[:, None]).astype(np.float32)     # keeps the projection, moves the depth
    _check(run_ours(sc), ref.forward(sc))


@pytest.mark.parametrize("P,scale", [(15000, 6.0), (30000, 40.0)])
def test_long_tile_lists(ref, P, scale):
    """(15000, x6): lists of 2-4 k entries (several shared-memory chunks); (30000, x40) on a 64x48 image: every tile
    holds ~all Gaussians (> 12288 entries: the global ping-pong path)."""
    W, H = (160, 128) if P == 15000 else (64, 48)
    sc = _identity_scene(W, H, P, seed=6, f=120.0)
    sc["scales"] = (sc["scales"] * scale).astype(np.float32)
    o, r = run_ours(sc), ref.forward(sc)
    n = r["ranges"][:, 1].astype(np.int64) - r["ranges"][:, 0]
    assert n.max() > (2048 if P == 15000 else 12288)
    _check(o, r)


def test_sparse_tiles(ref):
    """Few tiny Gaussians: most tiles empty (range (0,0) like the reference's memset), many single-entry lists."""
    sc = _identity_scene(320, 240, 300, seed=7)
    sc["scales"] = (sc["scales"] * 0.2).astype(np.float32)
    o, r = run_ours(sc), ref.forward(sc)
    n = r["ranges"][:, 1].astype(np.int64) - r["ranges"][:, 0]
    assert (n == 0).sum() > 0 and (n == 1).sum() > 0
    _check(o, r)


def _order_for(sc, device="cuda"):
    """gsr_spatial_order for scene `sc` through the C-ABI (plan first), as a numpy permutation."""
    import ctypes as C

    import torch

    import diff_gaussian_rasterization as dgr
    import scenes as S
    from common import settings_from_scene

    t = S.to_torch(sc, device)
    e = torch.empty(0)
    call = dgr._Call(settings_from_scene(t), t["means3D"], t["shs"], e, t["opacities"], t["scales"], t["rotations"], e)
    P = call.P
    gb = dgr._L.gsr_geometry_bytes(P, call.W, call.H)
    geom = torch.empty(gb, dtype=torch.uint8, device=device)
    radii = torch.empty(P, dtype=torch.int32, device=device)
    nt = torch.empty(P, dtype=torch.int32, device=device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda x: C.c_void_p(x.data_ptr())
    dgr._cabi.check(dgr._L.gsr_forward_plan(C.byref(call.scene), p(geom), gb, p(radii), p(nt), st), "plan")
    order = torch.empty(P, dtype=torch.int32, device=device)
    dgr._cabi.check(dgr._L.gsr_spatial_order(C.byref(call.scene), p(geom), gb, p(order), st), "order")
    torch.cuda.synchronize()
    return order, radii.cpu().numpy()


@pytest.mark.parametrize("W,H,big", [(640, 480, False), (1208, 680, True)])
def test_scatter_in_spatial_order(ref, W, H, big):
    """gsr_scene.spatial_order (the scatter walks the Gaussians bucketed by home tile, box-local histogram): ranges and
    complete lists bit-exact against the reference kernels -- with the order of THIS view, with the order of another pose (stale:
    only locality suffers) and with an arbitrary permutation; a few Gaussians scaled up 25x make some CTAs' boxes exceed the
    shared-memory box (per-instance path inside the same launch)."""
    import torch

    sc = _identity_scene(W, H, 20000, seed=11, f=0.8 * W)
    rng = np.random.default_rng(12)
    s = sc["scales"].copy()
    if big:
        s[rng.random(20000) < 0.02] *= 25.0
    s[rng.random(20000) < 0.3] *= 0.15
    sc["scales"] = s.astype(np.float32)
    r = ref.forward(sc)
    order, radii = _order_for(sc)
    o_np = order.cpu().numpy()
    assert np.array_equal(np.sort(o_np), np.arange(20000))                    # a permutation
    vis = radii[o_np] > 0
    # culled ones share the first bucket with the few visible Gaussians whose home is tile 0
    assert np.flatnonzero(~vis).max() < int((~vis).sum()) + 200
    other = _identity_scene(W, H, 20000, seed=11, f=0.8 * W)
    other["scales"] = sc["scales"]
    other["means3D"] = (sc["means3D"] + np.array([0.3, -0.2, 0.1], np.float32)).astype(np.float32)
    stale, _ = _order_for(other)
    perm = torch.from_numpy(rng.permutation(20000).astype(np.int32)).cuda()
    for name, od in (("own", order), ("stale", stale), ("random", perm)):
        o = run_ours(sc, spatial_order=od)
        _check(o, r)
