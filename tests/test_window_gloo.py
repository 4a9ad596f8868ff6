"""N > 1 host logic of the keyframe-sharded mapping window on CPU: world_size-2 (and 3) gloo process groups.
The CUDA engine cannot run here, so every rank fabricates deterministic per-view gradients; what is under
test is the owner map, the accumulate-then-one-all-reduce protocol and that ranks without views still take
part in the collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_grad(v, n):
    g = torch.Generator().manual_seed(1000 + v)
    return torch.randn(n, generator=g, dtype=torch.float32)


GRID_Y = 43      # tile rows of the 1200x680 mapping shape


def _worker(rank, world, port, V, n, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import allreduce_window_gradients, plan_units

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = plan_units(V, world, GRID_Y)[rank]
    # engine protocol: every unit adds into ONE flat buffer whose tail holds a row of dL/dtau per view; a unit that is a band
    # of tile rows contributes the band's share of its view (here: in proportion to its rows)
    flat = torch.zeros(n + 8 * V)
    for (v, y0, y1) in units:
        share = 1.0 if y1 == 0 else (y1 - y0) / GRID_Y
        flat[:n] += share * _view_grad(v, n)
        flat[n + 8 * v:n + 8 * v + 6] += share * _view_grad(100 + v, 6)
    allreduce_window_gradients(flat)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), flat.numpy())
    np.save(os.path.join(out_dir, "u%d.npy" % rank), np.asarray(units, np.int64).reshape(-1, 3))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,V", [(2, 10), (3, 2), (4, 10)])
def test_window_allreduce_gloo(tmp_path, world, V):
    n = 4099
    mp.spawn(_worker, args=(world, _free_port(), V, n, str(tmp_path)), nprocs=world, join=True)
    expect = torch.zeros(n + 8 * V)
    expect[:n] = sum(_view_grad(v, n) for v in range(V))
    for v in range(V):
        expect[n + 8 * v:n + 8 * v + 6] = _view_grad(100 + v, 6)
    rows = np.zeros((V, GRID_Y), np.int64)      # how often every tile row of every view is rendered
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npy" % r))
        np.testing.assert_allclose(got, expect.numpy(), rtol=1e-5, atol=1e-5)
        for v, y0, y1 in np.load(tmp_path / ("u%d.npy" % r)):
            rows[v, (y0 if y1 else 0):(y1 if y1 else GRID_Y)] += 1
    assert (rows == 1).all()                    # every tile row of every view rendered exactly once


def test_shard_map():
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import band_rows, owner_of, plan_units, shard_views

    assert [len(shard_views(10, 8, r)) for r in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]     # whole views only
    assert [len(shard_views(32, 8, r)) for r in range(8)] == [4] * 8
    assert shard_views(3, 4, 3) == []
    for v in range(10):
        assert v in shard_views(10, 8, owner_of(v, 8))
    # split plan: 10 keyframes on 8 ranks -> one whole view + a quarter of view 8 or 9 each
    plan = plan_units(10, 8, 43)
    assert all(len(u) == 2 and u[0] == (r, 0, 0) for r, u in enumerate(plan))
    assert [u[1][0] for u in plan] == [8, 8, 8, 8, 9, 9, 9, 9]
    assert [(u[1][1], u[1][2]) for u in plan[:4]] == band_rows(43, 4) and band_rows(43, 4)[0][0] == 0 and band_rows(43, 4)[-1][1] == 43
    assert plan_units(32, 8, 68) == [[(v, 0, 0) for v in range(r, 32, 8)] for r in range(8)]      # C4: nothing to split
    assert plan_units(10, 8, 43, split=False) == [[(v, 0, 0) for v in shard_views(10, 8, r)] for r in range(8)]
    assert plan_units(10, 1, 43) == [[(v, 0, 0) for v in range(10)]]
    # equal-work bands follow the weights
    assert band_rows(8, 2, [1, 1, 1, 1, 1, 1, 1, 9]) == [(0, 7), (7, 8)]
    for parts in (2, 3, 5):
        b = band_rows(43, parts, list(range(43)))
        assert b[0][0] == 0 and b[-1][1] == 43 and all(x[1] == y[0] for x, y in zip(b, b[1:])) and all(y1 > y0 for y0, y1 in b)


def _reducer_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diff_gaussian_rasterization.window import SwitchReducer      # needs libgsr_b200.so (built in-tree)

        red = SwitchReducer.create(1024, "cpu")      # no multicast memory on the CPU: every rank must agree on "unavailable"
        ok = red is None and SwitchReducer.last_error is not None
    except ImportError:
        ok = True      # library not built here: nothing to check
    np.save(os.path.join(out_dir, "s%d.npy" % rank), np.array([1 if ok else 0]))
    dist.barrier()
    dist.destroy_process_group()


def test_switch_reducer_reports_unavailable_consistently_without_multicast_memory(tmp_path):
    """SwitchReducer.create is a collective that must return None on EVERY rank (the caller then uses dist.all_reduce) when the
    symmetric / multicast allocation cannot be made -- here: CPU tensors under gloo."""
    mp.spawn(_reducer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert int(np.load(tmp_path / ("s%d.npy" % r))[0]) == 1


def test_plan_units_whole_bands():
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import band_rows, plan_units

    plan = plan_units(10, 8, 43, whole_bands=2)
    assert plan[0][:2] == [(0,) + band_rows(43, 2)[0], (0,) + band_rows(43, 2)[1]] and len(plan[0]) == 3
    rows = np.zeros((10, 43), np.int64)
    for units in plan:
        for v, y0, y1 in units:
            rows[v, (y0 if y1 else 0):(y1 if y1 else 43)] += 1
    assert (rows == 1).all()


def test_tau_rows_give_every_band_of_a_view_on_one_rank_its_own_row():
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import plan_units, tau_rows

    # default plan (10 keyframes on 8 ranks): every unit of a rank is a different view -> row = view, nothing to merge
    for units in plan_units(10, 8, 43):
        rows, merges = tau_rows(units, 10)
        assert rows == [u[0] for u in units] and merges == []
    # both bands of a whole view on the same rank: the second one writes a spare row behind the window's views
    units = plan_units(10, 8, 43, whole_bands=2)[0]
    rows, merges = tau_rows(units, 10)
    assert rows[:2] == [0, 10] and merges[0] == (10, 0)
    assert len(set(rows)) == len(rows) and all(r >= 10 for r, _ in merges)
    assert sorted(v for _, v in merges) == sorted(u[0] for u in units[1::2] if u[0] < 8)


def test_plans_cover_every_tile_row_of_every_view_exactly_once_randomised():
    """Property of plan_units / band_rows over random window sizes, rank counts, image heights and row weights: the units of all
    ranks partition (view, tile row); bands are consecutive and, while there are enough rows, non-empty."""
    import random

    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import band_rows, plan_units, tau_rows

    rnd = random.Random(7)
    for _ in range(3000):
        gy, parts = rnd.randint(1, 80), rnd.randint(1, 12)
        wts = rnd.choice([None, [0.0] * gy, [rnd.choice([0, 0, 0, 5.0]) for _ in range(gy)], [rnd.random() ** 4 * 100 for _ in range(gy)]])
        b = band_rows(gy, parts, wts)
        assert len(b) == parts and b[0][0] == 0 and b[-1][1] == gy
        assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1)) and all(y1 >= y0 for y0, y1 in b)
        assert gy < parts or all(y1 > y0 for y0, y1 in b)
    for _ in range(1500):
        V, N, gy, wb = rnd.randint(0, 40), rnd.randint(1, 9), rnd.randint(1, 70), rnd.randint(1, 3)
        plan = plan_units(V, N, gy, split=rnd.random() < 0.8, whole_bands=wb)
        cover = np.zeros((V, gy), np.int64)
        for units in plan:
            rows, merges = tau_rows(units, V)
            assert len(set(rows)) == len(rows) and all(0 <= v < V <= s for s, v in merges)
            for v, y0, y1 in units:
                cover[v, (y0 if y1 else 0):(y1 if y1 else gy)] += 1
        assert (cover == 1).all()


class _StandInEngine:
    """The part of RasterEngine a KeyframeWindow drives, on the CPU, with the kernels' write semantics: per-Gaussian gradients
    are overwritten unless `accumulate`, dL/dtau is STORED into tau_out; a band contributes its rows' share of the view."""

    def __init__(self, n, tau_slots=64):
        self.dev, self.H, self.tau_slots, self.n = torch.device("cpu"), GRID_Y * 16, tau_slots, n
        self.grad_flat = torch.full((n + 8 * tau_slots,), 7.0)      # stale contents: the window must not rely on zeros
        self.tau_block = self.grad_flat[n:].view(tau_slots, 8)
        self.view, self.band, self.calls = -1, (0, 0), 0

    def set_camera(self, cam):
        self.view = int(cam[0])

    def set_band(self, y0=0, y1=0):
        self.band = (int(y0), int(y1))

    def use_order(self, key):
        pass

    def calibrate(self, build_order=True):
        return 0

    def launch_forward(self, fused_loss=None):
        pass

    def launch_backward(self, dL_dcolor=None, dL_ddepth=None, accumulate=False, overlap_forward=False, upstream_ready=None, tau_out=None):
        share = 1.0 if self.band[1] == 0 else (self.band[1] - self.band[0]) / GRID_Y
        g = share * _view_grad(self.view, self.n)
        if accumulate:
            self.grad_flat[:self.n] += g
        else:
            self.grad_flat[:self.n] = g
        tau_out[:6] = share * _view_grad(100 + self.view, 6)
        self.calls += 1


def _window_worker(rank, world, port, V, n, whole_bands, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import KeyframeWindow

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _StandInEngine(n)
    cams = torch.zeros(V, 52)
    cams[:, 0] = torch.arange(V, dtype=torch.float32)
    win = KeyframeWindow(eng, cams, rank=rank, world_size=world, whole_bands=whole_bands)
    win.calibrate()
    up = (torch.zeros(V, 1), torch.zeros(V, 1))
    for _ in range(2):      # twice: nothing of the first iteration may leak into the second
        flat = win.iteration(up)
    assert eng.calls == 2 * len(win.units)
    np.save(os.path.join(out_dir, "w%d.npy" % rank), torch.cat([flat[:n], win.tau_all.reshape(-1)]).numpy())
    np.save(os.path.join(out_dir, "t%d.npy" % rank), win.tau.numpy())
    np.save(os.path.join(out_dir, "v%d.npy" % rank), np.asarray(win.views, np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,V,whole_bands", [(2, 5, 1), (2, 4, 2), (3, 7, 3)])
def test_keyframe_window_iteration_over_gloo(tmp_path, world, V, whole_bands):
    """The real KeyframeWindow (plan, per-unit launches, dL/dtau rows, the one all-reduce) over a stand-in engine on world_size-2 / 3
    gloo groups: every rank ends with the window gradient and every view's dL/dtau, also when a rank holds several bands of a view."""
    n = 1031
    mp.spawn(_window_worker, args=(world, _free_port(), V, n, whole_bands, str(tmp_path)), nprocs=world, join=True)
    want_g = sum(_view_grad(v, n) for v in range(V)).numpy()
    want_tau = np.zeros((V, 8), np.float32)
    for v in range(V):
        want_tau[v, :6] = _view_grad(100 + v, 6).numpy()
    for r in range(world):
        got = np.load(tmp_path / ("w%d.npy" % r))
        np.testing.assert_allclose(got[:n], want_g, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(got[n:].reshape(V, 6), want_tau[:, :6], rtol=1e-5, atol=1e-6)
        views = np.load(tmp_path / ("v%d.npy" % r))
        np.testing.assert_allclose(np.load(tmp_path / ("t%d.npy" % r)), want_tau[views, :6], rtol=1e-5, atol=1e-6)


def _densify_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import allreduce_densification_stats, shard_views

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, V = 257, 5
    accum, denom, radii = torch.zeros(P), torch.zeros(P), torch.zeros(P)
    for v in shard_views(V, world, rank):      # what the backward epilogue of every local view does (gaussian_model.py:767-771)
        g = _view_grad(v, P).abs()
        vis = _view_grad(50 + v, P) > 0
        accum[vis] += g[vis]
        denom[vis] += 1
        radii[vis] = torch.maximum(radii[vis], (10 * _view_grad(70 + v, P).abs())[vis])
    allreduce_densification_stats(accum, denom, radii)
    allreduce_densification_stats(None, None, None)      # nothing attached: nothing to do
    np.save(os.path.join(out_dir, "d%d.npy" % rank), torch.stack([accum, denom, radii]).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_densification_stats_reduce_over_the_ranks(tmp_path):
    mp.spawn(_densify_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    P, V = 257, 5
    accum, denom, radii = torch.zeros(P), torch.zeros(P), torch.zeros(P)
    for v in range(V):
        g, vis = _view_grad(v, P).abs(), _view_grad(50 + v, P) > 0
        accum[vis] += g[vis]
        denom[vis] += 1
        radii[vis] = torch.maximum(radii[vis], (10 * _view_grad(70 + v, P).abs())[vis])
    want = torch.stack([accum, denom, radii]).numpy()
    for r in range(2):
        np.testing.assert_allclose(np.load(tmp_path / ("d%d.npy" % r)), want, rtol=1e-6, atol=1e-6)
