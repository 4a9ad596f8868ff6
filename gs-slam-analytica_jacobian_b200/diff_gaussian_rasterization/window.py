"""Keyframe-window mapping sharded by keyframe across GPUs (BASELINE.json north_star).

One mapping iteration of the reference (utils/slam_backend.py:168-232) renders every keyframe of the
window through the same Gaussians, sums the losses and back-propagates once: autograd adds the V per-view
gradients of every Gaussian parameter.  The views are independent given the (replicated) map, so the path
shards by view:

  * the work is cut into UNITS (view, tile_row_begin, tile_row_end).  floor(V / world) whole views go to every
    rank round robin (rank r owns views r, r + world, ...); the V mod world views left over are not handed to a
    few ranks whole (10 keyframes on 8 GPUs: two ranks with two views -> 5x at best) but split into bands of tile
    rows (gsr_scene.tile_row_begin / tile_row_end), one band per rank (10 on 8: every rank renders one whole view
    plus a quarter of a ninth / tenth view).  The bands of a view sum to the view: per-Gaussian gradients, dL/dtau
    and n_touched are linear in the pixels;
  * each rank runs forward + backward of its units on its RasterEngine(s); every unit ADDS its per-Gaussian
    gradients into ONE flat buffer in the backward kernel itself (gsr_scene.accumulate_grads) -- read-modify-write
    on one stream, REDs when several engines on as many streams share the buffer -- and writes its dL/dtau into
    row `view` of the [V, 8] block at the tail of that buffer;
  * ONE all-reduce(sum) of the flat buffer per iteration gives every rank the window gradient and every view's dL/dtau:
    the library's own kernel over NVSwitch multicast memory when the buffer lives in a symmetric allocation
    (SwitchReducer: multimem.ld_reduce + multimem.st, csrc/window_reduce.cu), else dist.all_reduce (NCCL; gloo for the
    CPU tests of the host logic);
  * per-unit results (radii, n_touched, dL/dmeans2D, images) stay on the owning rank -- like the reference keeps
    them per viewpoint (utils/slam_backend.py:195-198,236-240,277-284).
Candidate-pose batches for tracking (C3) use the same sharding with reduce=False and no bands: no collective at all.
"""
import math

import torch


def shard_views(num_views, world_size, rank):
    """Round-robin owner map of WHOLE views: the views rank `rank` renders when nothing is split."""
    return list(range(rank, num_views, world_size))


def owner_of(view, world_size):
    return view % world_size


def band_rows(grid_y, parts, weights=None):
    """Cut grid_y tile rows into `parts` consecutive bands [(begin, end), ...]: equal row counts, or -- weights = work per
    tile row (e.g. instances per row from a calibration pass) -- equal work.  Every band holds at least one row while
    grid_y >= parts; bands may be empty (begin == end) beyond that."""
    parts = int(parts)
    if weights is None:
        cuts = [(grid_y * k) // parts for k in range(parts + 1)]
    else:
        w = [max(float(x), 0.0) for x in weights]
        assert len(w) == grid_y
        total = sum(w) or 1.0
        cuts, acc, k = [0], 0.0, 1
        for y in range(grid_y):
            acc += w[y]
            while k < parts and acc >= total * k / parts and len(cuts) <= k:
                cuts.append(y + 1)
                k += 1
        while len(cuts) < parts:
            cuts.append(grid_y)
        cuts.append(grid_y)
        if grid_y >= parts:        # no empty band: push cuts apart
            for k in range(1, parts):
                cuts[k] = min(max(cuts[k], cuts[k - 1] + 1), grid_y - (parts - k))
    return [(cuts[k], cuts[k + 1]) for k in range(parts)]


def plan_units(num_views, world_size, grid_y, split=True, row_weights=None, whole_bands=1):
    """The units (view, tile_row_begin, tile_row_end) of every rank: list (per rank) of lists.  (view, 0, 0) = whole view.
    split=False reproduces shard_views.  row_weights: optional {view: work per tile row} for equal-work bands.
    whole_bands > 1: a rank's whole views are themselves cut into that many bands (more units in flight on a rank that
    holds a single view: the latency-bound per-Gaussian stages of one band overlap the compositing of another, at the price
    of running the per-Gaussian kernels once per band)."""
    V, N = int(num_views), int(world_size)
    units = [[] for _ in range(N)]
    whole = V if not split else (V // N) * N
    wb = max(1, min(int(whole_bands), int(grid_y)))
    for v in range(whole):
        if wb == 1:
            units[v % N].append((v, 0, 0))
        else:
            units[v % N].extend((v, y0, y1) for (y0, y1) in band_rows(grid_y, wb, None if row_weights is None else row_weights.get(v)))
    rem = V - whole
    if rem:
        b = N // math.gcd(rem, N)                 # bands per left-over view: rem * b units, rem / gcd per rank
        b = min(b, max(int(grid_y), 1))
        j = 0
        for v in range(whole, V):
            bands = band_rows(grid_y, b, None if row_weights is None else row_weights.get(v))
            for (y0, y1) in bands:
                if y1 > y0:
                    units[j % N].append((v, y0, y1) if b > 1 else (v, 0, 0))
                j += 1
    return units


def tau_rows(units, num_views):
    """Row of the [tau_slots, 8] pose-gradient block every local unit writes its dL/dtau into (the backward kernel STORES its
    six sums): the first unit of a view on this rank takes row `view`, every further band of the same view (whole_bands > 1)
    a spare row behind the window's views; `merges` = [(spare row, view), ...] are added into the view's row before the
    reduction.  Returns (rows, merges)."""
    rows, merges, seen, spare = [], [], set(), int(num_views)
    for (v, _, _) in units:
        if v in seen:
            rows.append(spare)
            merges.append((spare, v))
            spare += 1
        else:
            seen.add(v)
            rows.append(v)
    return rows, merges


def allreduce_window_gradients(grad_flat, group=None):
    """Sum the packed per-Gaussian gradient buffer over all ranks, in place (one collective per window
    iteration).  No-op without an initialised process group (single GPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=group)
    return grad_flat


def allreduce_densification_stats(xyz_gradient_accum=None, denom=None, max_radii2D=None, group=None):
    """The densification statistics of a keyframe-sharded window: every rank's engines accumulate
    xyz_gradient_accum += ||dL/dmeans2D||, denom += 1 and max_radii2D = max(., radii) for the views THEY render
    (RasterEngine.attach_densification_stats; gaussian_model.py:767-771, utils/slam_backend.py:115-121), so a replicated optimiser
    needs the sum / sum / maximum over the ranks before it densifies -- every `gaussian_update_every` iterations, not per
    iteration.  Call with per-rank DELTAS for the two sums (tensors zeroed at the last call) or reset them afterwards.
    The statistics are per whole view (a norm is not linear in the pixels): run the iterations that feed them with
    KeyframeWindow(split=False).  In place; no-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    for t, op in ((xyz_gradient_accum, dist.ReduceOp.SUM), (denom, dist.ReduceOp.SUM), (max_radii2D, dist.ReduceOp.MAX)):
        if t is not None:
            dist.all_reduce(t, op=op, group=group)


class SwitchReducer:
    """The window's gradient buffer in a symmetric allocation + its sum over the ranks by ONE kernel over NVSwitch multicast
    memory (gsr_window_allreduce: multimem.ld_reduce / multimem.st, no NCCL on the data path).

        red = SwitchReducer.create(n_floats, device, group)      # collective; None when multicast memory is unavailable
        eng = RasterEngine(..., grad_flat=red.buffer)            # the engine accumulates straight into the symmetric buffer
        win = KeyframeWindow(eng, cams, rank, world, reducer=red)

    torch.distributed._symmetric_memory is used for the plumbing only (allocation, handle exchange, multicast mapping,
    signal pads); the reduction itself is the library's kernel."""
    last_error = None

    def __init__(self, buffer, padded, handle, ctas):
        import ctypes as C

        from . import _cabi

        self._C, self._cabi, self._L = C, _cabi, _cabi.load()
        self.buffer, self.padded, self.handle, self.ctas = buffer, padded, handle, int(ctas)
        self.rank, self.world = int(handle.rank), int(handle.world_size)
        self.dev = buffer.device
        self.mc = int(handle.multicast_ptr) + (padded.data_ptr() - int(handle.buffer_ptrs[self.rank]))
        self.pads = int(handle.signal_pad_ptrs_dev)
        self.pad_bytes = int(handle.signal_pad_size)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)

    @staticmethod
    def create(n_floats, device, group=None, ctas=64):
        """Collective over `group`.  Returns None (on every rank) when symmetric / multicast memory cannot be set up."""
        import torch.distributed as dist

        try:
            import torch.distributed._symmetric_memory as symm

            group = dist.group.WORLD if group is None else group
            n_pad = (int(n_floats) + 3) // 4 * 4
            padded = symm.empty(n_pad, dtype=torch.float32, device=torch.device(device))
            handle = symm.rendezvous(padded, group)
            ok = bool(handle.multicast_ptr) and handle.world_size > 1
        except Exception as ex:      # no symmetric-memory support in this torch build / on this box: the caller falls back to NCCL
            ok, padded, handle = False, None, None
            SwitchReducer.last_error = "%s: %s" % (type(ex).__name__, ex)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device(device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)          # all ranks or none
        if int(flag.item()) == 0:
            return None
        padded.zero_()
        return SwitchReducer(padded[:int(n_floats)], padded, handle, ctas)

    def all_reduce(self):
        """In-place sum of the buffer over the ranks on the current stream (a collective: every rank calls it)."""
        C = self._C
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            self._cabi.check(self._L.gsr_window_allreduce(C.c_void_p(self.mc), C.c_void_p(self.pads), self.rank, self.world,
                                                          self.padded.numel(), self.ctas, self.pad_bytes,
                                                          C.c_void_p(self.status.data_ptr()), st), "window_allreduce")
        return self.buffer

    def timed_out(self):
        """True when a peer did not arrive in some all_reduce since the last call (synchronises)."""
        bad = bool(self.status.item())
        self.status.zero_()
        return bad


class KeyframeWindow:
    """Runs the units a rank owns through its RasterEngine(s) and reduces the window gradient.

    cameras: [V, 52] packed camera blocks (RasterEngine.pack_camera) resident on the engine's device.
    upstream(v) -> (dL_dcolor[3,H,W], dL_ddepth[1,H,W]) device tensors for view v, called AFTER view v's
    forward so it may depend on engine.color / engine.depth (the loss gradient); or a pair of
    [V,3,H,W] / [V,1,H,W] tensors.  For a band unit only the band's pixel rows of the engine's images are valid and
    only the band's rows of the upstream gradients are read.

    extra_engines: further RasterEngines over the SAME Gaussians, constructed with grad_flat=engine.grad_flat (own
    workspaces and output buffers, ONE gradient buffer).  The local units are then dealt round robin to the engines, each
    on its own CUDA stream, so that the latency-bound stages of one view (per-Gaussian kernels, scatter) overlap the
    issue-bound compositing of another; all of them add into the shared buffer with REDs.  With several engines
    `upstream` is called as upstream(v, engine).

    split=True (default): views left over after every rank has floor(V / world) whole ones are split into bands of tile
    rows over the ranks (plan_units).  `self.views` lists the views of this rank's units, `self.units` the units.
    """

    def __init__(self, engine, cameras, rank=0, world_size=1, group=None, extra_engines=(), split=True, row_weights=None,
                 reducer=None, whole_bands=1):
        """reducer: a SwitchReducer whose buffer IS engine.grad_flat -- the window gradient is then summed by the library's
        NVSwitch kernel instead of dist.all_reduce."""
        self.engine, self.cameras = engine, cameras
        assert reducer is None or reducer.buffer.data_ptr() == engine.grad_flat.data_ptr(), "the reducer's buffer must be the engine's grad_flat"
        self.reducer = reducer
        self.engines = [engine] + list(extra_engines)
        for e in self.engines[1:]:
            assert e.grad_flat.data_ptr() == engine.grad_flat.data_ptr(), "extra engines must share the first engine's grad_flat"
        self.rank, self.world, self.group = rank, world_size, group
        self.num_views = int(cameras.shape[0])
        assert self.num_views <= engine.tau_slots, "more views than tau_slots of the engine"
        grid_y = (engine.H + 15) // 16
        self.plan = plan_units(self.num_views, world_size, grid_y, split=split, row_weights=row_weights, whole_bands=whole_bands)
        self.units = self.plan[rank]
        self.views = [u[0] for u in self.units]
        n = len(self.units)
        self.tau_row, self.tau_merges = tau_rows(self.units, self.num_views)
        assert self.num_views + len(self.tau_merges) <= engine.tau_slots, "tau_slots too small for the bands of this rank"
        self.num_rendered = [0] * n
        self.streams = [torch.cuda.Stream(engine.dev) for _ in self.engines] if len(self.engines) > 1 else None
        # dL/dtau of EVERY view of the window (complete after the all-reduce; rows of other ranks' whole views are theirs)
        self.tau_all = engine.tau_block[:self.num_views, :6]

    def engine_of(self, local_unit):
        """The engine that runs local unit i (units are dealt round robin to the engines)."""
        return self.engines[local_unit % len(self.engines)]

    @property
    def tau(self):
        """[local units, 6]: dL/dtau = [rho, theta] of the views of this rank's units (a view split into bands holds the
        sum over its bands once the iteration has been reduced)."""
        return self.tau_all[self.views] if self.views else self.tau_all[:0]

    def calibrate(self):
        """Size the binning workspaces for the largest local unit (one exact plan per unit and engine)."""
        for e in self.engines:
            for i, (v, y0, y1) in enumerate(self.units):
                e.set_camera(self.cameras[v])
                e.set_band(y0, y1)
                e.use_order(i)
                e.calibrate(build_order=e is self.engine_of(i))      # the scatter's spatial order: on the engine that runs the unit

    def capture(self, upstream, **kw):
        """One window iteration (same arguments as iteration()) as ONE CUDA graph over the persistent buffers: every unit's
        launches on every engine stream and the gradient reduction.  Returns the torch.cuda.CUDAGraph; replay() it instead
        of calling iteration() -- a dozen launches and stream fork / joins per unit cost the host more than a short
        iteration costs the GPU (10 keyframes on 8 GPUs: 1.27 ms host-driven against 1.02 ms of device time).  The camera
        blocks are read from self.cameras at replay time; `upstream` tensors / fused-loss structures must stay alive.
        Every rank of the group must capture and replay in step (the reduction is a collective).  Call iteration() once
        before capturing (kernel attributes, lazy allocations)."""
        dev = self.engine.dev
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.iteration(upstream, **kw)
        return g

    def iteration(self, upstream, reduce=True, on_view=None, upstream_precomputed=False, fused_loss=None):
        """One window iteration.  Returns engine.grad_flat (summed over all units of all ranks when reduce=True; its tail
        holds every view's dL/dtau, self.tau_all).  on_view(local_index, view) can read the engine's per-unit outputs.
        upstream_precomputed=True: the upstream gradients do not depend on this iteration's renders and upstream() launches
        nothing -- every unit's compositing backward then starts tile by tile behind its forward
        (RasterEngine.launch_backward(overlap_forward=True)).
        fused_loss: callable (local unit index, view) -> _cabi.GsrFusedLoss writing into the upstream buffers of the engine
        that runs the unit (self.engine_of(i); slam_ops.FusedLoss): the loss of every unit is evaluated inside its forward,
        `upstream` is not called, the backward overlaps the forward."""
        eng = self.engine
        if self.streams is None:
            if not self.units:
                eng.grad_flat.zero_()        # a rank without units still takes part in the collective
            else:
                eng.tau_block.zero_()
            for i, (v, y0, y1) in enumerate(self.units):
                eng.set_camera(self.cameras[v])
                eng.set_band(y0, y1)
                eng.use_order(i)
                if fused_loss is not None:
                    eng.launch_forward(fused_loss=fused_loss(i, v))
                    eng.launch_backward(accumulate=(i > 0), overlap_forward=True, tau_out=eng.tau_block[self.tau_row[i]])
                else:
                    eng.launch_forward()
                    if callable(upstream):
                        gc, gd = upstream(v)
                    else:
                        gc, gd = upstream[0][v], upstream[1][v]
                    eng.launch_backward(gc, gd, accumulate=(i > 0), overlap_forward=upstream_precomputed, tau_out=eng.tau_block[self.tau_row[i]])
                if on_view is not None:
                    on_view(i, v)
        else:
            main = torch.cuda.current_stream(eng.dev)
            eng.grad_flat.zero_()            # every unit ADDS (REDs): the engines run concurrently into one buffer
            for st in self.streams:
                st.wait_stream(main)
            for i, (v, y0, y1) in enumerate(self.units):
                k = i % len(self.engines)
                e = self.engines[k]
                with torch.cuda.stream(self.streams[k]):
                    e.set_camera(self.cameras[v])
                    e.set_band(y0, y1)
                    e.use_order(i)
                    if fused_loss is not None:
                        e.launch_forward(fused_loss=fused_loss(i, v))
                        e.launch_backward(accumulate="atomic", overlap_forward=True, tau_out=eng.tau_block[self.tau_row[i]])
                    else:
                        e.launch_forward()
                        if callable(upstream):
                            gc, gd = upstream(v, e)
                        else:
                            gc, gd = upstream[0][v], upstream[1][v]
                        e.launch_backward(gc, gd, accumulate="atomic", overlap_forward=upstream_precomputed, tau_out=eng.tau_block[self.tau_row[i]])
                    if on_view is not None:
                        on_view(i, v)
            for st in self.streams:
                main.wait_stream(st)
        for (spare, v) in self.tau_merges:       # several bands of one view on this rank: their dL/dtau add up
            eng.tau_block[v] += eng.tau_block[spare]
        if reduce:
            if self.reducer is not None:
                self.reducer.all_reduce()
            else:
                allreduce_window_gradients(eng.grad_flat, self.group)
        return eng.grad_flat
