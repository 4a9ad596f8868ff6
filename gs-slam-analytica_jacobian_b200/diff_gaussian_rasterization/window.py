"""Keyframe-window mapping sharded by keyframe across GPUs (BASELINE.json north_star).

One mapping iteration of the reference (utils/slam_backend.py:168-232) renders every keyframe of the
window through the same Gaussians, sums the losses and back-propagates once: autograd adds the V per-view
gradients of every Gaussian parameter.  The views are independent given the (replicated) map, so the path
shards by view:

  * rank r owns views r, r + world, r + 2 world, ...   (round robin; V=10 on 8 GPUs -> 2/2/1/1/1/1/1/1)
  * each rank runs forward + backward for its views on its own RasterEngine; the first local view
    overwrites the flat per-Gaussian gradient buffer, the following ones accumulate into it in the
    backward kernel itself (gsr_scene.accumulate_grads), so no extra add / memset passes;
  * ONE all-reduce(sum) of that flat fp32 buffer per iteration (NCCL over NVLink / NVSwitch when the
    process group is NCCL; gloo works for CPU tests of the host logic) gives every rank the window gradient;
  * per-view results (dL/dtau = pose gradient, radii, n_touched, dL/dmeans2D, images) stay on the owning
    rank -- each pose lives on one GPU, exactly like the reference keeps them per viewpoint
    (utils/slam_backend.py:195-198,236-240,277-284).
Candidate-pose batches for tracking (C3) use the same sharding with reduce=False: no collective at all.
"""
import torch


def shard_views(num_views, world_size, rank):
    """Round-robin owner map: the views rank `rank` renders."""
    return list(range(rank, num_views, world_size))


def owner_of(view, world_size):
    return view % world_size


def allreduce_window_gradients(grad_flat, group=None):
    """Sum the packed per-Gaussian gradient buffer over all ranks, in place (one collective per window
    iteration).  No-op without an initialised process group (single GPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=group)
    return grad_flat


class KeyframeWindow:
    """Runs the views a rank owns through a RasterEngine and reduces the window gradient.

    cameras: [V, 52] packed camera blocks (RasterEngine.pack_camera) resident on the engine's device.
    upstream(v) -> (dL_dcolor[3,H,W], dL_ddepth[1,H,W]) device tensors for view v, called AFTER view v's
    forward so it may depend on engine.color / engine.depth (the loss gradient); or a pair of
    [V,3,H,W] / [V,1,H,W] tensors.

    extra_engines: further RasterEngines over the SAME Gaussians (own workspaces and output buffers).  The local views
    are then dealt round robin to the engines, each on its own CUDA stream, so that the latency-bound stages of one view
    (per-Gaussian kernels, scatter, sort) overlap the issue-bound compositing of another; every engine accumulates into
    its own flat gradient buffer and the buffers are summed into the first engine's before the all-reduce.  With several
    engines `upstream` is called as upstream(v, engine).
    """

    def __init__(self, engine, cameras, rank=0, world_size=1, group=None, extra_engines=()):
        self.engine, self.cameras = engine, cameras
        self.engines = [engine] + list(extra_engines)
        self.rank, self.world, self.group = rank, world_size, group
        self.views = shard_views(int(cameras.shape[0]), world_size, rank)
        V = len(self.views)
        dev = engine.dev
        self.tau = torch.zeros((V, 6), dtype=torch.float32, device=dev)             # per local view [rho, theta]
        self.num_rendered = [0] * V
        self.streams = [torch.cuda.Stream(dev) for _ in self.engines] if len(self.engines) > 1 else None

    def calibrate(self):
        """Size the binning workspace for the largest local view (one exact plan per view)."""
        for e in self.engines:
            for v in self.views:
                e.set_camera(self.cameras[v])
                e.calibrate()

    def iteration(self, upstream, reduce=True, on_view=None, upstream_precomputed=False, fused_loss=None):
        """One window iteration.  Returns engine.grad_flat (summed over all views of all ranks when
        reduce=True).  on_view(local_index, view) can read the engine's per-view outputs.
        upstream_precomputed=True: the upstream gradients do not depend on this iteration's renders and upstream() launches
        nothing -- every view's compositing backward then starts tile by tile behind its forward
        (RasterEngine.launch_backward(overlap_forward=True)).
        fused_loss: callable view -> _cabi.GsrFusedLoss writing into the engine's upstream buffers (slam_ops.FusedLoss): the
        loss of every view is evaluated inside its forward, `upstream` is not called, the backward overlaps the forward
        (single-engine windows)."""
        eng = self.engine
        if not self.views:
            eng.grad_flat.zero_()        # a rank without views still takes part in the collective
        if self.streams is None:
            for i, v in enumerate(self.views):
                eng.set_camera(self.cameras[v])
                if fused_loss is not None:
                    eng.launch_forward(fused_loss=fused_loss(v))
                    eng.launch_backward(accumulate=(i > 0), overlap_forward=True)
                else:
                    eng.launch_forward()
                    if callable(upstream):
                        gc, gd = upstream(v)
                    else:
                        gc, gd = upstream[0][v], upstream[1][v]
                    eng.launch_backward(gc, gd, accumulate=(i > 0), overlap_forward=upstream_precomputed)
                self.tau[i].copy_(eng.g_tau, non_blocking=True)
                if on_view is not None:
                    on_view(i, v)
        else:
            main = torch.cuda.current_stream(eng.dev)
            used = [False] * len(self.engines)
            for st in self.streams:
                st.wait_stream(main)
            for i, v in enumerate(self.views):
                k = i % len(self.engines)
                e = self.engines[k]
                with torch.cuda.stream(self.streams[k]):
                    e.set_camera(self.cameras[v])
                    e.launch_forward()
                    if callable(upstream):
                        gc, gd = upstream(v, e)
                    else:
                        gc, gd = upstream[0][v], upstream[1][v]
                    e.launch_backward(gc, gd, accumulate=used[k], overlap_forward=upstream_precomputed)
                    used[k] = True
                    self.tau[i].copy_(e.g_tau, non_blocking=True)
                    if on_view is not None:
                        on_view(i, v)
            for st in self.streams:
                main.wait_stream(st)
            for k in range(1, len(self.engines)):
                if used[k]:
                    eng.grad_flat.add_(self.engines[k].grad_flat)
        if reduce:
            allreduce_window_gradients(eng.grad_flat, self.group)
        return eng.grad_flat
