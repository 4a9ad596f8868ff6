"""The callers on either side of the rasterizer in the SLAM loops, on the device:

  slam_loss(...)        the reference's tracking / mapping losses (utils/slam_utils.py:56-128) and their gradients
                        w.r.t. the rendered images and the exposure parameters in ONE kernel;
  tracking_step(...)    torch.optim.Adam on [cam_rot_delta, cam_trans_delta, exposure_a, exposure_b]
                        (utils/slam_frontend.py:129-162) + update_pose (utils/pose_utils.py:76-93) + the camera tensors
                        of the next render (utils/camera_utils.py:96-109) in ONE kernel;
  TrackingLoop          render -> loss -> backward -> optimiser -> pose update captured as ONE CUDA graph per iteration
                        on top of a RasterEngine: the reference's tracking loop (utils/slam_frontend.py:163-192) without a
                        host round trip per iteration; convergence is polled every `check_every` iterations.
All compute goes through the C-ABI (include/gsr_b200.h); torch owns memory and streams.
"""
import ctypes as C

import torch

from . import _cabi

_L = _cabi.load()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class LossWorkspace:
    """Reusable outputs + scratch of slam_loss for one image size."""

    def __init__(self, W, H, device="cuda"):
        dev = torch.device(device)
        self.W, self.H, self.dev = int(W), int(H), dev
        self.dL_dcolor = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        self.dL_ddepth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        self.sums = torch.zeros(4, dtype=torch.float32, device=dev)      # loss, dL/da, dL/db, 0
        self.scratch = torch.zeros(_L.gsr_slam_loss_scratch_bytes(W, H), dtype=torch.uint8, device=dev)


def slam_loss(ws, color, depth, opacity, gt_color, gt_depth=None, grad_mask=None, exposure=None, rgb_boundary_threshold=0.01,
              alpha=0.95, tracking=True, dL_dcolor=None, dL_ddepth=None):
    """One kernel: loss + dL/dcolor + dL/ddepth + dL/dexposure.  tracking=True: get_loss_tracking (opacity-weighted,
    grad_mask); False: get_loss_mapping.  gt_depth=None selects the monocular variants.  exposure: device tensor (a, b) or
    None.  Gradients land in ws.dL_dcolor / ws.dL_ddepth unless explicit outputs are given; returns ws.sums (device)."""
    gc = ws.dL_dcolor if dL_dcolor is None else dL_dcolor
    gd = ws.dL_ddepth if dL_ddepth is None else dL_ddepth
    with torch.cuda.device(ws.dev):
        _cabi.check(_L.gsr_slam_loss(ws.W, ws.H, _p(color), _p(depth), _p(opacity), _p(gt_color), _p(gt_depth), _p(grad_mask),
                                     _p(exposure), float(rgb_boundary_threshold), float(alpha), 0 if gt_depth is None else 1,
                                     1 if tracking else 0, _p(gc), _p(gd), _p(ws.sums), _p(ws.scratch), _stream(ws.dev)), "slam_loss")
    return ws.sums


class FusedLoss:
    """The same loss evaluated inside the forward compositing kernel's epilogue (gsr_fused_loss): build once per
    (workspace, targets), hand `.struct` to RasterEngine.launch_forward(fused_loss=...).  The gradients land in the given
    dL_dcolor / dL_ddepth (normally the engine's own upstream buffers), the sums in ws.sums."""

    def __init__(self, ws, gt_color, gt_depth=None, grad_mask=None, exposure=None, rgb_boundary_threshold=0.01, alpha=0.95,
                 tracking=True, dL_dcolor=None, dL_ddepth=None):
        gc = ws.dL_dcolor if dL_dcolor is None else dL_dcolor
        gd = ws.dL_ddepth if dL_ddepth is None else dL_ddepth
        self.scratch = torch.zeros(_L.gsr_fused_loss_scratch_bytes(ws.W, ws.H), dtype=torch.uint8, device=ws.dev)
        self._keep = (ws, gt_color, gt_depth, grad_mask, exposure, gc, gd)
        ptr = lambda t: None if t is None else t.data_ptr()
        self.struct = _cabi.GsrFusedLoss(ptr(gt_color), ptr(gt_depth), ptr(grad_mask), ptr(exposure), float(rgb_boundary_threshold),
                                         float(alpha), 0 if gt_depth is None else 1, 1 if tracking else 0, ptr(gc), ptr(gd),
                                         ptr(ws.sums), ptr(self.scratch))


class PoseState:
    """World-to-camera pose (R row-major, T), exposure (a, b), Adam state and status of one tracked frame, on the device."""

    def __init__(self, R, T, proj_raw, device="cuda", exposure=(0.0, 0.0)):
        dev = torch.device(device)
        self.dev = dev
        self.RT = torch.cat([torch.as_tensor(R, dtype=torch.float32).reshape(9), torch.as_tensor(T, dtype=torch.float32).reshape(3)]).to(dev)
        self.proj_raw = torch.as_tensor(proj_raw, dtype=torch.float32).reshape(16).to(dev).contiguous()
        self.exposure = torch.tensor(exposure, dtype=torch.float32, device=dev)
        self.adam = torch.zeros(17, dtype=torch.float32, device=dev)
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)     # converged, iterations, first converged iteration

    def reset_optimizer(self):
        self.adam.zero_()
        self.status.zero_()


def tracking_step(pose, dL_dtau, dL_dexposure, camera_block, lr_rot=0.003, lr_trans=0.001, lr_exposure=0.01,
                  converged_threshold=1e-4):
    """Adam + update_pose + next camera block, one kernel.  dL_dtau: the rasterizer's [rho, theta] gradient;
    dL_dexposure: the `sums` tensor of slam_loss (or None)."""
    with torch.cuda.device(pose.dev):
        _cabi.check(_L.gsr_tracking_step(_p(dL_dtau), _p(dL_dexposure), _p(pose.exposure), _p(pose.adam), _p(pose.RT), _p(pose.proj_raw),
                                         _p(camera_block), _p(pose.status), float(lr_rot), float(lr_trans), float(lr_exposure),
                                         float(converged_threshold), _stream(pose.dev)), "tracking_step")


def camera_block_from_pose(pose, camera_block):
    """Fill the engine's camera block from the current R, T without moving the pose (zero gradients, fresh Adam state)."""
    z6 = torch.zeros(6, dtype=torch.float32, device=pose.dev)
    adam, status, expo = pose.adam.clone(), pose.status.clone(), pose.exposure.clone()
    pose.status.zero_()      # (a frame that has converged ignores further steps)
    tracking_step(pose, z6, None, camera_block, 0.0, 0.0, 0.0, -1.0)
    pose.adam.copy_(adam); pose.status.copy_(status); pose.exposure.copy_(expo)


class TrackingLoop:
    """The reference's per-frame tracking loop (utils/slam_frontend.py:129-192) as one CUDA graph per iteration."""

    def __init__(self, engine, pose, gt_color, gt_depth=None, grad_mask=None, rgb_boundary_threshold=0.01, alpha=0.95,
                 lr_rot=0.003, lr_trans=0.001, lr_exposure=0.01, converged_threshold=1e-4, fused=True):
        """fused=True: the loss is evaluated in the forward compositing kernel's epilogue and the backward starts tile by
        tile behind the forward (three kernels + the pose step per iteration); False: the stand-alone loss kernel."""
        self.eng, self.pose = engine, pose
        self.gt_color, self.gt_depth, self.grad_mask = gt_color, gt_depth, grad_mask
        self.ws = LossWorkspace(engine.W, engine.H, engine.dev)
        self.kw = dict(rgb_boundary_threshold=rgb_boundary_threshold, alpha=alpha)
        self.lrs = (lr_rot, lr_trans, lr_exposure, converged_threshold)
        self.graph = None
        self.fused = FusedLoss(self.ws, gt_color, gt_depth, grad_mask, pose.exposure, tracking=True, dL_dcolor=engine.dL_dcolor,
                               dL_ddepth=engine.dL_ddepth, **self.kw) if fused else None
        camera_block_from_pose(pose, engine.cam)

    def _iteration(self):
        eng = self.eng
        if self.fused is not None:
            eng.launch_forward(fused_loss=self.fused.struct)
            eng.launch_backward(overlap_forward=True)
        else:
            eng.launch_forward()
            slam_loss(self.ws, eng.color, eng.depth, eng.opacity, self.gt_color, self.gt_depth, self.grad_mask, self.pose.exposure,
                      tracking=True, dL_dcolor=eng.dL_dcolor, dL_ddepth=eng.dL_ddepth, **self.kw)
            eng.launch_backward()
        tracking_step(self.pose, eng.g_tau, self.ws.sums, eng.cam, *self.lrs)

    def capture(self):
        eng = self.eng
        if eng.binning is None:
            eng.calibrate()
        # warm-up on a side stream with the state restored afterwards (the warm-up moves the pose)
        keep = [t.clone() for t in (self.pose.RT, self.pose.exposure, self.pose.adam, self.pose.status, eng.cam)]
        side = torch.cuda.Stream(eng.dev)
        side.wait_stream(torch.cuda.current_stream(eng.dev))
        with torch.cuda.stream(side):
            self._iteration()
        torch.cuda.current_stream(eng.dev).wait_stream(side)
        torch.cuda.synchronize(eng.dev)
        for dst, src in zip((self.pose.RT, self.pose.exposure, self.pose.adam, self.pose.status, eng.cam), keep):
            dst.copy_(src)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        for dst, src in zip((self.pose.RT, self.pose.exposure, self.pose.adam, self.pose.status, eng.cam), keep):
            dst.copy_(src)

    def run(self, max_iters=100, check_every=10, use_graph=True):
        """Iterate until converged (polled every `check_every` iterations like the reference's GUI cadence) or max_iters.
        Returns (iterations run, first converged iteration or 0, overflow flag of the last forward)."""
        if use_graph and self.graph is None:
            self.capture()
        done = 0
        while done < max_iters:
            n = min(check_every, max_iters - done)
            for _ in range(n):
                if use_graph:
                    self.graph.replay()
                else:
                    self._iteration()
            done += n
            st = self.pose.status.cpu()
            if int(st[2]) != 0:
                break
        _, overflow = self.eng.header()
        st = self.pose.status.cpu()
        return int(st[1]), int(st[2]), overflow


class MappingWindow:
    """One mapping iteration over a keyframe window (utils/slam_backend.py:156-232) with the loss of every view computed by
    the fused kernel right behind its forward: render -> get_loss_mapping + gradients -> backward, views sharded over the
    ranks by a KeyframeWindow, per-Gaussian gradients accumulated in the backward kernel and summed by one all-reduce.
    Per view (kept on the owning rank, like the reference keeps them per viewpoint): loss, dL/dexposure (a, b), dL/dtau,
    and -- through on_view -- radii / n_touched / dL/dmeans2D of the engine.
    The isotropic scale regulariser (10 * |s - mean(s)|.mean(), :228-230) is a per-Gaussian torch expression outside the
    rasterizer path; callers add its gradient to the scale part of the flat buffer."""

    def __init__(self, window, gt_colors, gt_depths=None, exposures=None, rgb_boundary_threshold=0.01, alpha=0.95, initialization=False,
                 fused=None):
        """gt_colors [V,3,H,W], gt_depths [V,1,H,W] or None (monocular), exposures [V,2] device tensor or None.
        fused (default): the loss of every unit is evaluated in its forward's epilogue; a unit that is a band of tile rows
        yields the band's share of the view's loss / dL/dexposure (the shares of a view add up)."""
        self.win, self.eng = window, window.engine
        self.gt_colors, self.gt_depths, self.exposures = gt_colors, gt_depths, exposures
        self.kw = dict(rgb_boundary_threshold=rgb_boundary_threshold, alpha=alpha)
        self.initialization = initialization
        self.ws = LossWorkspace(self.eng.W, self.eng.H, self.eng.dev)
        n = max(len(window.views), 1)
        self.view_sums = torch.zeros((n, 4), dtype=torch.float32, device=self.eng.dev)    # per local unit: loss, dL/da, dL/db, 0
        self.fused = None
        if fused is None:
            fused = True
        if fused:
            self.fused = []
            for i, v in enumerate(window.views):
                expo = None if (initialization or exposures is None) else exposures[v]
                e = window.engine_of(i)

                class _Slot:      # a LossWorkspace whose sums are this unit's row and whose gradients are its engine's buffers
                    pass
                slot = _Slot()
                slot.W, slot.H, slot.dev, slot.sums = e.W, e.H, e.dev, self.view_sums[i]
                slot.dL_dcolor, slot.dL_ddepth = e.dL_dcolor, e.dL_ddepth
                self.fused.append(FusedLoss(slot, gt_colors[v], None if gt_depths is None else gt_depths[v], None, expo, tracking=False,
                                            **self.kw))

    def capture(self, reduce=True):
        """The iteration as ONE CUDA graph (KeyframeWindow.capture); results land where iteration() leaves them
        (window.engine.grad_flat, self.view_sums, window.tau_all).  Run iteration() once before."""
        dev = self.eng.dev
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch(reduce, None)
        return g

    def iteration(self, reduce=True, on_view=None):
        """Returns (grad_flat summed over all views of all ranks, per-local-view sums [n,4], per-local-view dL/dtau [n,6])."""
        flat = self._launch(reduce, on_view)
        return flat, self.view_sums, self.win.tau

    def _launch(self, reduce, on_view):
        eng, local = self.eng, {v: i for i, v in enumerate(self.win.views)}

        def upstream(v, e=None):
            # (stand-alone loss kernel: whole views only -- it reads every pixel of the rendered images)
            e = eng if e is None else e
            expo = None if (self.initialization or self.exposures is None) else self.exposures[v]
            slam_loss(self.ws, e.color, e.depth, e.opacity, self.gt_colors[v], None if self.gt_depths is None else self.gt_depths[v],
                      None, expo, tracking=False, dL_dcolor=e.dL_dcolor, dL_ddepth=e.dL_ddepth, **self.kw)
            self.view_sums[local[v]].copy_(self.ws.sums, non_blocking=True)
            return e.dL_dcolor, e.dL_ddepth

        if self.fused is None:
            assert all(y1 == 0 for (_, _, y1) in self.win.units) and len(self.win.engines) == 1, \
                "the stand-alone loss kernel needs whole views on one engine; use fused=True"
        return self.win.iteration(upstream, reduce=reduce, on_view=on_view,
                                  fused_loss=None if self.fused is None else (lambda i, v: self.fused[i].struct))
