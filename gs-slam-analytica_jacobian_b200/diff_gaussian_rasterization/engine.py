"""RasterEngine: the launch-lean way to drive the hot path for SLAM loops.

The reference calls the rasterizer once per view through autograd, allocates every scratch/grad tensor
per call and blocks on a D2H read of num_rendered in every forward (rasterizer_impl.cu:331).  A tracking
loop (utils/slam_frontend.py:163, 100 iterations on one frozen map) or a mapping window iteration
(utils/slam_backend.py:168-232) repeats the same shapes, so the engine
  * owns persistent workspaces (geometry / binning / image) and output + gradient buffers,
  * runs forward and backward WITHOUT any host synchronisation: the binning capacity is calibrated once
    (exact run) with head-room, the kernels read num_rendered on the device, and an overflow flag is
    read back with the results (overflow -> the step is re-run exactly, never silently wrong),
  * captures forward(+backward) into CUDA graphs: one launch per step instead of ~14,
  * takes the camera as one 208-byte device block whose contents are replaced in place per step,
  * keeps all per-Gaussian parameter gradients in ONE flat buffer (`grad_flat`) so that a keyframe window
    can accumulate views in place (accumulate=True) and all-reduce a single tensor across GPUs.
All compute goes through the C-ABI (include/gsr_b200.h); torch only owns memory and streams.
"""
import ctypes as C
import os

import torch

from . import _cabi
from ._cabi import GsrScene

_L = _cabi.load()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _RawDeviceBytes:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class RasterEngine:
    def __init__(self, gaussians, image_width, image_height, tanfovx, tanfovy, bg, sh_degree=0, scale_modifier=1.0,
                 device="cuda", headroom=1.25, tau_slots=64, grad_flat=None):
        """gaussians: dict with means3D[P,3], opacities[P,1], shs[P,M,3] or colors_precomp[P,3],
        scales[P,3]+rotations[P,4] or cov3D_precomp[P,6] (fp32 tensors; moved to `device`).
        tau_slots: rows of the [tau_slots, 8] pose-gradient block kept at the tail of grad_flat (a keyframe window reduces
        the per-view dL/dtau of views split over ranks together with the per-Gaussian gradients).
        grad_flat: share the flat gradient buffer of another engine over the SAME Gaussians (views of a window running on
        several streams then add into one buffer: launch_backward(accumulate="atomic"))."""
        self.dev = torch.device(device)
        if self.dev.type == "cuda" and self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        f = lambda k: (gaussians[k].to(self.dev, torch.float32).contiguous() if gaussians.get(k) is not None else None)
        self.g = {k: f(k) for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp")}
        if self.g["colors_precomp"] is not None:
            self.g["shs"] = None
        if self.g["cov3D_precomp"] is not None:
            self.g["scales"] = self.g["rotations"] = None
        self.P = int(self.g["means3D"].shape[0])
        self.M = int(self.g["shs"].shape[1]) if self.g["shs"] is not None else 0
        self.W, self.H = int(image_width), int(image_height)
        self.headroom = float(headroom)
        f32 = dict(dtype=torch.float32, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.bg = torch.as_tensor(bg, dtype=torch.float32).to(self.dev).contiguous()
        # camera block: view(16) | proj(16) | proj_raw(16) | campos(4)  -- one 208-byte H2D per pose
        self.cam = torch.zeros(52, **f32)
        P, W, H, M = self.P, self.W, self.H, self.M
        self.color = torch.empty((3, H, W), **f32)
        self.depth = torch.empty((1, H, W), **f32)
        self.opacity = torch.empty((1, H, W), **f32)
        self.radii = torch.empty((P,), **i32)
        self.n_touched = torch.empty((P,), **i32)
        self.dL_dcolor = torch.zeros((3, H, W), **f32)
        self.dL_ddepth = torch.zeros((1, H, W), **f32)
        # flat per-Gaussian gradient buffer: [means3D 3P | colour part | opacity P | covariance part]
        ncol = 3 * M if self.g["shs"] is not None else 3
        ncov = 7 if self.g["scales"] is not None else 6
        self.tau_slots = int(tau_slots)
        n_flat = P * (3 + ncol + 1 + ncov) + 32 + 8 * self.tau_slots            # +32: keeps every segment 16-B aligned
        if grad_flat is not None:
            assert grad_flat.numel() == n_flat and grad_flat.device == self.dev and grad_flat.dtype == torch.float32
        self.grad_flat = torch.zeros(n_flat, **f32) if grad_flat is None else grad_flat
        self.tau_block = self.grad_flat[n_flat - 8 * self.tau_slots:].view(self.tau_slots, 8)   # row v: [rho, theta, 0, 0] of view v
        o = 0

        def take(n, shape):
            nonlocal o
            o = (o + 3) // 4 * 4
            v = self.grad_flat[o:o + n].view(shape)
            o += n
            return v

        self.g_means3D = take(3 * P, (P, 3))
        self.g_sh = take(3 * M * P, (P, M, 3)) if self.g["shs"] is not None else None
        self.g_colors = take(3 * P, (P, 3)) if self.g["colors_precomp"] is not None else None
        self.g_opacity = take(P, (P, 1))
        self.g_rot = take(4 * P, (P, 4)) if self.g["scales"] is not None else None
        self.g_scales = take(3 * P, (P, 3)) if self.g["scales"] is not None else None
        self.g_cov = take(6 * P, (P, 6)) if self.g["cov3D_precomp"] is not None else None
        self.g_means2D = torch.empty((P, 3), **f32)     # per view (densification statistic), not part of grad_flat
        self.g_tau = torch.zeros((6,), **f32)           # per view: [rho, theta]
        self.geom_bytes = _L.gsr_geometry_bytes(P, W, H)
        self.img_bytes = _L.gsr_image_bytes(W, H)
        self.geom = torch.empty((self.geom_bytes,), **u8)
        self.img = torch.empty((self.img_bytes,), **u8)
        self.capacity = 0
        self.max_tile_hint = 0
        self.bin_bytes = 0
        self.binning = None
        s = GsrScene()
        s.P, s.D, s.M, s.W, s.H = P, int(sh_degree), M, W, H
        s.background = self.bg.data_ptr()
        for k in ("means3D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp"):
            setattr(s, k, self.g[k].data_ptr() if self.g[k] is not None else None)
        base = self.cam.data_ptr()
        s.viewmatrix, s.projmatrix, s.projmatrix_raw, s.campos = base, base + 64, base + 128, base + 192
        s.scale_modifier, s.tan_fovx, s.tan_fovy = float(scale_modifier), float(tanfovx), float(tanfovy)
        s.prefiltered, s.debug, s.accumulate_grads, s.overlap_forward = 0, 0, 0, 0
        s.upstream_ready = None
        s.spatial_order = None
        s.depth_cut = None
        s.densify_grad_accum = s.densify_denom = s.max_radii2D = None
        self.scene = s
        self.graph_fwd = self.graph_bwd = self.graph_all = None
        self.last_num_rendered = None
        # screen-coherent processing orders of the scatter kernel (gsr_scene.spatial_order), one per key (a window unit);
        # built by calibrate() for maps too large for the cooperative preprocess + scatter kernel
        self.spatial_orders = {}
        self.depth_cuts = {}      # per key: the per-tile depth hints that go with the order (gsr_scene.depth_cut)
        self.depth_cut_min_list = 4096
        self.order_key = 0
        self.use_spatial_order = (not os.environ.get("GSR_NO_SPATIAL_ORDER")) and not _L.gsr_forward_nosync_fuses_scatter(P, W, H)
        # step(): the compositing backward overlaps the tail of the compositing forward (GSR_NO_OVERLAP=1: plain order)
        self.overlap = not os.environ.get("GSR_NO_OVERLAP")

    @staticmethod
    def flat_size(P, M=1, colors_precomp=False, cov3D_precomp=False, tau_slots=64):
        """Length (floats) of grad_flat for a model of P Gaussians with M SH coefficients: what a caller-provided buffer
        (e.g. a symmetric allocation, window.SwitchReducer) must hold."""
        ncol = 3 if colors_precomp else 3 * int(M)
        ncov = 6 if cov3D_precomp else 7
        return int(P) * (3 + ncol + 1 + ncov) + 32 + 8 * int(tau_slots)

    # ---- camera ---------------------------------------------------------------------------------------
    @staticmethod
    def pack_camera(viewmatrix, projmatrix, projmatrix_raw, campos, out=None):
        """Pack the four settings tensors (any device) into the 52-float camera block layout."""
        out = torch.empty(52, dtype=torch.float32) if out is None else out
        out[0:16] = viewmatrix.reshape(-1)
        out[16:32] = projmatrix.reshape(-1)
        out[32:48] = projmatrix_raw.reshape(-1)
        out[48:51] = campos.reshape(-1)
        out[51] = 0
        return out

    def set_camera(self, packed, non_blocking=True):
        self.cam.copy_(packed, non_blocking=non_blocking)

    def set_band(self, tile_row_begin=0, tile_row_end=0):
        """Render only the tile rows [begin, end) of the view from now on (gsr_scene.tile_row_begin / tile_row_end; 0, 0 =
        the whole image): Gaussians outside the band are culled (radii 0), pixels outside are not written, gradients and
        n_touched are the band's share.  The captured graphs belong to the previous band."""
        if (int(tile_row_begin), int(tile_row_end)) != (self.scene.tile_row_begin, self.scene.tile_row_end):
            self.scene.tile_row_begin, self.scene.tile_row_end = int(tile_row_begin), int(tile_row_end)
            self.graph_fwd = self.graph_bwd = self.graph_all = None

    def use_order(self, key):
        """Select the spatial order built for `key` (e.g. the index of a window unit: every view / band has its own); a key
        without an order yet runs the plain scatter until calibrate() builds one.  The captured graphs belong to the
        previous order."""
        self.order_key = key
        t = self.spatial_orders.get(key)
        ptr = None if t is None else t.data_ptr()
        c = self.depth_cuts.get(key) if t is not None else None
        cptr = None if c is None else c.data_ptr()
        if ptr != self.scene.spatial_order or cptr != self.scene.depth_cut:
            self.scene.spatial_order = ptr
            self.scene.depth_cut = cptr
            self.graph_fwd = self.graph_bwd = self.graph_all = None

    # ---- capacity -------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def ensure_capacity(self, num_rendered):
        need = int(num_rendered * self.headroom) + 4096
        if need > self.capacity:
            self.capacity = need
            self.bin_bytes = _L.gsr_binning_bytes(self.P, self.W, self.H, self.capacity)
            self.binning = torch.empty((self.bin_bytes,), dtype=torch.uint8, device=self.dev)
            self.graph_fwd = self.graph_bwd = self.graph_all = None   # pointers changed
        return self.capacity

    def calibrate(self, build_order=True):
        """One exact (synchronising) forward plan at the current camera to size the binning workspace (and, for maps beyond
        the cooperative preprocess + scatter kernel, to build the scatter's spatial order for the current order key)."""
        with torch.cuda.device(self.dev):
            _cabi.check(_L.gsr_forward_plan(C.byref(self.scene), _p(self.geom), self.geom_bytes, _p(self.radii),
                                            _p(self.n_touched), self._stream()), "forward_plan")
            R, mt = C.c_longlong(0), C.c_longlong(0)
            _cabi.check(_L.gsr_forward_num_rendered(_p(self.geom), self._stream(), C.byref(R), C.byref(mt)), "num_rendered")
        self.last_num_rendered = int(R.value)
        self.longest_list = max(getattr(self, "longest_list", 0), int(mt.value))
        hint = int(mt.value * self.headroom) + 64      # longest per-tile list -> smem capacity of the tile sort
        if hint > self.max_tile_hint:
            self.max_tile_hint = hint
            self.graph_fwd = self.graph_bwd = self.graph_all = None
        self.ensure_capacity(R.value)
        if self.use_spatial_order and build_order:
            # Gaussians bucketed by the tile at the centre of their rectangle AT THIS CAMERA; later poses reuse it (any
            # permutation gives the same lists, a stale one only loses locality)
            t = self.spatial_orders.get(self.order_key)
            if t is None:
                t = self.spatial_orders[self.order_key] = torch.empty((self.P,), dtype=torch.int32, device=self.dev)
            with torch.cuda.device(self.dev):
                _cabi.check(_L.gsr_spatial_order(C.byref(self.scene), _p(self.geom), self.geom_bytes, _p(t), self._stream()), "spatial_order")
            # per-tile depth hints (gsr_scene.depth_cut) pay where the forward sweeps lists too long for its registers (> 4096
            # entries: C4 view 2.41 -> 2.20 ms); at 2-3 k entries per tile the partition costs the scatter what it saves the
            # forward (C2: +0.2 %), so shorter lists run without
            if self.order_key not in self.depth_cuts and self.longest_list > self.depth_cut_min_list and not os.environ.get("GSR_NO_DEPTH_CUT"):
                tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
                self.depth_cuts[self.order_key] = torch.full((tiles,), 0x7f800000, dtype=torch.int32, device=self.dev)      # +inf: no cut yet
            self.use_order(self.order_key)
        return int(R.value)

    # ---- un-synchronised launches ---------------------------------------------------------------------
    def launch_forward(self, fused_loss=None):
        """fused_loss: a _cabi.GsrFusedLoss (slam_ops.FusedLoss.struct): the view's SLAM loss and its gradients w.r.t. the
        rendered images are evaluated in the epilogue of the compositing kernel (no kernel between forward and backward)."""
        self.scene.fused_loss = None if fused_loss is None else C.pointer(fused_loss)
        try:
            self._launch_forward()
        finally:
            self.scene.fused_loss = None

    def _launch_forward(self):
        _cabi.check(_L.gsr_forward_nosync(C.byref(self.scene), _p(self.geom), self.geom_bytes, _p(self.binning), self.bin_bytes,
                                          self.capacity, self.max_tile_hint, _p(self.img), self.img_bytes, _p(self.color), _p(self.depth),
                                          _p(self.opacity), _p(self.radii), _p(self.n_touched), self._stream()), "forward_nosync")

    def attach_densification_stats(self, xyz_gradient_accum=None, denom=None, max_radii2D=None):
        """fp32 [P] (or [P,1]) device tensors updated in the backward's epilogue for the visible Gaussians of each view:
        xyz_gradient_accum += ||dL/dmeans2D[:2]||, denom += 1, max_radii2D = max(., radii)
        (gaussian_model.py:767-771, utils/slam_backend.py:115-121).  None detaches."""
        for name, t in (("densify_grad_accum", xyz_gradient_accum), ("densify_denom", denom), ("max_radii2D", max_radii2D)):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.numel() == self.P and t.is_contiguous()
            setattr(self.scene, name, None if t is None else t.data_ptr())
        self._densify_refs = (xyz_gradient_accum, denom, max_radii2D)      # keep the tensors alive
        self.graph_fwd = self.graph_bwd = self.graph_all = None

    def launch_backward(self, dL_dcolor=None, dL_ddepth=None, accumulate=False, overlap_forward=False, upstream_ready=None,
                        tau_out=None):
        """dL_dcolor / dL_ddepth default to the engine's own buffers; accumulate=True adds this view's
        per-Gaussian gradients into grad_flat (mapping window) instead of overwriting; accumulate="atomic" does so with
        REDs (several engines on different streams sharing one pre-zeroed grad_flat).  tau_out: a 6-float device tensor
        that receives dL/dtau instead of g_tau (e.g. a row of tau_block).
        overlap_forward=True: ONLY directly behind launch_forward() on the same stream, with upstream gradients that were
        complete before it -- the compositing backward then starts tile by tile behind the compositing forward
        (programmatic dependent launch + per-tile flags, gsr_scene.overlap_forward) instead of waiting for its tail."""
        gc = self.dL_dcolor if dL_dcolor is None else dL_dcolor
        gd = self.dL_ddepth if dL_ddepth is None else dL_ddepth
        self.scene.accumulate_grads = 2 if accumulate == "atomic" else (1 if accumulate else 0)
        self.scene.overlap_forward = 1 if (overlap_forward and self.overlap) else 0
        self.scene.upstream_ready = None if upstream_ready is None else upstream_ready.data_ptr()
        try:
            _cabi.check(_L.gsr_rasterize_gaussians_backward(
                C.byref(self.scene), _p(self.radii), _p(self.geom), _p(self.binning), self.capacity, _p(self.img),
                _p(gc), _p(gd), _p(self.g_means3D), _p(self.g_means2D), _p(self.g_sh), _p(self.g_colors),
                _p(self.g_opacity), _p(self.g_scales), _p(self.g_rot), _p(self.g_cov),
                _p(self.g_tau if tau_out is None else tau_out), self._stream()), "backward")
        finally:
            self.scene.accumulate_grads = 0
            self.scene.overlap_forward = 0
            self.scene.upstream_ready = None

    def capture(self):
        """Capture forward, backward and forward+backward CUDA graphs over the persistent buffers."""
        if self.binning is None:
            self.calibrate()
        with torch.cuda.device(self.dev):
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self.launch_forward()
                self.launch_backward()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            self.graph_fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fwd):
                self.launch_forward()
            self.graph_bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_bwd):
                self.launch_backward()
            self.graph_all = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_all):
                self.launch_forward()
                self.launch_backward(overlap_forward=True)

    def capture_host_step(self, cam_host, dL_dcolor_host=None, dL_ddepth_host=None):
        """One CUDA graph for a step driven from HOST buffers: H2D of the pinned 52-float camera block (and, when given,
        of the pinned upstream gradients, on a forked branch that overlaps the forward), forward, backward, D2H of
        dL/dtau and the (num_rendered, overflow) header into pinned tensors.  The caller rewrites the pinned inputs in
        place and calls step_host(): two Python calls per step instead of ten."""
        for t in (cam_host, dL_dcolor_host, dL_ddepth_host):
            assert t is None or (t.is_pinned() and t.is_contiguous())
        if self.binning is None:
            self.calibrate()
        self.h_tau = torch.empty(6, dtype=torch.float32).pin_memory()
        self.h_hdr = torch.empty(2, dtype=torch.int32).pin_memory()
        hdr_dev = torch.as_tensor(_RawDeviceBytes(self.geom.data_ptr(), 8), device=self.dev).view(torch.int32)
        self._host_refs = (cam_host, dL_dcolor_host, dL_ddepth_host, hdr_dev)
        with torch.cuda.device(self.dev):
            warm = torch.cuda.Stream(self.dev)
            warm.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(warm):
                self.launch_forward()
                self.launch_backward()
            torch.cuda.current_stream(self.dev).wait_stream(warm)
            torch.cuda.synchronize(self.dev)
            side = torch.cuda.Stream(self.dev)
            # the upstream gradients arrive on a side branch; a 4-byte copy queued behind them sets a device word the backward
            # waits for ON THE DEVICE, so that the branch joins behind the backward and the backward keeps its programmatic
            # (tile by tile) dependency on the forward
            flagged = dL_dcolor_host is not None and self.overlap
            if flagged:
                self.h_one = torch.ones(1, dtype=torch.int32).pin_memory()
                self.up_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cur = torch.cuda.current_stream(self.dev)
                self.cam.copy_(cam_host, non_blocking=True)
                if flagged:
                    self.up_flag.zero_()
                if dL_dcolor_host is not None:
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        self.dL_dcolor.copy_(dL_dcolor_host, non_blocking=True)
                        if dL_ddepth_host is not None:
                            self.dL_ddepth.copy_(dL_ddepth_host, non_blocking=True)
                        if flagged:
                            self.up_flag.copy_(self.h_one, non_blocking=True)
                self.launch_forward()
                if flagged:
                    self.launch_backward(overlap_forward=True, upstream_ready=self.up_flag)
                    cur.wait_stream(side)
                else:
                    if dL_dcolor_host is not None:
                        cur.wait_stream(side)
                    self.launch_backward()
                self.h_tau.copy_(self.g_tau, non_blocking=True)
                self.h_hdr.copy_(hdr_dev, non_blocking=True)
            self.graph_host = g

    def step_host(self):
        """Replay the host-driven step and wait for it; returns (dL_dtau[6] pinned, (num_rendered, overflow) pinned)."""
        self.graph_host.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        return self.h_tau, self.h_hdr

    def step(self, use_graph=True):
        """forward + backward at the current camera / upstream gradients; no host synchronisation."""
        if use_graph:
            if self.graph_all is None:
                self.capture()
            self.graph_all.replay()
        else:
            if self.binning is None:
                self.calibrate()
            with torch.cuda.device(self.dev):
                self.launch_forward()
                self.launch_backward(overlap_forward=True)

    def status(self):
        """Raw header words of the last step -- synchronises: dict(num_rendered, overflow, spin_timeout, max_tile)."""
        out = (C.c_uint * 4)()
        with torch.cuda.device(self.dev):
            _cabi.check(_L.gsr_step_status(_p(self.geom), self._stream(), out), "step_status")
        return dict(num_rendered=int(out[0]), overflow=int(out[1]), spin_timeout=int(out[2]), max_tile=int(out[3]))

    def header(self):
        """(num_rendered, overflow) of the last forward -- synchronises.  Raises RuntimeError (GSR_ERR_TIMEOUT) when a
        compositing backward gave up waiting for a tile flag / the upstream_ready word: that step's gradients are incomplete."""
        ov, need = C.c_int(0), C.c_longlong(0)
        with torch.cuda.device(self.dev):
            _cabi.check(_L.gsr_forward_overflowed(_p(self.geom), self._stream(), C.byref(ov), C.byref(need)), "overflowed")
        self.last_num_rendered = int(need.value)
        return int(need.value), bool(ov.value)

    def step_checked(self, use_graph=True):
        """step() + overflow validation; an overflowing step is re-run with a larger workspace."""
        self.step(use_graph)
        R, ov = self.header()
        if ov:
            self.ensure_capacity(R)
            self.step(use_graph)
            R, ov = self.header()
            assert not ov
        return R
