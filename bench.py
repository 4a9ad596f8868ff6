#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: "fwd+bwd raster iters/s incl. pose dL/dtau, 1/2/4/8 B200;
% of HBM roofline").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2_replica_mapping]

Default workload, for EVERY N (the driver must measure the multi-GPU mapping path): C2 = Replica-shaped RGB-D
mapping, 1200x680, 500 000 Gaussians, 10-keyframe window.  A *step* is one window iteration (utils/slam_backend.py:160-232):
forward + backward incl. dL/dtau of all 10 views through one replicated map, per-Gaussian gradients summed over the views;
the loss is excluded from `value` (dL/dpixel tensors are pre-generated, seed 1).  The views are sharded over the N ranks
(whole keyframes round robin, left-over keyframes split into bands of tile rows, window.py) and the packed gradient buffer
is summed by ONE collective inside the timed region -- the library's own NVSwitch kernel (csrc/window_reduce.cu; multimem.ld_reduce +
multimem.st over symmetric memory), dist.all_reduce / NCCL with --nccl-reduce or without multicast memory: strong scaling, value =
window iterations/s of the whole job (max-over-ranks device time).  At N = 1 the line also carries `also_C1`: the C1 tracking step (640x480, 100 k Gaussians,
one view, pose perturbed every step), last round's headline.

  value     device-timed (CUDA events per step, L2 flushed between steps) with everything resident in HBM
  e2e       the same iteration through the public host-driven call: slam_ops.MappingWindow.iteration -- camera blocks of
            all keyframes copied H2D from pinned memory, loss of every view evaluated in its forward's epilogue (dL/dpixel
            produced on the device, as a mapping iteration does), backward, all-reduce, per-view losses + dL/dtau + the
            gradient norm copied D2H; host blocks on every step.  (C1: RasterEngine.step_host, pinned pose + dL/dpixel.)
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json; the per-instance
            gather terms are charged on the instances the kernels actually consume (R_consumed = sum over tiles of the
            deepest contributor), not on the whole lists
  cpu_baseline  the CPU oracle port (oracle/gs_oracle.c, OpenMP) on a bounded sample of the same workload (N = 1 only)
--impl reference: the reference's own implementation of the path on this box.  The reference path has no
CPU implementation (it IS a CUDA extension), so this arm drives the UNMODIFIED reference kernels compiled
for sm_100a (oracle/_ref/libgsref.so) on the GPU, including the zero-fills and the torch.sum its binding
performs and autograd's accumulation over the views; if that library is absent it times the CPU oracle port instead.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "fwd+bwd raster iters/s incl. pose dL/dtau"
UNIT = "iters/s"
WINDOW_METRIC = "fwd+bwd raster window iters/s incl. pose dL/dtau"
WINDOW_UNIT = "window iters/s"
DEFAULT_WORKLOAD = "C2_replica_mapping"
L2_NOTE = "flushed between steps (256 MiB fill, outside the per-step events)"


def workload_config(name, V=None):
    """The `config` object of the JSON line: identical in both arms (the driver compares them)."""
    import scenes as S

    cfg = S.CONFIGS[name]
    if name.startswith(("C0", "C1")):
        what = "1 view/step, pose perturbed every step"
    else:
        what = "V=%d views/step over one replicated map, per-Gaussian gradients summed over the views" % (V or cfg["V"])
    return {"workload": "%s: %dx%d, P=%d Gaussians, SH deg %d, %s" % (name, cfg["W"], cfg["H"], cfg["P"], cfg["sh_degree"], what),
            "l2": L2_NOTE}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
def build_workload(name, steps_total, rank, device):
    import scenes as S      # (nothing of the product package: the reference arm builds its inputs here too)

    cfg = S.CONFIGS[name]
    sc = S.make_scene(name, seed=0)
    W, H = cfg["W"], cfg["H"]
    # pose perturbed every iteration: Exp(noise) * base, seeded per rank (pose seed 2)
    poses = S.noisy_poses(steps_total, sigma_rho=0.01, sigma_theta=np.radians(0.5), seed=2 + 1000 * rank)
    cams = []
    for w2c in poses:
        cam = S.make_camera(W, H, cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], w2c)
        blk = np.zeros(52, np.float32)
        blk[0:16] = cam["viewmatrix"].reshape(-1)
        blk[16:32] = cam["projmatrix"].reshape(-1)
        blk[32:48] = cam["projmatrix_raw"].reshape(-1)
        blk[48:51] = cam["campos"]
        cams.append(blk)
    cams = np.stack(cams)
    dc, dd = S.make_pixel_grads(W, H, seed=1)
    return cfg, sc, cams, dc, dd


def l2_flusher(device):
    buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def flush():
        buf.fill_(1)
    return flush


def roofline_bytes(P, R, HW, fused_sort=False, fused_scatter=False, R_consumed=None):
    """Algorithmic bytes per stage (SH degree 0).  The 44 B/instance of binning are 12 (write
    key + value) + 24 (one ideal sort pass: 12 read + 12 write) + 8 (key read for the ranges); when the forward compositing
    kernel orders its own tile the "binning" stage is the scatter alone and the ordering bytes move to "render_forward".
    R_consumed (sum over the tiles of the deepest contributor's list position): front-to-back compositing stops reading a
    list once the tile is opaque and the backward starts at the last contributor (forward.cu:497-502, backward.cu:763), so
    the per-instance GATHER terms (44 B forward, 44 + 40 B backward) and the ordering of the consumed prefix are charged on
    the instances that are actually consumed; only the 8 B/instance selection read (every pair is looked at once to find the
    front slab) stays on R.  R_consumed=None reproduces the whole-list model of §8(d)."""
    Rc = R if R_consumed is None else R_consumed
    if fused_sort:
        order = (32 * R) if R_consumed is None else (8 * R + 24 * Rc)     # selection read of every pair + one ideal pass over the prefix
    else:
        order = 0
    scat = 12 * R if fused_scatter else 0       # cooperative preprocess + scatter: the pair writes belong to "preprocess"
    return {
        "preprocess": (56 + 8 + 44) * P + scat,
        "binning": 44 * R - (32 * R if fused_sort else 0) - scat,
        "render_forward": 44 * Rc + 28 * HW + order,
        "render_backward": (44 + 40) * Rc + 24 * HW + 40 * P,
        "preprocess_backward": (60 + 40 + 68) * P,
    }


def consumed_instances(eng):
    """R_consumed of the engine's last forward: sum over the tiles of the deepest contributor (max n_contrib of the tile)."""
    from common import dev_view
    from diff_gaussian_rasterization import _cabi

    L = _cabi.load()
    ptrs = (C.c_ulonglong * 8)()
    L.gsr_debug_pointers(eng.P, eng.W, eng.H, C.c_void_p(eng.geom.data_ptr()), C.c_void_p(eng.binning.data_ptr()), eng.capacity,
                         C.c_void_p(eng.img.data_ptr()), ptrs)
    W, H = eng.W, eng.H
    nc = dev_view(ptrs[6], W * H * 4, torch.int32, eng.dev).view(H, W)
    gy, gx = (H + 15) // 16, (W + 15) // 16
    pad = torch.zeros((gy * 16, gx * 16), dtype=torch.int32, device=eng.dev)
    pad[:H, :W] = nc
    return int(pad.view(gy, 16, gx, 16).amax(dim=(1, 3)).sum().item())


def on_demand_min():
    from diff_gaussian_rasterization import _cabi
    return 0 if os.environ.get("GSR_NO_FUSED_SORT") else int(_cabi.load().gsr_sort_on_demand(-1))


def fused_sort_active(eng):
    """the forward compositing kernel orders its own tile: on demand (lists above the threshold) or completely"""
    if os.environ.get("GSR_NO_FUSED_SORT"):
        return False
    return eng.max_tile_hint > on_demand_min() > 0 or 0 < eng.max_tile_hint <= 2048


def sort_path(eng):
    if not fused_sort_active(eng):
        return "in its own kernels"
    if eng.max_tile_hint > on_demand_min() > 0:
        return "inside the forward compositing kernel, on demand (depth slab by depth slab, lists > %d entries)" % on_demand_min()
    return "fused into the forward compositing kernel"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
def cpu_oracle_iters_per_s(sc, dc, dd, max_seconds=30.0, views_per_iter=1):
    """The CPU oracle port on full forward+backward passes of ONE view of the workload; a window iteration of V views is
    V such passes (views_per_iter), so the sample is bounded to a few views and scaled."""
    from oracle.gs_oracle import Oracle

    o = Oracle(np.float32)
    o.forward(dict(sc, means3D=sc["means3D"][:2000], scales=sc["scales"][:2000], rotations=sc["rotations"][:2000],
                   opacities=sc["opacities"][:2000], shs=sc["shs"][:2000]))   # warm the library
    n, t0 = 0, time.perf_counter()
    while True:
        st = o.forward(sc)
        o.backward(st, dc, dd)
        n += 1
        el = time.perf_counter() - t0
        if el > 10.0 or n >= 3 or el + el / n > max_seconds:
            break
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    what = "%d full fwd+bwd pass(es) over one view of the same workload (P=%d, %dx%d), fp32 oracle, OpenMP" % (
        n, sc["means3D"].shape[0], sc["image_width"], sc["image_height"])
    if views_per_iter > 1:
        what += "; a window iteration is %d such views: value = views/s / %d" % (views_per_iter, views_per_iter)
    return n / el / views_per_iter, threads, what


def script_c0_baseline():
    """BASELINE.json configs[0]: the reference's own CPU script (Loss_Derivative_wrt_mu_and_cov.py, numpy) on one 640x480
    view with 15 Gaussians -- timed through its vectorised restatement oracle/loss_derivative_2d.py (pinned to the
    reference's output); the reference's nested Python loops over pixels and Gaussian pairs take minutes for the same work."""
    from oracle import loss_derivative_2d as LD

    rng = np.random.default_rng(0)
    H, W, N = 480, 640, 15
    gs = []
    for _ in range(N):
        A = rng.normal(size=(2, 2))
        gs.append(dict(mu_I=np.array([rng.uniform(0, W), rng.uniform(0, H)]), Sigma_I=A @ A.T * 40.0 + 25.0 * np.eye(2),
                       opacity=float(rng.uniform(0.2, 0.9)), color=rng.uniform(0, 1, 3), depth=float(rng.uniform(0.8, 1.5))))
    rc, rd = rng.uniform(0, 1, (H, W, 3)), rng.uniform(0, 2, (H, W))
    gc, gd = rng.uniform(0, 1, (H, W, 3)), rng.uniform(0, 2, (H, W))
    t0 = time.perf_counter()
    n = 0
    while True:
        LD.compute_gradients_2D(gs, rc, rd, gc, gd, (H, W))
        n += 1
        el = time.perf_counter() - t0
        if el > 5.0 or n >= 5:
            break
    return {"value": n / el, "unit": "dL/dmu_I + dL/dSigma_I evaluations/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d evaluation(s), 640x480, 15 Gaussians, float64 numpy (BLAS threads = all cores)" % n}


# ----------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, device):
    from diff_gaussian_rasterization import _cabi
    import scenes as S
    from diff_gaussian_rasterization.engine import RasterEngine

    L = _cabi.load()
    K, Wm = args.steps, args.warmup
    cfg, sc, cams, dc, dd = build_workload(args.workload, K + Wm, rank, device)
    t = S.to_torch(sc, device)
    eng = RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                       cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=cfg["sh_degree"], device=device)
    cams_dev = torch.from_numpy(cams).to(device)
    cams_pin = torch.from_numpy(cams).pin_memory()
    dc_pin, dd_pin = torch.from_numpy(dc).pin_memory(), torch.from_numpy(dd).pin_memory()
    eng.dL_dcolor.copy_(dc_pin)
    eng.dL_ddepth.copy_(dd_pin)
    # exact instance counts over all poses -> capacity that cannot overflow
    Rs = []
    for i in range(K + Wm):
        eng.set_camera(cams_dev[i])
        Rs.append(eng.calibrate())
    eng.capture()
    flush = l2_flusher(device)
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    # ---------------- value: device timed, inputs resident ----------------
    l0 = L.gsr_kernel_launch_count()
    eng.set_camera(cams_dev[0])
    eng.step(use_graph=False)
    torch.cuda.synchronize(device)
    launches_per_step = int(L.gsr_kernel_launch_count() - l0)
    for i in range(Wm):
        flush()
        eng.set_camera(cams_dev[i])
        eng.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(torch.cuda.current_device())
    barrier()
    sampler.start()
    for i in range(K):
        flush()
        ev[i][0].record(stream)
        eng.set_camera(cams_dev[Wm + i])
        eng.step()
        ev[i][1].record(stream)
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(step_ms))
    R_last, overflow = eng.header()
    assert not overflow, "binning capacity overflow inside the timed region"
    tau_check = eng.g_tau.cpu().numpy().copy()

    # ---------------- e2e: host buffers, copies inside the timed region ----------------
    # The public call for a host-driven step: RasterEngine.capture_host_step / step_host -- ONE graph holding the H2D of the
    # pinned camera block and upstream gradients (forked branch, overlaps the forward), forward, backward and the D2H of
    # dL/dtau + header.  Per step the host writes the next pose into the pinned block, replays, and blocks (the next pose
    # depends on dL/dtau).
    cam_step = torch.empty(52, dtype=torch.float32).pin_memory()
    cam_step.copy_(cams_pin[0])
    eng.capture_host_step(cam_step, dc_pin, dd_pin)

    def e2e_step(i):
        cam_step.copy_(cams_pin[i])            # host -> pinned staging (208 B)
        return eng.step_host()

    for i in range(Wm):
        flush()
        torch.cuda.synchronize(device)
        e2e_step(i)
    barrier()
    e2e_s = 0.0
    for i in range(K):
        flush()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        tau_pin, hdr_pin = e2e_step(Wm + i)
        e2e_s += time.perf_counter() - t0
        assert hdr_pin[1].item() == 0
    barrier()
    assert np.allclose(tau_pin.numpy(), tau_check, rtol=1e-3, atol=1e-9), "e2e path disagrees with the resident path"
    h2d = 52 * 4 + dc.nbytes + dd.nbytes
    d2h = 6 * 4 + 8

    # ---------------- per-stage CUDA-event timing (roofline) ----------------
    L.gsr_stage_timing(1)
    stage = np.zeros(5)
    out5 = (C.c_float * 5)()
    nst = max(3, min(K, 20))
    for i in range(nst):
        flush()
        eng.set_camera(cams_dev[Wm + (i % K)])
        eng.step(use_graph=False)
        _cabi.check(L.gsr_stage_times_ms(out5), "stage_times")
        stage += np.array(list(out5))
    stage /= nst
    L.gsr_stage_timing(0)
    names = ["preprocess", "binning", "render_forward", "render_backward", "preprocess_backward"]
    R_mean = float(np.mean(Rs[Wm:]))
    HW = cfg["W"] * cfg["H"]
    # instances the compositing kernels consume (deepest contributor per tile), averaged over a few poses of the timed region
    Rc = []
    for i in range(min(K, 8)):
        eng.set_camera(cams_dev[Wm + i])
        eng.launch_forward()
        Rc.append(consumed_instances(eng))
    R_cons = float(np.mean(Rc))
    fs_on, fsc_on = fused_sort_active(eng), bool(L.gsr_forward_nosync_fuses_scatter(cfg["P"], cfg["W"], cfg["H"]))
    rb = roofline_bytes(cfg["P"], R_mean, HW, fs_on, fsc_on, R_consumed=R_cons)
    rb_full = roofline_bytes(cfg["P"], R_mean, HW, fs_on, fsc_on)
    peak, peak_src = peaks()
    dom = int(np.argmax([stage[2], stage[3]])) + 2       # dominant single kernel: one of the two composite kernels
    achieved = rb[names[dom]] / (stage[dom] * 1e-3) / 1e9
    # the same pose as the reference arm's check.dL_dtau_last: the last timed step
    eng.set_camera(cams_dev[Wm + K - 1])
    eng.step()
    torch.cuda.synchronize(device)
    tau_last = [float(x) for x in eng.g_tau.cpu().numpy()]
    stages = {n: {"ms": round(float(ms), 4), "alg_bytes": int(rb[n]), "GB/s": round(rb[n] / (ms * 1e-3) / 1e9, 1) if (ms > 0 and rb[n] > 0) else None}
              for n, ms in zip(names, stage)}
    traffic, issue_pct = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, issue_pct = tj.get(names[dom]), tj.get("issue_pct_" + names[dom])

    # ---------------- reductions over ranks ----------------
    tmax, emax = dev_ms, e2e_s
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        tmax, emax = float(tt[0]), float(tt[1])
    if rank != 0:
        return None
    line = {
        "metric": METRIC, "value": world * K / (tmax * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": tmax / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.workload),
        "details": {"num_rendered_mean": R_mean, "num_consumed_mean": R_cons, "tiles": ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16),
                    "parallelism": "pose-parallel x%d (one independent tracking stream per GPU, no collective)" % world,
                    "path": "RasterEngine: CUDA graph of forward+backward, no host sync, capacity %d; per-tile sort %s"
                            % (eng.capacity, sort_path(eng))},
        "e2e": {"value": world * K / emax, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": emax / K * 1e3,
                "what": "RasterEngine.step_host: one graph with pinned pose block + pinned dL/dcolor,dL/ddepth H2D, forward, backward, dL/dtau + header D2H; host sync every step"},
        "gpu_launches": launches_per_step * K,
        "launches_per_step": {"kernels": launches_per_step, "graph_launches": 1},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src, "issue_slot_pct_ncu": issue_pct,
                     "alg_bytes_per_launch": int(rb[names[dom]]), "kernel_ms": round(float(stage[dom]), 4),
                     "byte_model": "gather terms on consumed instances (R_consumed = %.0f of R = %.0f)" % (R_cons, R_mean),
                     "frac_whole_list_model": round(rb_full[names[dom]] / (stage[dom] * 1e-3) / 1e9 / peak, 4),
                     "whole_step": {"alg_bytes": int(sum(rb.values())), "frac": round(sum(rb.values()) / (tmax / K * 1e-3) / 1e9 / peak, 4)},
                     "stages": stages},
        "clocks": clocks, "check": {"dL_dtau_last": tau_last},
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_oracle_iters_per_s(sc, dc, dd)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "script_C0": script_c0_baseline()}
    return line


def window_inputs(name, V):
    """Scene, packed camera blocks [V,52] and pre-generated upstream gradients of a window workload (both arms)."""
    import scenes as S

    cfg = S.CONFIGS[name]
    sc = S.make_scene(name, seed=0)
    W, H = cfg["W"], cfg["H"]
    poses = S.noisy_poses(V, seed=2) if name.startswith("C3") else S.arc_poses(V, radius=0.5, seed=2)
    cams = np.stack([np.concatenate([c["viewmatrix"].reshape(-1), c["projmatrix"].reshape(-1), c["projmatrix_raw"].reshape(-1),
                                     c["campos"], [0.0]]).astype(np.float32)
                     for c in (S.make_camera(W, H, cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], w2c) for w2c in poses)])
    dc, dd = S.make_pixel_grads(W, H, seed=1)
    return cfg, sc, cams, dc, dd


def run_window(args, rank, world, device):
    """Mapping-window / candidate-pose workloads (C2, C3, C4): V views over one replicated map, sharded over the ranks
    (window.py: whole views round robin, left-over views split into bands of tile rows).  One step =
    forward+backward of ALL V views; per-Gaussian gradients are accumulated in the backward kernel and summed over ranks by
    ONE all-reduce (NCCL over NVLink) per step for the mapping shapes; the candidate-pose batch (C3) needs no collective.
    Strong scaling: total work is fixed."""
    from diff_gaussian_rasterization import _cabi
    import scenes as S
    from diff_gaussian_rasterization import slam_ops as SO
    from diff_gaussian_rasterization.engine import RasterEngine
    from diff_gaussian_rasterization.window import KeyframeWindow, SwitchReducer

    L = _cabi.load()
    K, Wm = args.steps, args.warmup
    name = args.workload
    V = args.views or S.CONFIGS[name]["V"]
    reduce = not name.startswith("C3")
    plan_rank, plan_world = rank, world
    if args.as_rank_of and world == 1:      # tuning aid: rank 0's share of an N-rank window on ONE GPU, no collective
        plan_rank, plan_world, reduce = 0, int(args.as_rank_of), False
    cfg, sc, cams, dc, dd = window_inputs(name, V)
    W, H = cfg["W"], cfg["H"]
    t = S.to_torch(sc, device)
    mk = lambda **kw: RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                                   W, H, sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=cfg["sh_degree"], device=device,
                                   tau_slots=max(64, V), **kw)
    # the window gradient lives in a symmetric allocation and is summed over the ranks by the library's own NVSwitch kernel
    # (multimem.ld_reduce / multimem.st, csrc/window_reduce.cu); --nccl-reduce (or no multicast memory): dist.all_reduce
    reducer = None
    if reduce and world > 1 and not args.nccl_reduce:
        reducer = SwitchReducer.create(RasterEngine.flat_size(cfg["P"], 1, tau_slots=max(64, V)), device)
        if reducer is None:
            log("switch reducer unavailable (%s): NCCL all-reduce" % SwitchReducer.last_error)
    eng = mk(grad_flat=reducer.buffer) if reducer is not None else mk()
    cams_dev = torch.from_numpy(cams).to(device)
    cams_pin = torch.from_numpy(cams).pin_memory()
    gc_dev, gd_dev = torch.from_numpy(dc).to(device), torch.from_numpy(dd).to(device)      # ONE resident copy, read by every view
    # further engines over the same Gaussians: units overlap on separate streams, all add into ONE gradient buffer (REDs)
    n_units_max = (-(-V // plan_world)) * max(1, args.whole_bands) + 1
    extra = [mk(grad_flat=eng.grad_flat) for _ in range(min(max(args.engines, 1), n_units_max) - 1)]
    win = KeyframeWindow(eng, cams_dev, rank=plan_rank, world_size=plan_world, extra_engines=extra, split=not args.no_split,
                         reducer=reducer, whole_bands=args.whole_bands)
    win.calibrate()
    Rs = []
    for (v, y0, y1) in win.units:
        eng.set_camera(cams_dev[v]); eng.set_band(y0, y1)
        Rs.append(eng.calibrate(build_order=False))
    up = (lambda v, e: (gc_dev, gd_dev)) if extra else (lambda v: (gc_dev, gd_dev))
    flush = l2_flusher(device)
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    l0 = L.gsr_kernel_launch_count()
    win.iteration(up, reduce=reduce, upstream_precomputed=True)
    torch.cuda.synchronize(device)
    launches_per_step = int(L.gsr_kernel_launch_count() - l0)
    # the whole iteration -- every unit on every engine stream + the reduction -- as ONE CUDA graph (--no-graph: eager launches)
    if args.no_graph:
        step_value = lambda: win.iteration(up, reduce=reduce, upstream_precomputed=True)
    else:
        barrier()
        try:
            graph_value = win.capture(up, reduce=reduce, upstream_precomputed=True)
            step_value = graph_value.replay
        except Exception as ex:      # e.g. a collective that cannot be captured: eager launches, same work
            log("window graph capture failed (%s): eager launches" % ex)
            args.no_graph = True
            torch.cuda.synchronize(device)
            step_value = lambda: win.iteration(up, reduce=reduce, upstream_precomputed=True)
    for _ in range(Wm):
        flush()
        step_value()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(torch.cuda.current_device())
    barrier()
    sampler.start()
    for i in range(K):
        flush()
        ev[i][0].record(stream)
        step_value()
        ev[i][1].record(stream)
    barrier()
    clocks = sampler.stop()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    for e in win.engines:
        _, overflow = e.header()
        assert not overflow
    gnorm_check = float(eng.grad_flat[:eng.grad_flat.numel() - 8 * eng.tau_slots].double().norm())
    tau_last = [float(x) for x in win.tau_all[V - 1].cpu().numpy()] if reduce or world == 1 else None
    # the collective alone (same buffer, device-timed): what part of the step it is
    coll_ms, nccl_ms = 0.0, 0.0
    if reduce and world > 1:
        def time_collective(fn):
            for _ in range(3):
                fn()
            ce = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
            for a_, b_ in ce:
                barrier()
                a_.record(stream); fn(); b_.record(stream)
            torch.cuda.synchronize(device)
            tt = torch.tensor(float(np.median([a_.elapsed_time(b_) for a_, b_ in ce])), device=device)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            return float(tt)
        probe = torch.zeros_like(eng.grad_flat)
        nccl_ms = time_collective(lambda: torch.distributed.all_reduce(probe))
        del probe
        coll_ms = nccl_ms
        if reducer is not None:
            keep = eng.grad_flat.clone()
            coll_ms = time_collective(reducer.all_reduce)
            eng.grad_flat.copy_(keep)
            assert not reducer.timed_out(), "a rank did not arrive in the switch reduction"
            del keep

    # ---------------- e2e: the host-driven mapping iteration (slam_ops.MappingWindow) ----------------
    # Per step from pinned HOST memory: the camera blocks of all keyframes (the poses are optimisation variables).  The loss
    # of every unit is evaluated in its forward's epilogue against device-resident ground truth (uploaded once per keyframe,
    # like the reference keeps viewpoint.original_image on the GPU), so dL/dpixel never crosses PCIe.  Back to the host:
    # per-unit {loss, dL/da, dL/db}, every view's dL/dtau and the gradient norm; the host blocks every step.
    e2e_s, h2d, d2h = 0.0, 0, 0
    if reduce:
        g = torch.Generator().manual_seed(7)
        gt_c = torch.rand((V, 3, H, W), generator=g).to(device)
        gt_d = (torch.rand((V, 1, H, W), generator=g) * 3.0).to(device)
        expo = torch.zeros((V, 2), dtype=torch.float32, device=device)
        mw = SO.MappingWindow(win, gt_c, gt_d, expo, alpha=0.95)
        n_loc = max(len(win.units), 1)
        sums_pin = torch.empty((n_loc, 4), dtype=torch.float32).pin_memory()
        tau_pin = torch.empty((V, 6), dtype=torch.float32).pin_memory()
        norm_pin = torch.empty(1, dtype=torch.float32).pin_memory()
        body = eng.grad_flat[:eng.grad_flat.numel() - 8 * eng.tau_slots]

        mw.iteration(reduce=True)
        if args.no_graph:
            step_e2e = lambda: mw.iteration(reduce=True)
        else:
            barrier()
            graph_e2e = mw.capture(reduce=True)
            step_e2e = graph_e2e.replay
        sums = mw.view_sums

        def e2e_step():
            cams_dev.copy_(cams_pin, non_blocking=True)
            step_e2e()
            sums_pin.copy_(sums, non_blocking=True)
            tau_pin.copy_(win.tau_all, non_blocking=True)
            norm_pin.copy_(body.norm().reshape(1), non_blocking=True)
            stream.synchronize()

        h2d, d2h = cams.nbytes, n_loc * 16 + V * 24 + 4
    else:      # candidate-pose batch: pinned poses in, dL/dtau of the local poses out
        tau_pin = torch.empty((V, 6), dtype=torch.float32).pin_memory()

        def e2e_step():
            cams_dev.copy_(cams_pin, non_blocking=True)
            step_value()
            tau_pin.copy_(win.tau_all, non_blocking=True)
            stream.synchronize()

        h2d, d2h = cams.nbytes, V * 24
    for _ in range(Wm):
        flush()
        torch.cuda.synchronize(device)
        e2e_step()
    barrier()
    for _ in range(K):
        flush()
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        e2e_s += time.perf_counter() - t0
    barrier()
    loss_check = float(sums_pin[:, 0].sum()) if reduce else None

    # per-stage timing + consumed instances on this rank's first whole view
    names = ["preprocess", "binning", "render_forward", "render_backward", "preprocess_backward"]
    stage = np.zeros(5)
    R_view, R_cons = 0.0, 0.0
    whole = [i for i, u in enumerate(win.units) if u[2] == 0 and win.engine_of(i) is eng]
    if whole:
        eng.set_camera(cams_dev[win.units[whole[0]][0]]); eng.set_band(0, 0)
        eng.use_order(whole[0])      # the spatial order built for this unit on this engine
        eng.dL_dcolor.copy_(gc_dev); eng.dL_ddepth.copy_(gd_dev)
        R_view = float(eng.calibrate(build_order=False))
        eng.launch_forward()
        R_cons = float(consumed_instances(eng))
        L.gsr_stage_timing(1)
        out5 = (C.c_float * 5)()
        for i in range(3):
            flush()
            eng.step(use_graph=False)
            _cabi.check(L.gsr_stage_times_ms(out5), "stage_times")
            stage += np.array(list(out5))
        stage /= 3
        L.gsr_stage_timing(0)
    tmax, emax = dev_ms, e2e_s
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        tmax, emax = float(tt[0]), float(tt[1])
    if rank != 0:
        return None
    HW = W * H
    fs_on, fsc_on = fused_sort_active(eng), bool(L.gsr_forward_nosync_fuses_scatter(cfg["P"], W, H))
    rb = roofline_bytes(cfg["P"], R_view, HW, fs_on, fsc_on, R_consumed=R_cons)
    rb_full = roofline_bytes(cfg["P"], R_view, HW, fs_on, fsc_on)
    peak, peak_src = peaks()
    dom = int(np.argmax([stage[2], stage[3]])) + 2
    achieved = rb[names[dom]] / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
    grad_bytes = int(eng.grad_flat.numel() * 4)
    line = {
        "metric": WINDOW_METRIC, "value": K / (tmax * 1e-3), "unit": WINDOW_UNIT,
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": tmax / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, V),
        "details": {"views_per_s": V * K / (tmax * 1e-3), "num_rendered_per_view": R_view, "num_consumed_per_view": R_cons,
                    "units_per_rank": [["v%d" % u[0] if u[2] == 0 else "v%d[rows %d:%d]" % u for u in r] for r in win.plan],
                    "collective": ("all_reduce(sum) of %d B (packed per-Gaussian gradients + every view's dL/dtau) per step, inside the timed region, by %s; "
                                   "alone %.3f ms (dist.all_reduce / NCCL on the same buffer: %.3f ms)"
                                   % (grad_bytes, "the library's NVSwitch kernel (gsr_window_allreduce: multimem.ld_reduce + multimem.st over symmetric memory)"
                                      if reducer is not None else "dist.all_reduce (NCCL)", coll_ms, nccl_ms)) if (reduce and world > 1)
                                  else ("none (one rank: the views of the window add up in the gradient buffer of this GPU, %d B)" % grad_bytes if reduce else "none"),
                    "parallelism": "keyframe-parallel x%d (left-over views split into bands of tile rows), %d engine(s) / stream(s) per GPU"
                                   % (world, len(win.engines)),
                    "path": "KeyframeWindow over RasterEngine(s), %s, no host sync; per-tile lists ordered %s"
                            % ("eager launches" if args.no_graph else "one CUDA graph per window iteration (all units, all engine streams, the reduction)", sort_path(eng))},
        "e2e": {"value": K / emax, "unit": WINDOW_UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": emax / K * 1e3,
                "what": ("slam_ops.MappingWindow.iteration: pinned camera blocks H2D, mapping loss of every view fused into its forward "
                         "(ground truth resident), backward, all-reduce, per-view loss + dL/dtau + gradient norm D2H; host sync every step")
                        if reduce else "KeyframeWindow.iteration(reduce=False): pinned candidate poses H2D, dL/dtau of every pose D2H"},
        "gpu_launches": launches_per_step * K,
        "launches_per_step": {"kernels": launches_per_step, "graph_launches": 0 if args.no_graph else 1},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                     "alg_bytes_per_launch": int(rb[names[dom]]), "kernel_ms": round(float(stage[dom]), 4),
                     "byte_model": "gather terms on consumed instances (R_consumed = %.0f of R = %.0f per view)" % (R_cons, R_view),
                     "frac_whole_list_model": round(rb_full[names[dom]] / (stage[dom] * 1e-3) / 1e9 / peak, 4) if stage[dom] > 0 else None,
                     "whole_step": {"alg_bytes": int(sum(rb.values()) * V),
                                    "frac": round(sum(rb.values()) * V / (tmax / K * 1e-3) / 1e9 / (peak * world), 4)},
                     "stages_one_view": {n: {"ms": round(float(ms), 4), "alg_bytes": int(rb[n]),
                                             "GB/s": round(rb[n] / (ms * 1e-3) / 1e9, 1) if (ms > 0 and rb[n] > 0) else None}
                                         for n, ms in zip(names, stage)}},
        "clocks": clocks, "check": {"grad_norm": gnorm_check, "dL_dtau_last": tau_last, "e2e_loss_sum_rank0": loss_check},
    }
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        line["roofline"]["traffic"] = tj.get(name[:2] + "_" + names[dom])
        # the compositing kernels are bound by instruction issue, not by bytes: issue-slot utilisation of the
        # dominant kernel in the committed ncu capture of this shape
        line["roofline"]["issue_slot_pct_ncu"] = tj.get(name[:2] + "_issue_pct_" + names[dom])
        line["roofline"]["traffic_source"] = tj.get(name[:2] + "__source")
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_oracle_iters_per_s(S.with_camera(sc, S.make_camera(W, H, cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], S.arc_poses(V, radius=0.5, seed=2)[0])),
                                                  dc, dd, views_per_iter=V)
        line["cpu_baseline"] = {"value": v, "unit": WINDOW_UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "script_C0": script_c0_baseline()}      # BASELINE.json configs[0], the reference's own CPU script
    return line


def _tracking_setup(device):
    """C1 scene, ground-truth images rendered at the base pose, start pose = Exp(noise) * base."""
    import scenes as S
    from diff_gaussian_rasterization.engine import RasterEngine

    cfg = S.CONFIGS["C1_tum_tracking"]
    sc = S.make_scene("C1_tum_tracking", seed=0)
    t = S.to_torch(sc, device)
    mk = lambda: RasterEngine(dict(means3D=t["means3D"], opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]),
                              cfg["W"], cfg["H"], sc["tanfovx"], sc["tanfovy"], sc["bg"], sh_degree=cfg["sh_degree"], device=device)
    base = S.base_pose()
    cam = S.make_camera(cfg["W"], cfg["H"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"], base)
    eng = mk()
    eng.set_camera(RasterEngine.pack_camera(*(torch.from_numpy(cam[k]) for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos"))).to(device))
    eng.calibrate()
    eng.launch_forward()
    gt_color, gt_depth = eng.color.clone(), eng.depth.clone()
    start = S.se3_exp([0.01, -0.008, 0.012, 0.004, -0.003, 0.002]) @ base
    gmask = torch.ones((cfg["H"], cfg["W"]), dtype=torch.uint8, device=device)
    return cfg, sc, t, cam, mk, start, gt_color, gt_depth, gmask


def run_tracking_loop(args, rank, world, device):
    """Whole tracking iterations (render -> loss -> backward -> Adam -> update_pose -> camera tensors) per second:
    ours = one CUDA graph replay per iteration (slam_ops.TrackingLoop); value is device time per replay, e2e is the wall
    clock of the loop with the convergence status polled every 10 iterations."""
    from diff_gaussian_rasterization import _cabi
    from diff_gaussian_rasterization import slam_ops as SO

    L = _cabi.load()
    K, Wm = args.steps, args.warmup
    cfg, sc, t, cam, mk, start, gt_color, gt_depth, gmask = _tracking_setup(device)
    eng = mk()
    pose = SO.PoseState(start[:3, :3], start[:3, 3], cam["projmatrix_raw"], device=device)
    loop = SO.TrackingLoop(eng, pose, gt_color, gt_depth, gmask, alpha=0.9, converged_threshold=0.0)   # never "converged": fixed work
    eng.calibrate()
    l0 = L.gsr_kernel_launch_count()
    loop.capture()
    launches = int(L.gsr_kernel_launch_count() - l0) // 2
    flush = l2_flusher(device)
    stream = torch.cuda.current_stream(device)
    for _ in range(Wm):
        flush(); loop.graph.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(torch.cuda.current_device())
    torch.cuda.synchronize(device)
    sampler.start()
    for i in range(K):
        flush()
        ev[i][0].record(stream); loop.graph.replay(); ev[i][1].record(stream)
    torch.cuda.synchronize(device)
    clocks = sampler.stop()
    ms = float(sum(a.elapsed_time(b) for a, b in ev))
    t0 = time.perf_counter()
    n, first, overflow = loop.run(max_iters=K, check_every=10)
    e2e_s = time.perf_counter() - t0
    assert not overflow
    return {"metric": "tracking iterations/s (render + loss + backward + Adam + pose update)", "value": K / (ms * 1e-3), "unit": "iters/s",
            "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C1_tracking_loop: C1 scene, RGB-D tracking loss (alpha 0.9), Adam lr 0.003/0.001/0.01, one CUDA graph per iteration",
                       "l2": "flushed between iterations (device-timed value); e2e runs back to back"},
            "e2e": {"value": K / e2e_s, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16 / 10,
                    "what": "wall clock of TrackingLoop.run, status D2H + host sync every 10 iterations"},
            "gpu_launches": launches * K, "clocks": clocks,
            "check": {"loss": float(loop.ws.sums[0]), "trans_err_m": float(np.linalg.norm(pose.RT.cpu().numpy()[9:] - cam["w2c"][:3, 3]))}}


def run_tracking_loop_reference(args, device):
    """The same loop the way the reference runs it (utils/slam_frontend.py:163-192): its rasterizer kernels, torch autograd
    for the loss, torch.optim.Adam, update_pose and the camera properties as eager torch ops, `if converged` sync each iteration."""
    from oracle import slam_ref as SR

    K, Wm = args.steps, args.warmup
    cfg, sc, t, cam, mk, start, gt_color, gt_depth, gmask = _tracking_setup(device)
    Lr = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgsref.so"))
    Lr.gsref_create.restype = C.c_void_p
    h = C.c_void_p(Lr.gsref_create())
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    f32 = dict(dtype=torch.float32, device=device)
    p = lambda x: C.c_void_p(x.data_ptr())
    Rm = torch.tensor(start[:3, :3], **f32); Tm = torch.tensor(start[:3, 3], **f32)
    proj = torch.from_numpy(cam["projmatrix_raw"]).to(device)
    rot = torch.zeros(3, requires_grad=True, **f32); trans = torch.zeros(3, requires_grad=True, **f32)
    ea = torch.zeros(1, requires_grad=True, **f32); eb = torch.zeros(1, requires_grad=True, **f32)
    opt = torch.optim.Adam([{"params": [rot], "lr": 0.003}, {"params": [trans], "lr": 0.001}, {"params": [ea], "lr": 0.01}, {"params": [eb], "lr": 0.01}])
    gshape = dict(m2d=(P, 3), conic=(P, 2, 2), opac=(P, 1), col=(P, 3), dep=(P, 1), m3d=(P, 3), cov=(P, 6), sh=(P, 1, 3), sc=(P, 3), rot=(P, 4), tau=(P, 6))
    gm = gmask.view(1, H, W).bool()

    def iteration():
        nonlocal Rm, Tm
        wvt, full, center = SR.camera_tensors(Rm, Tm, proj)                       # camera properties, camera_utils.py:96-109
        wvt, full, center = wvt.contiguous(), full.contiguous(), center.contiguous()
        out = dict(color=torch.zeros((3, H, W), **f32), depth=torch.zeros((1, H, W), **f32), opacity=torch.zeros((1, H, W), **f32),
                   radii=torch.zeros((P,), dtype=torch.int32, device=device), n_touched=torch.zeros((P,), dtype=torch.int32, device=device))
        Rn = Lr.gsref_forward(h, P, 0, 1, p(t["bg"]), W, H, p(t["means3D"]), p(t["shs"]), None, p(t["opacities"]), p(t["scales"]),
                              C.c_float(1.0), p(t["rotations"]), None, p(wvt), p(full), p(center), C.c_float(sc["tanfovx"]),
                              C.c_float(sc["tanfovy"]), 0, p(out["color"]), p(out["depth"]), p(out["opacity"]), p(out["radii"]), p(out["n_touched"]), 0)
        image = out["color"].requires_grad_(True); depth = out["depth"].requires_grad_(True)
        opt.zero_grad()
        loss = SR.loss_tracking(image, depth, out["opacity"], gt_color, gt_depth, gm, ea, eb, 0.01, 0.9, False)
        loss.backward()
        g = {k: torch.zeros(s, **f32) for k, s in gshape.items()}
        Lr.gsref_backward(h, P, 0, 1, Rn, p(t["bg"]), W, H, p(t["means3D"]), p(t["shs"]), None, p(t["scales"]), C.c_float(1.0),
                          p(t["rotations"]), None, p(wvt), p(full), p(proj), p(center), C.c_float(sc["tanfovx"]), C.c_float(sc["tanfovy"]),
                          p(out["radii"]), p(image.grad), p(depth.grad), p(g["m2d"]), p(g["conic"]), p(g["opac"]), p(g["col"]), p(g["dep"]),
                          p(g["m3d"]), p(g["cov"]), p(g["sh"]), p(g["sc"]), p(g["rot"]), p(g["tau"]), 0)
        tau = torch.sum(g["tau"].view(-1, 6), dim=0)
        rot.grad, trans.grad = tau[3:].clone(), tau[:3].clone()
        with torch.no_grad():
            opt.step()
            Rm, Tm, conv = SR.update_pose(Rm, Tm, trans.detach(), rot.detach(), converged_threshold=0.0)
            rot.zero_(); trans.zero_()
        return loss

    for _ in range(Wm):
        iteration()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(K):
        loss = iteration()
    torch.cuda.synchronize(device)
    el = time.perf_counter() - t0
    v = K / el
    return {"metric": "tracking iterations/s (render + loss + backward + Adam + pose update)", "value": v, "unit": "iters/s", "n_gpus": 1,
            "steps": K, "warmup": Wm, "ms_per_step": el / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "C1_tracking_loop: reference rasterizer kernels + eager torch loss / Adam / update_pose, wall clock"},
            "cpu_baseline": {"value": v, "unit": "iters/s", "cores": 1, "kind": "reference", "device": "cuda",
                             "sample": "all %d iterations; reference CUDA kernels + torch eager ops driven by one host thread" % K},
            "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "check": {"loss": float(loss.detach()), "trans_err_m": float(np.linalg.norm(Tm.cpu().numpy() - cam["w2c"][:3, 3]))}}


class _RawView:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


# ----------------------------------------------------------------------------------------------------
def run_reference(args, rank, world, device):
    """The reference arm (rank 0 only)."""
    import scenes as S

    K, Wm = args.steps, args.warmup
    window = not args.workload.startswith(("C0", "C1"))
    if window:      # V views per step over one map, per-view gradients summed like autograd does (slam_backend.py:168-232)
        V = args.views or S.CONFIGS[args.workload]["V"]
        cfg, sc, cams, dc, dd = window_inputs(args.workload, V)
    else:
        V = 1
        cfg, sc, cams, dc, dd = build_workload(args.workload, K + Wm, 0, device)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libgsref.so")
    unit = UNIT if not window else WINDOW_UNIT
    base = {"metric": METRIC if not window else WINDOW_METRIC, "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
            "higher_is_better": True, "scaling": "weak" if not window else "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": workload_config(args.workload, V if window else None),
            "details": {"where": "one GPU (rank 0): the reference has no multi-GPU path"}}
    if not (os.path.exists(ref_so) and torch.cuda.is_available()):
        v, cores, sample = cpu_oracle_iters_per_s(sc, dc, dd, views_per_iter=V)
        base.update(value=v, ms_per_step=1e3 / v,
                    cpu_baseline={"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        return base
    Lr = C.CDLL(ref_so)
    Lr.gsref_create.restype = C.c_void_p
    h = C.c_void_p(Lr.gsref_create())
    t = S.to_torch(sc, device)
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    f32 = dict(dtype=torch.float32, device=device)
    cams_dev = torch.from_numpy(cams).to(device)
    cam = torch.zeros(52, **f32)
    gc, gd = torch.from_numpy(dc).to(device), torch.from_numpy(dd).to(device)
    out = dict(color=torch.empty((3, H, W), **f32), depth=torch.empty((1, H, W), **f32), opacity=torch.empty((1, H, W), **f32),
               radii=torch.empty((P,), dtype=torch.int32, device=device), n_touched=torch.empty((P,), dtype=torch.int32, device=device))
    gshape = dict(m2d=(P, 3), conic=(P, 2, 2), opac=(P, 1), col=(P, 3), dep=(P, 1), m3d=(P, 3), cov=(P, 6), sh=(P, 1, 3),
                  sc=(P, 3), rot=(P, 4), tau=(P, 6))
    g = {k: torch.empty(s, **f32) for k, s in gshape.items()}
    p = lambda x: C.c_void_p(x.data_ptr())
    cp = cam.data_ptr()
    view, proj, praw, campos = C.c_void_p(cp), C.c_void_p(cp + 64), C.c_void_p(cp + 128), C.c_void_p(cp + 192)

    acc = {k: torch.zeros_like(g[k]) for k in ("m3d", "m2d", "sh", "opac", "sc", "rot")} if window else None

    def step(i):
        if not window:
            return view_step(i)
        for a in acc.values():
            a.zero_()
        tau = None
        for v in range(V):
            tau = view_step(v)
            for k, a in acc.items():      # autograd's accumulation of the per-view parameter gradients
                a.add_(g[k])
        return tau

    def view_step(i):
        cam.copy_(cams_dev[i])
        for v in out.values():          # torch::full(0) of the binding, rasterize_points.cu:84-88
            v.zero_()
        R = Lr.gsref_forward(h, P, 0, 1, p(t["bg"]), W, H, p(t["means3D"]), p(t["shs"]), None, p(t["opacities"]), p(t["scales"]),
                             C.c_float(1.0), p(t["rotations"]), None, view, proj, campos, C.c_float(sc["tanfovx"]),
                             C.c_float(sc["tanfovy"]), 0, p(out["color"]), p(out["depth"]), p(out["opacity"]), p(out["radii"]),
                             p(out["n_touched"]), 0)
        for v in g.values():            # torch::zeros of the binding, rasterize_points.cu:175-185
            v.zero_()
        Lr.gsref_backward(h, P, 0, 1, R, p(t["bg"]), W, H, p(t["means3D"]), p(t["shs"]), None, p(t["scales"]), C.c_float(1.0),
                          p(t["rotations"]), None, view, proj, praw, campos, C.c_float(sc["tanfovx"]), C.c_float(sc["tanfovy"]),
                          p(out["radii"]), p(gc), p(gd), p(g["m2d"]), p(g["conic"]), p(g["opac"]), p(g["col"]), p(g["dep"]),
                          p(g["m3d"]), p(g["cov"]), p(g["sh"]), p(g["sc"]), p(g["rot"]), p(g["tau"]), 0)
        return torch.sum(g["tau"].view(-1, 6), dim=0)     # __init__.py:162-164

    flush = l2_flusher(device)
    stream = torch.cuda.current_stream(device)
    for i in range(Wm):
        flush()
        step(i)
    torch.cuda.synchronize(device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    for i in range(K):
        flush()
        ev[i][0].record(stream)
        tau = step(Wm + i)
        ev[i][1].record(stream)
    torch.cuda.synchronize(device)
    clocks = sampler.stop()
    ms = float(sum(a.elapsed_time(b) for a, b in ev))
    v = K / (ms * 1e-3)
    Lr.gsref_destroy(h)
    base.update(value=v, ms_per_step=ms / K, clocks=clocks,
                cpu_baseline={"value": v, "unit": unit, "cores": 1, "kind": "reference", "device": "cuda",
                              "sample": "all %d steps of the workload; the reference path has no CPU implementation, so its own "
                                        "CUDA kernels (unmodified sources, sm_100a) run on the B200, driven by one host thread" % K},
                e2e={"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                check={"dL_dtau_last": [float(x) for x in tau.cpu().numpy()]})
    if window:
        base["check"]["grad_norm"] = float(torch.cat([acc[k].reshape(-1) for k in ("m3d", "sh", "opac", "rot", "sc")]).double().norm())
    return base


# ----------------------------------------------------------------------------------------------------
def _claim_stdout():
    """Libraries (NCCL prints its version banner there) must not add lines to stdout: the contract is ONE JSON line.
    File descriptor 1 is pointed at stderr for the whole run; the JSON line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--views", type=int, default=0, help="override the window size of C2/C3/C4")
    ap.add_argument("--engines", type=int, default=4,
                    help="window workloads: engines (streams) the local views are dealt to (device time is flat from 2 up, the "
                         "host-driven e2e gains from overlapping the per-view copies: C2 6.61 / 6.40 / 6.19 ms, C3 19.5 / 16.7 / 16.4 ms with 2 / 3 / 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-split", action="store_true", help="window workloads: whole views only (round robin), no bands of tile rows")
    ap.add_argument("--no-also-c1", action="store_true", help="skip the nested C1 tracking record of the default run at N = 1")
    ap.add_argument("--no-graph", action="store_true", help="window workloads: eager launches instead of one CUDA graph per window iteration")
    ap.add_argument("--whole-bands", type=int, default=1, help="window workloads: cut every whole view of a rank into this many bands of tile rows (window.plan_units)")
    ap.add_argument("--nccl-reduce", action="store_true", help="window workloads at N > 1: sum the gradients with dist.all_reduce instead of the library's NVSwitch kernel")
    ap.add_argument("--as-rank-of", type=int, default=0, help="tuning aid (N = 1 only): time rank 0's share of a window sharded over this many ranks, without the collective")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        device = "cuda:%d" % local if torch.cuda.is_available() else "cpu"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        if args.workload == "C1_tracking_loop":
            print(json.dumps(run_tracking_loop_reference(args, device)), file=out, flush=True)
            return 0
        print(json.dumps(run_reference(args, rank, world, device)), file=out, flush=True)
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    if args.workload == "C1_tracking_loop":
        line = run_tracking_loop(args, rank, world, device)
    elif args.workload.startswith(("C0", "C1")):
        line = run_ours(args, rank, world, device)
    else:
        line = run_window(args, rank, world, device)
        if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_also_c1:
            # last round's headline rides along: the C1 tracking step (same flags; its own value / e2e / roofline)
            torch.cuda.empty_cache()
            a1 = argparse.Namespace(**vars(args))
            a1.workload, a1.no_cpu_baseline = "C1_tum_tracking", True
            line["also_C1"] = run_ours(a1, rank, world, device)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), file=out, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
