/*
 * gsr_b200 -- C-ABI of the B200-native differentiable Gaussian-splatting rasterizer.
 *
 * Drop-in boundary for the hot path of notu97/GS-SLAM-Analytica_Jacobian: these entry points are
 * what the reference's native binding for this path exports through pybind
 * (submodules/diff-gaussian-rasterization/ext.cpp:15-19):
 *
 *   reference `rasterize_gaussians`           rasterize_points.h:18-39,  rasterize_points.cu:35-137
 *       -> gsr_rasterize_gaussians            (one call, same outputs, same "returns num_rendered")
 *          = gsr_forward_plan + gsr_forward_num_rendered + gsr_forward_render  (split so a caller
 *            can skip the reference's blocking D2H read of num_rendered, rasterizer_impl.cu:331)
 *   reference `rasterize_gaussians_backward`  rasterize_points.h:41-65,  rasterize_points.cu:139-226
 *       -> gsr_rasterize_gaussians_backward
 *   reference `mark_visible`                  rasterize_points.h:67-70,  rasterize_points.cu:228-247
 *       -> gsr_mark_visible
 *   reference scratch growth through std::function<char*(size_t)> resize callbacks
 *       (rasterize_points.cu:27-33, rasterizer_impl.cu:275-276,288-289,333-334)
 *       -> gsr_geometry_bytes / gsr_image_bytes / gsr_binning_bytes size queries + caller-owned
 *          buffers, and a gsr_alloc_fn callback for the one size that is data dependent.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer unless it is
 * named *_host.  All arrays are fp32, contiguous, laid out as the reference documents them
 * (means3D[P,3], shs[P,M,3], opacities[P], scales[P,3], rotations[P,4] (w,x,y,z),
 * cov3D_precomp[P,6], 4x4 matrices as 16 floats column-major).  A null pointer selects the
 * alternative path exactly like the reference's empty tensors (forward.cu:207,382).
 * All work is enqueued on `stream` (a cudaStream_t; 0 = legacy default stream).  Functions return
 * GSR_OK or a negative error code and never throw; gsr_error_string() explains the last failure
 * of the calling thread.
 * Process-wide state, all of it optional and off the data path: the default threshold of gsr_sort_on_demand (per call:
 * gsr_scene.sort_on_demand), the launch counter, and the opt-in stage timing (events kept per device).  Everything a call
 * computes lives in the caller's workspaces, so calls on different workspaces may run concurrently from different threads.
 */
#ifndef GSR_B200_H
#define GSR_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_OK 0
#define GSR_ERR_ARG -1        /* inconsistent arguments (mirrors the Python-side exceptions / AT_ERROR) */
#define GSR_ERR_CUDA -2       /* a CUDA call failed */
#define GSR_ERR_WORKSPACE -3  /* a caller-provided workspace is too small */
#define GSR_ERR_OVERFLOW -4   /* num_rendered exceeded the binning capacity (no-sync mode) */
#define GSR_ERR_TIMEOUT -5    /* a compositing-backward CTA gave up waiting for a tile flag / the upstream_ready word */

/* Optional: the SLAM loss of the view evaluated in the epilogue of the forward compositing kernel (the pixel values are
 * still in registers), see gsr_slam_loss below for the arithmetic (utils/slam_utils.py:56-128).  The forward then also
 * writes dL_dcolor / dL_ddepth -- the upstream gradients of the backward -- and sums[4] = {loss, dL/da, dL/db, 0}, so that
 * no kernel stands between forward and backward and the backward may start tile by tile (gsr_scene.overlap_forward). */
typedef struct gsr_fused_loss {
	const float* gt_color;          /* [3,H,W] */
	const float* gt_depth;          /* [1,H,W], read when use_depth != 0 */
	const unsigned char* grad_mask; /* [H,W] bool or null */
	const float* exposure;          /* [2] = (a, b) device pointer or null */
	float rgb_boundary_threshold, alpha;
	int use_depth, opacity_weighted;
	float* dL_dcolor;               /* [3,H,W] out */
	float* dL_ddepth;               /* [1,H,W] out */
	float* sums;                    /* [4] out */
	void* scratch;                  /* gsr_fused_loss_scratch_bytes(W, H) bytes, zero-filled ONCE by the caller */
} gsr_fused_loss;

typedef struct gsr_scene {
	int P;                       /* number of Gaussians */
	int D;                       /* active SH degree                      (settings.sh_degree) */
	int M;                       /* SH coefficients per Gaussian, 0 if shs is null */
	int W, H;                    /* image_width, image_height */
	const float* background;     /* [3] */
	const float* means3D;        /* [P,3] */
	const float* shs;            /* [P,M,3] or null */
	const float* colors_precomp; /* [P,3]  or null  (exactly one of shs / colors_precomp) */
	const float* opacities;      /* [P] */
	const float* scales;         /* [P,3]  or null */
	const float* rotations;      /* [P,4]  or null */
	const float* cov3D_precomp;  /* [P,6]  or null  (exactly one of scales+rotations / cov3D_precomp) */
	const float* viewmatrix;     /* [16] */
	const float* projmatrix;     /* [16] */
	const float* projmatrix_raw; /* [16] (only read by the backward) */
	const float* campos;         /* [3] */
	float scale_modifier;
	float tan_fovx, tan_fovy;
	int prefiltered;
	int debug;                   /* synchronise + check after every stage (reference CHECK_CUDA) */
	int accumulate_grads;        /* backward only: ADD into dL_dmeans3D/dL_dsh/dL_dcolors/dL_dopacity/dL_dscales/
	                                dL_drotations/dL_dcov3D instead of overwriting them (what autograd's AccumulateGrad
	                                does across the views of a mapping window, utils/slam_backend.py:168-232);
	                                rows of culled Gaussians are then left untouched.  dL_dmeans2D and dL_dtau are
	                                per-view quantities and are always overwritten.
	                                1: read-modify-write -- the views adding into one buffer run in stream order;
	                                2: atomic REDs -- views on DIFFERENT streams may add into one (pre-zeroed) buffer
	                                   concurrently (the summation order, hence the last bits, then varies run to run) */
	/* backward only, optional (null = off): densification statistics of the view, updated for the Gaussians with
	 * radii > 0 in the backward's per-Gaussian epilogue instead of by masked torch ops afterwards
	 * (gaussian_splatting/scene/gaussian_model.py:767-771, utils/slam_backend.py:115-121) */
	float* densify_grad_accum;   /* [P]  += || dL/dmeans2D[:2] ||   (xyz_gradient_accum) */
	float* densify_denom;        /* [P]  += 1                       (denom) */
	float* max_radii2D;          /* [P]   = max(., radii)           (max_radii2D) */
	int overlap_forward;         /* backward only: set ONLY when this call directly follows, on the same stream, the forward
	                                (gsr_forward_render / gsr_forward_nosync) of the same workspaces, and dL_dout_color /
	                                dL_dout_depth were complete before that forward was launched.  The compositing backward is
	                                then launched as a programmatic dependent of the compositing forward (CUDA programmatic
	                                stream serialization) and starts tile by tile behind it -- per-tile release/acquire flags
	                                in the geometry workspace order the data -- instead of waiting for the forward's tail.
	                                0 (default): plain stream order. */
	const unsigned int* upstream_ready; /* backward only, optional (null = off): a device word that is non-zero once
	                                dL_dout_color / dL_dout_depth are in place -- e.g. written by a 4-byte host-to-device copy
	                                queued behind the copies of the gradients on another stream.  The compositing backward
	                                waits for it on the device, so that this dependency need not be a full edge in front of
	                                the kernel (which would undo overlap_forward).  The caller clears the word before the
	                                forward of the step and joins the other stream behind the backward.  If the word (or a
	                                tile flag) does not arrive within about a second the kernel gives up, the tile contributes
	                                no gradients and the header's spin_timeout word is set: gsr_forward_overflowed() then
	                                returns GSR_ERR_TIMEOUT, gsr_step_status() reports the raw word. */
	const gsr_fused_loss* fused_loss; /* forward (gsr_forward_render / gsr_forward_nosync) only, optional (null = off):
	                                host pointer, read during the call */
	int sort_on_demand;          /* forward only: per-call form of gsr_sort_on_demand().  0 = the process default; > 0: lists
	                                longer than this are ordered on demand; < 0: every list of this call is sorted completely */
	int exact_exp;               /* forward + backward: how alpha = min(0.99, o * exp(power)) is evaluated by the compositing kernels.
	                                0 = the library default (2; environment GSR_EXACT_EXP overrides);
	                                2: the reference's own expf (forward.cu:496, backward.cu:772) in forward and backward -- T, every
	                                   threshold decision, n_contrib, n_touched and final_T are bit-identical to the reference's and
	                                   the gradients agree with it to ~1e-6 (its own run-to-run spread is ~1e-7);
	                                1: exact in the forward only (bit-identical integer outputs; gradients to ~1e-4);
	                                < 0: one ex2.approx per pair in both (6 instructions less per pair, 6 % of a tracking step at
	                                   640x480 / 100 k Gaussians; n_contrib / n_touched differ from the reference's on a ~1e-5
	                                   fraction of entries, gradients to 1e-4 .. 2e-4 in the max norm).
	                                Pass the same value to the forward and the backward of a step. */
	int tile_row_begin, tile_row_end; /* forward + backward: render only the band of tile rows [begin, end) of the view
	                                (rows of 16 pixels; 0, 0 = the whole image).  Gaussians whose tile rectangle does not
	                                reach the band are culled for this call (radii 0, zero gradient rows); pixels outside the
	                                band are not written; n_touched, every per-Gaussian gradient and dL_dtau are the band's
	                                share -- the bands of a view sum to the whole view (a view of a mapping window split over
	                                several GPUs, window.py).  Pass the same values to the backward of the same workspaces. */
	const unsigned int* spatial_order; /* forward only, optional (null = off): a device PERMUTATION of [0, P) that lists the
	                                Gaussians so that neighbours in the list are neighbours on the screen (gsr_spatial_order).
	                                The scatter of the two-kernel path (maps too large for the cooperative preprocess) then
	                                walks the Gaussians in this order: the 256 Gaussians of a CTA touch a small box of tiles
	                                with dozens of instances each instead of every tile with about one, so that a CTA counts
	                                in a box-sized shared-memory histogram, claims one slice per box tile and its pairs land
	                                in runs.  ANY permutation gives the same lists (the order inside a segment is fixed later,
	                                by depth and id); a stale one -- built for an earlier pose -- only costs speed.  The caller
	                                guarantees that it IS a permutation. */
	unsigned int* depth_cut;     /* forward only, optional (null = off), used together with spatial_order: device unsigned int[tiles],
	                                in/out, initialised by the caller to 0x7f800000 (+inf) and then left alone.  A per-tile depth HINT
	                                carried from one forward of a view to the next: the scatter puts the pairs with depth <= the
	                                tile's value at the front of the tile's segment and the others at its back (same segment, same
	                                count), the compositing forward orders the front part first and touches the back part only if
	                                the tile is still open behind it, and writes the value for the next call (the depth a little
	                                behind the tile's deepest contributor).  SLAM loops render the same views again and again, and
	                                compositing reads 2-40 % of a list: the forward then skips its sweeps over the rest.  Results do
	                                not depend on the values (sorted front ++ sorted back is the sorted list). */
} gsr_scene;

/* device allocator callback: must return a device pointer to >= bytes, aligned to 256 B, or null */
typedef void* (*gsr_alloc_fn)(void* user, size_t bytes);

/* ---- workspace sizes (reference: required<GeometryState/ImageState/BinningState>) ---- */
size_t gsr_geometry_bytes(int P, int W, int H);
size_t gsr_image_bytes(int W, int H);
size_t gsr_binning_bytes(int P, int W, int H, long long num_rendered_capacity);

/* ---- forward ---- */
/* Stage A: per-Gaussian preprocess, per-tile instance counts and tile ranges.  Writes radii[P] and
 * zero-fills n_touched[P]; leaves num_rendered in the geometry workspace on the device. */
int gsr_forward_plan(const gsr_scene* s, void* geom, size_t geom_bytes, int* radii, int* n_touched, void* stream);
/* Blocks until stage A is done and returns num_rendered (the reference's only host sync) and, if
 * max_tile_host is not null, the length of the longest per-tile list. */
int gsr_forward_num_rendered(void* geom, void* stream, long long* num_rendered_host, long long* max_tile_host);
/* Stage B: binning + compositing.  `num_rendered_host` >= 0: the exact count read with
 * gsr_forward_num_rendered;  < 0: unknown -- the kernels read it from the device and the call needs no
 * host synchronisation at all; gsr_forward_overflowed() must then be checked before results are used.
 * `max_tile_hint`: expected longest per-tile list (sizes the shared memory of the per-tile sort; longer
 * tiles still sort correctly through global memory); <= 0 = default. */
int gsr_forward_render(const gsr_scene* s, void* geom, void* binning, size_t binning_bytes,
                       long long binning_capacity, long long num_rendered_host, long long max_tile_hint,
                       void* image, size_t image_bytes,
                       float* out_color /*[3,H,W]*/, float* out_depth /*[1,H,W]*/, float* out_opacity /*[1,H,W]*/,
                       int* n_touched /*[P]*/, void* stream);
/* Synchronises and reports whether the last no-sync forward ran out of binning capacity
 * (*needed_host receives the required capacity).  Returns GSR_ERR_TIMEOUT (outputs still filled in) when a compositing
 * backward of these workspaces gave up waiting since the last forward (see gsr_scene.upstream_ready). */
int gsr_forward_overflowed(void* geom, void* stream, int* overflowed_host, long long* needed_host);
/* Synchronises and returns the raw header words of the last step: out4 = {num_rendered, overflow (binning capacity),
 * spin_timeout (0 = none, 1 = a tile flag of the forward, 2 = the upstream_ready word never arrived), longest tile list}. */
int gsr_step_status(void* geom, void* stream, unsigned int* out4_host);

/* One-call forward with the reference's semantics: allocates the binning workspace through
 * `binning_alloc` once num_rendered is known, returns it in *num_rendered_host and the buffer in
 * *binning_out (keep both for the backward). */
int gsr_rasterize_gaussians(const gsr_scene* s, void* geom, size_t geom_bytes, void* image, size_t image_bytes,
                            gsr_alloc_fn binning_alloc, void* alloc_user, void** binning_out,
                            long long* num_rendered_host, float* out_color, float* out_depth, float* out_opacity,
                            int* radii, int* n_touched, void* stream);

/* Stage A + B in one call for callers that own a pre-sized binning workspace and do not want the host to see
 * num_rendered (RasterEngine: no synchronisation, graph capturable).  Same results as gsr_forward_plan +
 * gsr_forward_render(num_rendered_host = -1); knowing the binning workspace up front lets the preprocess kernel build the
 * per-tile segments itself (cooperative launch, one grid barrier) when every CTA of its grid can be resident at once.
 * Check gsr_forward_overflowed() afterwards. */
int gsr_forward_nosync(const gsr_scene* scene, void* geometry, size_t geometry_bytes, void* binning, size_t binning_bytes,
                       long long binning_capacity, long long max_tile_hint, void* image, size_t image_bytes, float* out_color,
                       float* out_depth, float* out_opacity, int* radii, int* n_touched, void* stream);

/* 1 if gsr_forward_nosync would use the cooperative preprocess + scatter kernel for this shape on the current device */
int gsr_forward_nosync_fuses_scatter(int P, int W, int H);

/* ---- backward ---- */
/* Consumes dL/dcolor[3,H,W] and dL/ddepth[1,H,W] only (the reference drops the gradient of the
 * opacity image, __init__.py:114,139-140).  Every output row is written; outputs need no zero fill.
 * Null outputs: dL_dsh when shs is null, dL_dcolors optional, dL_dscales/dL_drotations when scales is
 * null, dL_dcov3D optional.  dL_dtau[6] = [rho(3), theta(3)], already summed over Gaussians
 * (reference: torch.sum of a [P,6] buffer, __init__.py:162-164). */
int gsr_rasterize_gaussians_backward(const gsr_scene* s, const int* radii, void* geom, void* binning,
                                     long long binning_capacity, void* image, const float* dL_dout_color,
                                     const float* dL_dout_depth, float* dL_dmeans3D /*[P,3]*/,
                                     float* dL_dmeans2D /*[P,3]*/, float* dL_dsh /*[P,M,3]*/, float* dL_dcolors /*[P,3]*/,
                                     float* dL_dopacity /*[P,1]*/, float* dL_dscales /*[P,3]*/,
                                     float* dL_drotations /*[P,4]*/, float* dL_dcov3D /*[P,6]*/, float* dL_dtau /*[6]*/,
                                     void* stream);

/* ---- markVisible ---- */
int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present /*[P] bool*/, void* stream);

/* ---- the callers on either side of the path, so that a whole tracking / mapping iteration can
 *      run without host round trips ---- */
/* Loss of utils/slam_utils.py:56-128 and its gradients in one pass over the pixels.
 *   tracking, RGB-D:    use_depth=1 opacity_weighted=1 grad_mask!=null   (get_loss_tracking_rgbd :76-89)
 *   tracking, monocular use_depth=0 opacity_weighted=1 grad_mask!=null   (get_loss_tracking_rgb  :63-73)
 *   mapping,  RGB-D:    use_depth=1 opacity_weighted=0 grad_mask=null    (get_loss_mapping_rgbd  :115-128)
 *   mapping,  monocular use_depth=0 opacity_weighted=0 grad_mask=null    (get_loss_mapping_rgb   :100-112)
 * exposure = (a, b) device pointer: image_ab = exp(a) image + b; null = no exposure (initialization).
 * sums[4] <- {loss, dL/da, dL/db, 0}.  The gradient w.r.t. the opacity image is not produced: the reference
 * rasterizer drops it (diff_gaussian_rasterization/__init__.py:114,139-140).
 * scratch: gsr_slam_loss_scratch_bytes() bytes, zero-filled ONCE by the caller, reusable across calls. */
size_t gsr_slam_loss_scratch_bytes(int W, int H);
size_t gsr_fused_loss_scratch_bytes(int W, int H);      /* scratch of gsr_fused_loss: per-tile partial sums + ticket */
int gsr_slam_loss(int W, int H, const float* color, const float* depth, const float* opacity, const float* gt_color,
                  const float* gt_depth, const unsigned char* grad_mask, const float* exposure, float rgb_boundary_threshold,
                  float alpha, int use_depth, int opacity_weighted, float* dL_dcolor, float* dL_ddepth, float* sums,
                  void* scratch, void* stream);
/* torch.optim.Adam step on [cam_rot_delta, cam_trans_delta, exposure_a, exposure_b] (utils/slam_frontend.py:129-162,
 * torch defaults) + update_pose (utils/pose_utils.py:76-93) + the camera tensors of the next render
 * (utils/camera_utils.py:96-109) as the 52-float camera block view|proj|proj_raw|campos.
 * adam_state[17] = exp_avg[8], exp_avg_sq[8], step (zero-initialised by the caller per tracked frame);
 * RT[12] = R row-major, T (world-to-camera), updated in place; status[4] <- {converged, iterations, first converged
 * iteration (sticky, 0 = not yet), 0}; dL_dexposure = the loss call's `sums` (or null), exposure may be null. */
int gsr_tracking_step(const float* dL_dtau, const float* dL_dexposure, float* exposure, float* adam_state, float* RT,
                      const float* proj_raw, float* camera_block, int* status, float lr_rot, float lr_trans, float lr_exposure,
                      float converged_threshold, void* stream);

/* ---- introspection (tests / parity): device pointers into the workspaces of the last forward ---- */
/* out[0]=records (48 B each: x,y,conic.xx,conic.xy | conic.yy,opacity,depth,r | g,b,rect_min,rect_max)
 * out[1]=tiles_touched u32[P]  out[2]=clamped u8[P]  out[3]=point_list u32[R]  out[4]=ranges u32[2*tiles]
 * out[5]=final_T f32[HW]  out[6]=n_contrib u32[HW]  out[7]=header (u32: num_rendered, overflow, num_visible) */
int gsr_debug_pointers(int P, int W, int H, void* geom, void* binning, long long binning_capacity, void* image,
                       unsigned long long* out);

/* ---- per-tile list order on demand ----
 * The reference sorts every (tile, depth) key of a view (cub::DeviceRadixSort::SortPairs, rasterizer_impl.cu:353-358),
 * although front-to-back compositing stops reading a tile's list once all its pixels are opaque (forward.cu:497-502) and
 * the backward starts at the last contributor (backward.cu:763).  By default the forward here orders lists longer than
 * a threshold only as far as they are read, depth slab by depth slab: point_list then holds the reference's list exactly
 * up to (at least) each tile's deepest contributor and is unspecified behind it; all outputs and gradients are unchanged.
 * min_list_length > 0: lists longer than this are ordered on demand; 0: every list is sorted completely (what the
 * bit-exact list tests compare); < 0: query only.  Sets the process-wide DEFAULT (used by calls whose
 * gsr_scene.sort_on_demand is 0); returns the previous value.  Default 256 (environment: GSR_LAZY_MIN). */
int gsr_sort_on_demand(int min_list_length);

/* Builds gsr_scene.spatial_order for the view last planned / rendered into `geom` (gsr_forward_plan, gsr_forward_nosync or
 * gsr_rasterize_gaussians with the same scene): Gaussians bucketed by the tile at the centre of their tile rectangle, culled
 * ones first.  order_out: device unsigned int[P].  Overwrites the tile counters of `geom` (scratch of the forward's binning:
 * the next forward rebuilds them, the backward does not read them). */
int gsr_spatial_order(const gsr_scene* scene, void* geom, size_t geom_bytes, unsigned int* order_out, void* stream);

/* ---- keyframe-window gradient reduction over the GPUs of one NVSwitch domain ----
 * Sums, in place, the packed per-Gaussian gradient buffer every rank accumulated for its share of the window (the sum
 * autograd forms over the views of utils/slam_backend.py:160-232) with ONE kernel over peer memory: rank r reduces slice r
 * in the switch (multimem.ld_reduce) and broadcasts it (multimem.st); barriers over the ranks in front and behind.
 *   multicast        the MULTICAST address of the buffer (same offset on every rank; e.g. torch symmetric memory:
 *                    handle.multicast_ptr + offset of the tensor); 16-byte aligned
 *   signal_pads      DEVICE array of world_size pointers: pad[p] = signal pad of rank p as mapped on this GPU; zero-filled
 *                    once, signal_pad_bytes each (one 32-bit flag per (CTA, peer); the kernel leaves them zero)
 *   n_floats         buffer length, a multiple of 4 (pad the allocation)
 *   ctas             CTAs per rank (<= 0: 64); clamped to what the signal pad holds and to one CTA per SM.  EVERY rank must
 *                    pass the same value and call in the same order (it is a collective)
 *   status           device int, set to 1 when a peer did not arrive within 2 s (the buffer is then incomplete)
 * The caller's stream order makes the local accumulation complete before the kernel starts.  Graph-capturable. */
int gsr_window_allreduce(float* multicast, const void* signal_pads, int rank, int world_size, size_t n_floats, int ctas,
                         size_t signal_pad_bytes, int* status, void* stream);

/* ---- measurement hooks (bench.py) ---- */
/* number of CUDA kernels this library has launched since it was loaded (bench.py: gpu_launches) */
unsigned long long gsr_kernel_launch_count(void);
/* enable/disable CUDA-event stage timing (process-wide switch; the events are kept per device and created on first use
 * there); gsr_stage_times_ms returns the durations of the last forward+backward timed on the CURRENT device:
 * {preprocess, binning, render_forward, render_backward, preprocess_backward}.  While it is on, calls are launched without
 * programmatic dependent launches (the event records sit between the kernels). */
int gsr_stage_timing(int enable);
int gsr_stage_times_ms(float* out5);
/* phase probe of a -DGSR_PHASE_PROBE build (tools/phase_probe.py): copies the u64[3, 4 or 5][4096][8] %globaltimer table
 * (kernel, CTA, phase) to the host; returns an error in the product build, which carries no probes */
int gsr_debug_probe(unsigned long long* out, size_t bytes);

const char* gsr_error_string(void);
int gsr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H */
