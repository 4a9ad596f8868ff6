// C-ABI of gsr_b200 (declared in include/gsr_b200.h).  Thin host layer: argument checks, workspace
// carve-up, stream plumbing; no torch types.  Process-wide state: the thread-local error string, the launch counter, the
// default on-demand threshold (gsr_sort_on_demand; per call: gsr_scene.sort_on_demand) and the opt-in stage timing, whose
// events are kept per device.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/gsr_b200.h"
#include "gsr_params.h"

namespace gsr {
int probe_read_preprocess(unsigned long long*);
int probe_read_scatter(unsigned long long*);
int probe_read_preprocess_backward(unsigned long long*);
int probe_read_render(unsigned long long*);
int probe_read_render_backward(unsigned long long*);
}

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};   // kernels launched by this library (bench.py: gpu_launches)
// optional stage timing (bench.py roofline): CUDA events recorded on the launching stream between stages
std::atomic<bool> g_timing{false};
const bool g_no_fused_sort = getenv("GSR_NO_FUSED_SORT") != nullptr;   // A/B switch for measurements
const bool g_no_pdl = getenv("GSR_NO_PDL") != nullptr;                 // A/B switch: no programmatic dependent launches
const bool g_no_pdl_fwd = getenv("GSR_NO_PDL_FWD") != nullptr;         // A/B switch: forward not a programmatic dependent of the preprocess
// what gsr_scene.exact_exp == 0 means: 2 = exact forward + backward (built-in), 1 = exact forward only, <= 0 = ex2.approx
// (environment GSR_EXACT_EXP overrides the built-in default: A/B switch for measurements)
const int g_exact_exp_default = getenv("GSR_EXACT_EXP") ? (atoi(getenv("GSR_EXACT_EXP")) > 0 ? atoi(getenv("GSR_EXACT_EXP")) : -1)
                                                        : GSR_EXACT_EXP_DEFAULT;
// lists longer than this are ordered on demand inside the forward compositing kernel (gsr_sort_on_demand); 0 = every list
// is sorted completely
std::atomic<int> g_lazy_min{getenv("GSR_LAZY_MIN") ? atoi(getenv("GSR_LAZY_MIN")) : GSR_LAZY_MIN_DEFAULT};
// stage-timing events live per device and are created on first use there (a process may drive several GPUs: KeyframeWindow
// with engines on different devices, one thread per GPU); the mutex covers creation and teardown only
constexpr int kMaxDevices = 64;
struct StageEvents {
	cudaEvent_t ev[7];
	std::atomic<bool> made{false};      // set (release) once all seven events exist; read without the mutex
};
StageEvents g_stage[kMaxDevices];
std::mutex g_stage_mutex;
StageEvents* stage_events(bool create)
{
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
	StageEvents& se = g_stage[dev];
	if (!se.made.load(std::memory_order_acquire)) {
		if (!create) return nullptr;
		std::lock_guard<std::mutex> lock(g_stage_mutex);
		if (!se.made.load(std::memory_order_relaxed)) {
			for (int i = 0; i < 7; i++)
				if (cudaEventCreate(&se.ev[i]) != cudaSuccess) return nullptr;
			se.made.store(true, std::memory_order_release);
		}
	}
	return &se;
}
void stage_mark(int i, cudaStream_t st)
{
	if (!g_timing.load(std::memory_order_relaxed)) return;
	if (StageEvents* se = stage_events(true)) cudaEventRecord(se->ev[i], st);
}

// NVTX ranges around the stages of a call (host side: they bracket the launches; under a tracing tool the kernels carry
// their own names).  No-ops unless a tool is attached.
struct NvtxRange {
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
};

int fail(int code, const char* fmt, const char* detail = "")
{
	snprintf(g_err, sizeof(g_err), fmt, detail);
	return code;
}

int check_cuda(const char* where)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) {
		snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s", where, cudaGetErrorString(e));
		return GSR_ERR_CUDA;
	}
	return GSR_OK;
}

int debug_sync(const gsr_scene* s, cudaStream_t st, const char* where)
{
	if (!s->debug) return check_cuda(where);
	cudaError_t e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) {
		snprintf(g_err, sizeof(g_err), "CUDA error after %s: %s", where, cudaGetErrorString(e));
		return GSR_ERR_CUDA;
	}
	return check_cuda(where);
}

size_t tiles_of(int W, int H) { return (size_t)((W + GSR_TILE - 1) / GSR_TILE) * (size_t)((H + GSR_TILE - 1) / GSR_TILE); }

int make_scene(const gsr_scene* a, gsr::Scene& s)
{
	if (!a) return fail(GSR_ERR_ARG, "null scene");
	if (a->P < 0 || a->W <= 0 || a->H <= 0) return fail(GSR_ERR_ARG, "bad P/W/H");
	if (a->W > 16 * 65535 || a->H > 16 * 65535) return fail(GSR_ERR_ARG, "image too large: tile coordinates are 16-bit");
	if (a->P > 0) {
		if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->background)
			return fail(GSR_ERR_ARG, "means3D, opacities, viewmatrix, projmatrix and background are required");
		if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
			return fail(GSR_ERR_ARG, "Please provide excatly one of either SHs or precomputed colors!");
		if (((a->scales == nullptr || a->rotations == nullptr) && a->cov3D_precomp == nullptr) ||
		    ((a->scales != nullptr || a->rotations != nullptr) && a->cov3D_precomp != nullptr))
			return fail(GSR_ERR_ARG, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
		if (a->shs && (a->M <= 0 || a->M > 16 || (a->D + 1) * (a->D + 1) > a->M || a->D < 0 || a->D > 3))
			return fail(GSR_ERR_ARG, "SH degree / coefficient count mismatch");
		if (a->shs && !a->campos) return fail(GSR_ERR_ARG, "campos is required with SHs");
		if (a->rotations && ((size_t)a->rotations & 15)) return fail(GSR_ERR_ARG, "rotations must be 16-byte aligned");
	}
	s.P = a->P; s.D = a->D; s.M = a->shs ? a->M : 0; s.W = a->W; s.H = a->H;
	s.background = a->background; s.means3D = a->means3D; s.shs = a->shs; s.colors_precomp = a->colors_precomp;
	s.opacities = a->opacities; s.scales = a->scales; s.rotations = a->rotations; s.cov3D_precomp = a->cov3D_precomp;
	s.viewmatrix = a->viewmatrix; s.projmatrix = a->projmatrix; s.projmatrix_raw = a->projmatrix_raw; s.campos = a->campos;
	s.scale_modifier = a->scale_modifier; s.tan_fovx = a->tan_fovx; s.tan_fovy = a->tan_fovy;
	// focal lengths are derived, not passed (rasterizer_impl.cu:272-273)
	s.focal_y = a->H / (2.0f * a->tan_fovy);
	s.focal_x = a->W / (2.0f * a->tan_fovx);
	s.grid_x = (a->W + GSR_TILE - 1) / GSR_TILE;
	s.grid_y = (a->H + GSR_TILE - 1) / GSR_TILE;
	s.prefiltered = a->prefiltered;
	s.accumulate_grads = a->accumulate_grads;
	s.densify_grad_accum = a->densify_grad_accum; s.densify_denom = a->densify_denom; s.max_radii2D = a->max_radii2D;
	s.overlap_forward = a->overlap_forward;
	s.upstream_ready = a->upstream_ready;
	const int exact = a->exact_exp != 0 ? a->exact_exp : g_exact_exp_default;
	s.exact_exp = exact > 0 ? 1 : 0;
	s.exact_exp_bwd = exact > 1 ? 1 : 0;
	s.spatial_order = a->spatial_order;
	s.depth_cut = a->depth_cut;
	s.band_y0 = s.band_y1 = 0;
	if (a->tile_row_end != 0 || a->tile_row_begin != 0) {
		if (a->tile_row_begin < 0 || a->tile_row_end <= a->tile_row_begin || a->tile_row_end > s.grid_y)
			return fail(GSR_ERR_ARG, "tile_row_begin / tile_row_end: need 0 <= begin < end <= ceil(H / 16)");
		if (a->densify_grad_accum || a->densify_denom || a->max_radii2D)
			return fail(GSR_ERR_ARG, "densification statistics are per view: not available for a band of tile rows");
		if (a->tile_row_begin != 0 || a->tile_row_end != s.grid_y) { s.band_y0 = a->tile_row_begin; s.band_y1 = a->tile_row_end; }
	}
	s.has_loss = a->fused_loss != nullptr;
	if (s.has_loss) {
		const gsr_fused_loss* l = a->fused_loss;
		if (!l->gt_color || !l->dL_dcolor || !l->dL_ddepth || !l->sums || !l->scratch || (l->use_depth && !l->gt_depth))
			return fail(GSR_ERR_ARG, "fused_loss: gt_color, outputs and scratch are required (gt_depth with use_depth)");
		s.loss.gt_color = l->gt_color; s.loss.gt_depth = l->gt_depth; s.loss.grad_mask = l->grad_mask; s.loss.exposure = l->exposure;
		s.loss.rgb_boundary_threshold = l->rgb_boundary_threshold; s.loss.alpha = l->alpha;
		s.loss.use_depth = l->use_depth; s.loss.opacity_weighted = l->opacity_weighted;
		s.loss.dL_dcolor = l->dL_dcolor; s.loss.dL_ddepth = l->dL_ddepth; s.loss.sums = l->sums;
		s.loss.partials = (float*)l->scratch;
		s.loss.ticket = (unsigned*)((float*)l->scratch + 4 * (size_t)s.grid_x * s.grid_y);
	}
	return GSR_OK;
}

// shared-memory capacity (list entries) of the 256-thread per-tile sort for a given longest-tile hint: one chunk;
// lists of 2049..8192 entries are sorted by the long-list kernel, longer ones through global memory
int pick_cap_smem(long long max_tile_hint)
{
	return (max_tile_hint > 0 && max_tile_hint <= GSR_SORT_CHUNK / 2) ? GSR_SORT_CHUNK / 2 : GSR_SORT_CHUNK;
}

}  // namespace

extern "C" {

const char* gsr_error_string(void) { return g_err; }
int gsr_version(void) { return 102; }

size_t gsr_geometry_bytes(int P, int W, int H) { return gsr::geom_bytes((size_t)(P > 0 ? P : 0), tiles_of(W, H)); }
size_t gsr_image_bytes(int W, int H) { return gsr::image_bytes((size_t)W, (size_t)H); }
size_t gsr_binning_bytes(int P, int W, int H, long long cap)
{
	(void)P;
	return gsr::binning_bytes((size_t)(cap > 0 ? cap : 0), tiles_of(W, H));
}

int gsr_forward_plan(const gsr_scene* a, void* geom, size_t geom_bytes, int* radii, int* n_touched, void* stream)
{
	gsr::Scene s;
	int rc = make_scene(a, s);
	if (rc) return rc;
	const size_t tiles = (size_t)s.grid_x * s.grid_y;
	if (!geom || geom_bytes < gsr::geom_bytes(s.P, tiles)) return fail(GSR_ERR_WORKSPACE, "geometry workspace too small");
	if (s.P > 0 && !radii) return fail(GSR_ERR_ARG, "radii output is required");
	cudaStream_t st = (cudaStream_t)stream;
	gsr::GeomView g = gsr::geom_view(geom, s.P, tiles);
	NvtxRange nvtx_call("gsr_forward_plan: preprocess");
	stage_mark(0, st);
	gsr::launch_preprocess_forward(s, g, radii, n_touched, st);
	stage_mark(1, st);
	if (s.P > 0) g_launches += 1;
	return debug_sync(a, st, "preprocess");
}

int gsr_forward_num_rendered(void* geom, void* stream, long long* out, long long* max_tile_out)
{
	if (!geom || !out) return fail(GSR_ERR_ARG, "null argument");
	gsr::GeomView g = gsr::geom_view(geom, 0, 0);
	gsr::GeomHeader h;
	cudaError_t e = cudaMemcpyAsync(&h, g.hdr, 32, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
	if (e != cudaSuccess) return fail(GSR_ERR_CUDA, "reading num_rendered: %s", cudaGetErrorString(e));
	*out = (long long)h.num_rendered;
	if (max_tile_out) *max_tile_out = (long long)h.max_tile_count;
	return GSR_OK;
}

static int forward_render_impl(const gsr_scene* a, const gsr::Scene& s, void* geom, void* binning, size_t binning_bytes,
                               long long capacity, long long R_host, long long max_tile_hint, void* image, size_t image_bytes,
                               float* out_color, float* out_depth, float* out_opacity, int* n_touched, cudaStream_t st,
                               bool scatter_done, bool preprocess_just_launched = false)
{
	if (capacity < 0) return fail(GSR_ERR_ARG, "negative binning capacity");
	if (R_host > capacity) return fail(GSR_ERR_WORKSPACE, "binning capacity below num_rendered");
	// num_rendered unknown to the host: an empty workspace cannot even hold the lists' overflow bookkeeping
	if (R_host < 0 && capacity == 0 && s.P > 0) return fail(GSR_ERR_WORKSPACE, "binning capacity must be positive when num_rendered is read on the device");
	if (capacity >= (1ll << 31)) return fail(GSR_ERR_ARG, "more than 2^31 tile instances are not supported");
	if (!geom || !image || image_bytes < gsr::image_bytes(s.W, s.H)) return fail(GSR_ERR_WORKSPACE, "image workspace too small");
	if (!binning || binning_bytes < gsr::binning_bytes((size_t)capacity, tiles_of(s.W, s.H))) return fail(GSR_ERR_WORKSPACE, "binning workspace too small");
	if (!out_color || !out_depth || !out_opacity || (s.P > 0 && !n_touched)) return fail(GSR_ERR_ARG, "null output");
	const size_t tiles = (size_t)s.grid_x * s.grid_y;
	gsr::GeomView g = gsr::geom_view(geom, s.P, tiles);
	gsr::BinView b = gsr::bin_view(binning, (size_t)capacity, tiles);
	gsr::ImageView im = gsr::image_view(image, s.W, s.H);
	// lists known to fit one shared-memory chunk: sort them inside the compositing kernel (one launch less, overlap)
	// or ANY list length when lists are ordered on demand
	// (when the longest list expected does not reach the threshold, the plain fused kernel does: it still copes with longer ones)
	int lazy_min = g_no_fused_sort ? 0 : (a->sort_on_demand > 0 ? a->sort_on_demand : (a->sort_on_demand < 0 ? 0 : g_lazy_min.load()));
	if (max_tile_hint > 0 && max_tile_hint <= lazy_min) lazy_min = 0;
	const bool fuse_sort = !g_no_fused_sort && (lazy_min > 0 || (max_tile_hint > 0 && max_tile_hint <= GSR_SORT_CHUNK));
	nvtxRangePushA("gsr: binning");
	const bool pdl_ok = !g_timing.load() && !a->debug && !g_no_pdl && !g_no_pdl_fwd;
	g_launches += gsr::launch_binning(s, g, b, (size_t)capacity, pick_cap_smem(max_tile_hint), max_tile_hint, fuse_sort, st, scatter_done,
	                                  preprocess_just_launched && pdl_ok);
	stage_mark(2, st);
	nvtxRangePop();
	int rc = debug_sync(a, st, "binning");
	if (rc) return rc;
	NvtxRange nvtx_render("gsr: render forward");
	// directly behind the cooperative preprocess + scatter (no kernel in between): start inside its tail
	// (cooperative preprocess + scatter, or the stand-alone scatter: either way the kernel in front built the segments)
	const bool behind_preprocess = fuse_sort && pdl_ok && s.P > 0 && capacity > 0;
	// (the stand-alone scatter in spatial order partitions the segments by depth when the caller keeps a depth_cut array)
	gsr::launch_render_forward(s, g, b, im, out_color, out_depth, out_opacity, n_touched, fuse_sort, lazy_min, (size_t)capacity, st,
	                           behind_preprocess, !scatter_done && gsr::depth_partition_active(s));
	stage_mark(3, st);
	g_launches += 1;
	return debug_sync(a, st, "render");
}

int gsr_forward_render(const gsr_scene* a, void* geom, void* binning, size_t binning_bytes, long long capacity,
                       long long R_host, long long max_tile_hint, void* image, size_t image_bytes, float* out_color,
                       float* out_depth, float* out_opacity, int* n_touched, void* stream)
{
	gsr::Scene s;
	int rc = make_scene(a, s);
	if (rc) return rc;
	return forward_render_impl(a, s, geom, binning, binning_bytes, capacity, R_host, max_tile_hint, image, image_bytes, out_color,
	                           out_depth, out_opacity, n_touched, (cudaStream_t)stream, false);
}

int gsr_forward_nosync(const gsr_scene* a, void* geom, size_t geom_bytes, void* binning, size_t binning_bytes, long long capacity,
                       long long max_tile_hint, void* image, size_t image_bytes, float* out_color, float* out_depth,
                       float* out_opacity, int* radii, int* n_touched, void* stream)
{
	gsr::Scene s;
	int rc = make_scene(a, s);
	if (rc) return rc;
	const size_t tiles = (size_t)s.grid_x * s.grid_y;
	if (!geom || geom_bytes < gsr::geom_bytes(s.P, tiles)) return fail(GSR_ERR_WORKSPACE, "geometry workspace too small");
	if (s.P > 0 && !radii) return fail(GSR_ERR_ARG, "radii output is required");
	if (capacity < 0 || !binning || binning_bytes < gsr::binning_bytes((size_t)capacity, tiles)) return fail(GSR_ERR_WORKSPACE, "binning workspace too small");
	if (capacity == 0 && s.P > 0) return fail(GSR_ERR_WORKSPACE, "binning capacity must be positive when num_rendered is read on the device");
	cudaStream_t st = (cudaStream_t)stream;
	gsr::GeomView g = gsr::geom_view(geom, s.P, tiles);
	gsr::BinView b = gsr::bin_view(binning, (size_t)capacity, tiles);
	NvtxRange nvtx_call("gsr_forward_nosync");
	stage_mark(0, st);
	const bool scatter_done = gsr::launch_preprocess_forward(s, g, radii, n_touched, st, &b, (size_t)capacity);
	stage_mark(1, st);
	if (s.P > 0) g_launches += 1;
	rc = debug_sync(a, st, "preprocess");
	if (rc) return rc;
	return forward_render_impl(a, s, geom, binning, binning_bytes, capacity, -1, max_tile_hint, image, image_bytes, out_color,
	                           out_depth, out_opacity, n_touched, st, scatter_done, !scatter_done && s.P > 0 && !a->debug);
}

int gsr_forward_nosync_fuses_scatter(int P, int W, int H) { return gsr::fused_scatter_fits(P, (int)tiles_of(W, H)) ? 1 : 0; }

int gsr_forward_overflowed(void* geom, void* stream, int* overflowed, long long* needed)
{
	if (!geom || !overflowed) return fail(GSR_ERR_ARG, "null argument");
	gsr::GeomView g = gsr::geom_view(geom, 0, 0);
	gsr::GeomHeader h;
	cudaError_t e = cudaMemcpyAsync(&h, g.hdr, 32, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
	if (e != cudaSuccess) return fail(GSR_ERR_CUDA, "reading overflow flag: %s", cudaGetErrorString(e));
	*overflowed = h.overflow != 0;
	if (needed) *needed = h.num_rendered;
	if (h.spin_timeout)
		return fail(GSR_ERR_TIMEOUT, "the compositing backward gave up waiting for %s: gradients of this step are incomplete",
		            h.spin_timeout == 2 ? "the upstream_ready word" : "a tile flag of the forward");
	return GSR_OK;
}

int gsr_step_status(void* geom, void* stream, unsigned int* out4)
{
	if (!geom || !out4) return fail(GSR_ERR_ARG, "null argument");
	gsr::GeomView g = gsr::geom_view(geom, 0, 0);
	gsr::GeomHeader h;
	cudaError_t e = cudaMemcpyAsync(&h, g.hdr, 32, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
	if (e != cudaSuccess) return fail(GSR_ERR_CUDA, "reading the step status: %s", cudaGetErrorString(e));
	out4[0] = h.num_rendered; out4[1] = h.overflow; out4[2] = h.spin_timeout; out4[3] = h.max_tile_count;
	return GSR_OK;
}

int gsr_rasterize_gaussians(const gsr_scene* a, void* geom, size_t geom_bytes, void* image, size_t image_bytes,
                            gsr_alloc_fn alloc, void* user, void** binning_out, long long* R_out, float* out_color,
                            float* out_depth, float* out_opacity, int* radii, int* n_touched, void* stream)
{
	if (!alloc || !binning_out || !R_out) return fail(GSR_ERR_ARG, "null argument");
	int rc = gsr_forward_plan(a, geom, geom_bytes, radii, n_touched, stream);
	if (rc) return rc;
	long long R = 0, max_tile = 0;
	rc = gsr_forward_num_rendered(geom, stream, &R, &max_tile);
	if (rc) return rc;
	const size_t bytes = gsr_binning_bytes(a->P, a->W, a->H, R);
	void* bin = alloc(user, bytes);
	if (!bin) return fail(GSR_ERR_WORKSPACE, "binning allocator returned null");
	*binning_out = bin;
	*R_out = R;
	return gsr_forward_render(a, geom, bin, bytes, R, R, max_tile, image, image_bytes, out_color, out_depth, out_opacity,
	                          n_touched, stream);
}

int gsr_rasterize_gaussians_backward(const gsr_scene* a, const int* radii, void* geom, void* binning, long long capacity,
                                     void* image, const float* dL_dout_color, const float* dL_dout_depth,
                                     float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors,
                                     float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                                     float* dL_dtau, void* stream)
{
	gsr::Scene s;
	int rc = make_scene(a, s);
	if (rc) return rc;
	if (!dL_dtau) return fail(GSR_ERR_ARG, "dL_dtau output is required");
	cudaStream_t st = (cudaStream_t)stream;
	if (s.P == 0) {
		cudaMemsetAsync(dL_dtau, 0, 6 * sizeof(float), st);
		return check_cuda("backward(P=0)");
	}
	if (!geom || !binning || !image || !radii) return fail(GSR_ERR_ARG, "null workspace");
	if (capacity < 0 || capacity >= (1ll << 31)) return fail(GSR_ERR_ARG, "binning capacity out of range (pass the forward's)");
	if (!dL_dout_color || !dL_dout_depth) return fail(GSR_ERR_ARG, "null upstream gradient");
	if (!s.projmatrix_raw) return fail(GSR_ERR_ARG, "projmatrix_raw is required by the backward");
	if (!dL_dmeans3D || !dL_dmeans2D || !dL_dopacity) return fail(GSR_ERR_ARG, "null gradient output");
	if (s.shs && !dL_dsh) return fail(GSR_ERR_ARG, "dL_dsh is required with SHs");
	if (s.scales && (!dL_dscales || !dL_drotations)) return fail(GSR_ERR_ARG, "dL_dscales / dL_drotations required");
	if (dL_drotations && ((size_t)dL_drotations & 15)) return fail(GSR_ERR_ARG, "dL_drotations must be 16-byte aligned");
	const size_t tiles = (size_t)s.grid_x * s.grid_y;
	gsr::GeomView g = gsr::geom_view(geom, s.P, tiles);
	gsr::BinView b = gsr::bin_view(binning, (size_t)capacity, tiles);
	gsr::ImageView im = gsr::image_view(image, s.W, s.H);
	NvtxRange nvtx_call("gsr_rasterize_gaussians_backward");
	stage_mark(4, st);
	// (with stage timing the event record sits between the two kernels: plain ordering)
	gsr::launch_render_backward(s, g, b, im, dL_dout_color, dL_dout_depth, s.overlap_forward != 0 && !g_timing.load() && !a->debug && !g_no_pdl, st);
	stage_mark(5, st);
	g_launches += 2;
	rc = debug_sync(a, st, "render backward");
	if (rc) return rc;
	gsr::launch_preprocess_backward(s, g, radii, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacity, dL_dscales,
	                                dL_drotations, dL_dcov3D, dL_dtau, !g_timing.load() && !a->debug && !g_no_pdl, st);
	stage_mark(6, st);
	return debug_sync(a, st, "preprocess backward");
}

int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream)
{
	(void)projmatrix;   // the reference's test only uses the view matrix (auxiliary.h:152-154)
	if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(GSR_ERR_ARG, "bad argument");
	gsr::launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
	return check_cuda("mark_visible");
}

int gsr_debug_probe(unsigned long long* out, size_t bytes)
{
	const size_t one = 4096 * 8;
	if (!out || bytes < 3 * one * sizeof(unsigned long long)) return fail(GSR_ERR_ARG, "probe buffer too small");
	if (bytes >= 4 * one * sizeof(unsigned long long) && gsr::probe_read_render(out + 3 * one))
		return fail(GSR_ERR_ARG, "library built without GSR_PHASE_PROBE");
	if (bytes >= 5 * one * sizeof(unsigned long long) && gsr::probe_read_render_backward(out + 4 * one))
		return fail(GSR_ERR_ARG, "library built without GSR_PHASE_PROBE");
	if (cudaDeviceSynchronize() != cudaSuccess) return fail(GSR_ERR_CUDA, "sync failed");
	if (gsr::probe_read_preprocess(out) || gsr::probe_read_scatter(out + one) || gsr::probe_read_preprocess_backward(out + 2 * one))
		return fail(GSR_ERR_ARG, "library built without GSR_PHASE_PROBE");
	return GSR_OK;
}

size_t gsr_slam_loss_scratch_bytes(int W, int H) { return gsr::slam_loss_scratch_bytes(W, H); }
size_t gsr_fused_loss_scratch_bytes(int W, int H) { return (4 * tiles_of(W, H) + 4) * sizeof(float); }

int gsr_slam_loss(int W, int H, const float* color, const float* depth, const float* opacity, const float* gt_color,
                  const float* gt_depth, const unsigned char* grad_mask, const float* exposure, float rgb_boundary_threshold,
                  float alpha, int use_depth, int opacity_weighted, float* dL_dcolor, float* dL_ddepth, float* sums,
                  void* scratch, void* stream)
{
	if (W <= 0 || H <= 0) return fail(GSR_ERR_ARG, "bad W/H");
	if (!color || !depth || !opacity || !gt_color || !dL_dcolor || !dL_ddepth || !sums || !scratch) return fail(GSR_ERR_ARG, "null argument");
	if (use_depth && !gt_depth) return fail(GSR_ERR_ARG, "gt_depth is required by the RGB-D loss");
	gsr::SlamLossArgs a;
	a.W = W; a.H = H; a.color = color; a.depth = depth; a.opacity = opacity; a.gt_color = gt_color; a.gt_depth = gt_depth;
	a.grad_mask = grad_mask; a.exposure = exposure; a.rgb_boundary_threshold = rgb_boundary_threshold; a.alpha = alpha;
	a.use_depth = use_depth; a.opacity_weighted = opacity_weighted; a.dL_dcolor = dL_dcolor; a.dL_ddepth = dL_ddepth; a.sums = sums;
	gsr::launch_slam_loss(a, scratch, (cudaStream_t)stream);
	g_launches += 1;
	return check_cuda("slam_loss");
}

int gsr_tracking_step(const float* dL_dtau, const float* dL_dexposure, float* exposure, float* adam_state, float* RT,
                      const float* proj_raw, float* camera_block, int* status, float lr_rot, float lr_trans, float lr_exposure,
                      float converged_threshold, void* stream)
{
	if (!dL_dtau || !adam_state || !RT || !proj_raw || !camera_block || !status) return fail(GSR_ERR_ARG, "null argument");
	gsr::TrackingStepArgs a;
	a.dL_dtau = dL_dtau; a.dL_dexposure = dL_dexposure; a.exposure = exposure; a.adam_state = adam_state; a.RT = RT;
	a.proj_raw = proj_raw; a.camera_block = camera_block; a.status = status; a.lr_rot = lr_rot; a.lr_trans = lr_trans;
	a.lr_exposure = lr_exposure; a.converged_threshold = converged_threshold;
	gsr::launch_tracking_step(a, (cudaStream_t)stream);
	g_launches += 1;
	return check_cuda("tracking_step");
}

int gsr_spatial_order(const gsr_scene* a, void* geom, size_t geom_bytes, unsigned int* order_out, void* stream)
{
	gsr::Scene s;
	int rc = make_scene(a, s);
	if (rc) return rc;
	const size_t tiles = (size_t)s.grid_x * s.grid_y;
	if (!geom || geom_bytes < gsr::geom_bytes(s.P, tiles)) return fail(GSR_ERR_WORKSPACE, "geometry workspace too small");
	if (s.P > 0 && !order_out) return fail(GSR_ERR_ARG, "order_out is required");
	if (s.P == 0) return GSR_OK;
	gsr::launch_spatial_order(s, gsr::geom_view(geom, s.P, tiles), order_out, (cudaStream_t)stream);
	g_launches += 3;
	return check_cuda("spatial_order");
}

int gsr_window_allreduce(float* multicast, const void* signal_pads, int rank, int world_size, size_t n_floats, int ctas,
                         size_t signal_pad_bytes, int* status, void* stream)
{
	if (!multicast || !signal_pads || !status) return fail(GSR_ERR_ARG, "null argument");
	if (world_size < 2 || world_size > 32 || rank < 0 || rank >= world_size) return fail(GSR_ERR_ARG, "bad rank / world size");
	if ((reinterpret_cast<uintptr_t>(multicast) & 15) != 0 || (n_floats & 3) != 0)
		return fail(GSR_ERR_ARG, "the multicast buffer must be 16-byte aligned and hold a multiple of 4 floats");
	if (ctas <= 0) ctas = 64;
	const size_t fit = signal_pad_bytes / (sizeof(uint32_t) * (size_t)world_size);      // one flag per (CTA, peer)
	if (fit < 1) return fail(GSR_ERR_WORKSPACE, "signal pad too small");
	if ((size_t)ctas > fit) ctas = (int)fit;
	if (ctas > 148) ctas = 148;       // every CTA must be resident: the CTAs of all ranks meet pairwise
	gsr::launch_window_allreduce(multicast, signal_pads, rank, world_size, n_floats / 4, ctas, status, (cudaStream_t)stream);
	g_launches += 1;
	return check_cuda("window_allreduce");
}

unsigned long long gsr_kernel_launch_count(void) { return g_launches.load(); }

int gsr_sort_on_demand(int min_list_length)
{
	const int prev = g_lazy_min.load();
	if (min_list_length >= 0) g_lazy_min.store(min_list_length);
	return prev;
}

int gsr_stage_timing(int enable)
{
	if (enable) {
		if (!stage_events(true)) return fail(GSR_ERR_CUDA, "cudaEventCreate failed");
		g_timing.store(true);
	} else {
		g_timing.store(false);      // the events stay (per device, reused): another thread may still be recording into them
	}
	return GSR_OK;
}

int gsr_stage_times_ms(float* out5)
{
	StageEvents* se = g_timing.load() ? stage_events(false) : nullptr;
	if (!se || !out5) return fail(GSR_ERR_ARG, "stage timing is not enabled (or nothing was timed on the current device)");
	if (cudaEventSynchronize(se->ev[6]) != cudaSuccess) return fail(GSR_ERR_CUDA, "event sync failed");
	const int a[5] = {0, 1, 2, 4, 5}, b[5] = {1, 2, 3, 5, 6};
	for (int i = 0; i < 5; i++)
		if (cudaEventElapsedTime(&out5[i], se->ev[a[i]], se->ev[b[i]]) != cudaSuccess) return fail(GSR_ERR_CUDA, "elapsed time failed");
	return GSR_OK;
}

int gsr_debug_pointers(int P, int W, int H, void* geom, void* binning, long long capacity, void* image,
                       unsigned long long* out)
{
	if (!geom || !out) return fail(GSR_ERR_ARG, "null argument");
	gsr::GeomView g = gsr::geom_view(geom, P, tiles_of(W, H));
	out[0] = (unsigned long long)g.rec;
	out[1] = (unsigned long long)g.tiles_touched;
	out[2] = (unsigned long long)g.clamped;
	out[3] = out[5] = out[6] = 0;
	out[4] = (unsigned long long)g.ranges;
	if (binning) {
		gsr::BinView b = gsr::bin_view(binning, (size_t)capacity, tiles_of(W, H));
		out[3] = (unsigned long long)b.point_list;
	}
	if (image) {
		gsr::ImageView im = gsr::image_view(image, W, H);
		out[5] = (unsigned long long)im.final_T;
		out[6] = (unsigned long long)im.n_contrib;
	}
	out[7] = (unsigned long long)g.hdr;
	return GSR_OK;
}

}  // extern "C"
