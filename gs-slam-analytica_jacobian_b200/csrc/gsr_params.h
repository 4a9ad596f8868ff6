// Kernel-side view of one rasterization call (all pointers are device pointers).
#pragma once
#include "gsr_common.cuh"

namespace gsr {

// loss evaluated in the epilogue of the forward compositing kernel (gsr_fused_loss)
struct FusedLoss {
	const float* gt_color;
	const float* gt_depth;
	const unsigned char* grad_mask;
	const float* exposure;
	float rgb_boundary_threshold, alpha;
	int use_depth, opacity_weighted;
	float* dL_dcolor;
	float* dL_ddepth;
	float* sums;
	float* partials;       // [tiles][4]
	unsigned* ticket;
};

struct Scene {
	int P, D, M, W, H;
	const float* background;      // [3]
	const float* means3D;         // [P,3]
	const float* shs;             // [P,M,3] or null
	const float* colors_precomp;  // [P,3] or null
	const float* opacities;       // [P]
	const float* scales;          // [P,3] or null
	const float* rotations;       // [P,4] or null
	const float* cov3D_precomp;   // [P,6] or null
	const float* viewmatrix;      // [16] column-major
	const float* projmatrix;      // [16]
	const float* projmatrix_raw;  // [16] (backward only)
	const float* campos;          // [3]
	float scale_modifier, tan_fovx, tan_fovy, focal_x, focal_y;
	int grid_x, grid_y;
	int prefiltered;
	int accumulate_grads;
	const unsigned int* upstream_ready;   // backward: optional device word, non-zero once dL/dpixel are in place
	bool has_loss;
	FusedLoss loss;
	int overlap_forward;          // backward: launch the compositing backward as programmatic dependent of the forward before it
	int exact_exp_bwd;            // backward: the reference's expf and an exact division in the compositing backward as well
	int exact_exp;                // forward: alpha from the reference's expf instead of ex2.approx (bit-identical T, n_contrib, n_touched)
	int band_y0, band_y1;         // tile rows [band_y0, band_y1) this call renders (a band of the view); band_y1 == 0: the whole image
	const uint32_t* spatial_order; // optional permutation of [0, P): screen-coherent processing order of the scatter kernel
	uint32_t* depth_cut;           // optional [tiles] in/out: per-tile depth-key hint that splits a tile's segment into front / back (binning.cu)
	float* densify_grad_accum;    // [P] or null
	float* densify_denom;         // [P] or null
	float* max_radii2D;           // [P] or null
};

// slam_ops.cu -------------------------------------------------------------------------------------------------------
struct SlamLossArgs {
	int W, H;
	const float* color;        // [3,H,W] rendered
	const float* depth;        // [1,H,W] rendered
	const float* opacity;      // [1,H,W] rendered
	const float* gt_color;     // [3,H,W]
	const float* gt_depth;     // [1,H,W] or null when use_depth == 0
	const unsigned char* grad_mask;   // [H,W] bool or null
	const float* exposure;     // [2] = (a, b) or null (initialization: image_ab = image)
	float rgb_boundary_threshold, alpha;
	int use_depth;             // RGB-D loss (alpha * rgb + (1 - alpha) * depth) instead of rgb only
	int opacity_weighted;      // tracking variant: rgb term weighted by opacity, depth term masked by opacity > 0.95
	float* dL_dcolor;          // [3,H,W] out
	float* dL_ddepth;          // [1,H,W] out
	float* sums;               // [4] out: loss, dL/dexposure_a, dL/dexposure_b, 0
};
struct TrackingStepArgs {
	const float* dL_dtau;      // [6] = [rho, theta] gradient of the rasterizer
	const float* dL_dexposure; // [4] the loss kernel's `sums` (or null)
	float* exposure;           // [2] in/out (or null)
	float* adam_state;         // [17]: exp_avg[8], exp_avg_sq[8], step
	float* RT;                 // [12] in/out: R row-major (9), T (3)   (world-to-camera)
	const float* proj_raw;     // [16] projection_matrix as the reference stores it (transposed P)
	float* camera_block;       // [52] out: view | proj | proj_raw | campos (RasterEngine camera block)
	int* status;               // [4] out: converged, iterations done, first converged iteration, 0
	float lr_rot, lr_trans, lr_exposure, converged_threshold;
};
size_t slam_loss_scratch_bytes(int W, int H);
void launch_slam_loss(const SlamLossArgs& a, void* scratch, cudaStream_t stream);
void launch_tracking_step(const TrackingStepArgs& a, cudaStream_t stream);
void launch_window_allreduce(float* multicast, const void* signal_pads, int rank, int world, size_t n_float4, int ctas, int* status,
                             cudaStream_t stream);

// launchers (defined in the .cu files, all asynchronous on `stream`)
bool fused_scatter_fits(int P, int tiles);
// returns true when it also built the per-tile segments (cooperative fused scatter; needs the binning workspace)
bool launch_preprocess_forward(const Scene& s, const GeomView& g, int* radii, int* n_touched, cudaStream_t stream,
                               const BinView* bin = nullptr, size_t R_capacity = 0);
// returns the number of kernels launched; cap_smem = longest tile list the 256-thread sort holds in shared memory,
// max_tile_hint = longest list expected (<= 0: unknown)
// fuse_sort: launch only the scatter; the per-tile sort then runs inside the forward compositing kernel
int launch_binning(const Scene& s, const GeomView& g, const BinView& b, size_t R_capacity, int cap_smem, long long max_tile_hint,
                   bool fuse_sort, cudaStream_t stream, bool scatter_done = false, bool behind_preprocess = false);
size_t tile_sort_smem_bytes(int cap_smem);
bool depth_partition_active(const Scene& s);
// permutation of [0, P) by home tile from the records of the last forward plan (clobbers g.tile_count / g.tile_cursor)
void launch_spatial_order(const Scene& s, const GeomView& g, uint32_t* order_out, cudaStream_t stream);
// fused_sort: the forward kernel sorts each tile's segment itself (launch_binning was called with fuse_sort = true);
// lazy_min > 0: lists longer than lazy_min are ordered on demand, slab by slab, as far as the compositing gets (render.cu)
void launch_render_forward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, float* out_color,
                           float* out_depth, float* out_opacity, int* n_touched, bool fused_sort, int lazy_min, size_t R_capacity,
                           cudaStream_t stream, bool behind_preprocess = false, bool front_partition = false);
void launch_render_backward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im,
                            const float* dL_dpix, const float* dL_dpix_depth, bool overlap_forward, cudaStream_t stream);
void launch_preprocess_backward(const Scene& s, const GeomView& g, const int* radii, float* dL_dmeans3D,
                                float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors, float* dL_dopacity,
                                float* dL_dscales, float* dL_drotations, float* dL_dcov3D, float* dL_dtau,
                                bool behind_render_backward, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present, cudaStream_t stream);

}  // namespace gsr
