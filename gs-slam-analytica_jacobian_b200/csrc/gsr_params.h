// Kernel-side view of one rasterization call (all pointers are device pointers).
#pragma once
#include "gsr_common.cuh"

namespace gsr {

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
};

// launchers (defined in the .cu files, all asynchronous on `stream`)
void launch_preprocess_forward(const Scene& s, const GeomView& g, int* radii, int* n_touched, cudaStream_t stream);
// returns the number of kernels launched; cap_smem = longest tile list the 256-thread sort holds in shared memory,
// max_tile_hint = longest list expected (<= 0: unknown)
// fuse_sort: launch only the scatter; the per-tile sort then runs inside the forward compositing kernel
int launch_binning(const Scene& s, const GeomView& g, const BinView& b, size_t R_capacity, int cap_smem, long long max_tile_hint,
                   bool fuse_sort, cudaStream_t stream);
size_t tile_sort_smem_bytes(int cap_smem);
// fused_sort: the forward kernel sorts each tile's segment itself (launch_binning was called with fuse_sort = true)
void launch_render_forward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, float* out_color,
                           float* out_depth, float* out_opacity, int* n_touched, bool fused_sort, size_t R_capacity,
                           cudaStream_t stream);
void launch_render_backward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im,
                            const float* dL_dpix, const float* dL_dpix_depth, cudaStream_t stream);
void launch_preprocess_backward(const Scene& s, const GeomView& g, const int* radii, float* dL_dmeans3D,
                                float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors, float* dL_dopacity,
                                float* dL_dscales, float* dL_drotations, float* dL_dcov3D, float* dL_dtau,
                                cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present, cudaStream_t stream);

}  // namespace gsr
