// TEST INFRASTRUCTURE (oracle) -- not product code.
//
// C-ABI shim around the UNMODIFIED reference rasterizer library
// (CudaRasterizer::Rasterizer::{forward,backward,markVisible},
//  /root/reference/submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:24-88),
// replacing the reference's torch binding (rasterize_points.cu:35-247, an 8.5-minute
// torch-header compile) with plain pointers so the tests and `bench.py --impl reference`
// can drive the real reference kernels on the GPU box through ctypes.
// Scratch buffers are grow-only cudaMalloc blocks owned by a context handle (the reference
// grows torch byte tensors through resize callbacks, rasterize_points.cu:27-33).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct Buf {
	char* ptr = nullptr;
	size_t cap = 0;
	char* grow(size_t n)
	{
		if (n > cap) {
			if (ptr) cudaFree(ptr);
			size_t want = n + n / 4 + 256;
			if (cudaMalloc((void**)&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; return nullptr; }
			cap = want;
		}
		return ptr;
	}
	void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};
struct Ctx {
	Buf geom, binning, img;
	int P = 0, R = 0, W = 0, H = 0;
};
}  // namespace

extern "C" {

void* gsref_create() { return new Ctx(); }

void gsref_destroy(void* h)
{
	Ctx* c = (Ctx*)h;
	if (!c) return;
	c->geom.release(); c->binning.release(); c->img.release();
	delete c;
}

// Mirrors RasterizeGaussiansCUDA (rasterize_points.cu:35-137) minus the tensor allocation:
// outputs must be zero-filled by the caller (the reference fills them, :84-88).
int gsref_forward(void* h, int P, int D, int M, const float* bg, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos,
	float tan_fovx, float tan_fovy, int prefiltered,
	float* out_color, float* out_depth, float* out_opacity, int* radii, int* n_touched, int debug)
{
	Ctx* c = (Ctx*)h;
	c->P = P; c->W = W; c->H = H; c->R = 0;
	if (P == 0) return 0;
	std::function<char*(size_t)> g = [c](size_t n) { return c->geom.grow(n); };
	std::function<char*(size_t)> b = [c](size_t n) { return c->binning.grow(n); };
	std::function<char*(size_t)> i = [c](size_t n) { return c->img.grow(n); };
	int R = -1;
	try {
		R = CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, bg, W, H, means3D, shs,
			colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp,
			viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, prefiltered != 0,
			out_color, out_depth, out_opacity, radii, n_touched, debug != 0);
	} catch (...) {
		return -1;
	}
	c->R = R;
	return R;
}

// Mirrors RasterizeGaussiansBackwardCUDA (rasterize_points.cu:139-226): all gradient
// outputs must be zero-filled by the caller (torch::zeros there, :175-185).
int gsref_backward(void* h, int P, int D, int M, int R, const float* bg, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
	float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* projmatrix_raw, const float* campos,
	float tan_fovx, float tan_fovy, const int* radii,
	const float* dL_dpix, const float* dL_dpix_depth,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepths,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, float* dL_dtau,
	int debug)
{
	Ctx* c = (Ctx*)h;
	if (P == 0) return 0;
	try {
		CudaRasterizer::Rasterizer::backward(P, D, M, R, bg, W, H, means3D, shs, colors_precomp,
			scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, projmatrix_raw,
			campos, tan_fovx, tan_fovy, radii, c->geom.ptr, c->binning.ptr, c->img.ptr,
			dL_dpix, dL_dpix_depth, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepths,
			dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dtau, debug != 0);
	} catch (...) {
		return -1;
	}
	return 0;
}

int gsref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, unsigned char* present)
{
	if (P == 0) return 0;
	CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, (bool*)present);
	return 0;
}

// Device addresses of the reference's internal state of the last forward, decoded with the
// reference's own fromChunk carve-up (rasterizer_impl.cu:155-194).  Order of out[]:
// 0 depths f32[P]        1 clamped u8[3P]      2 means2D f32[2P]   3 cov3D f32[6P]
// 4 conic_opacity f32[4P] 5 rgb f32[3P]        6 tiles_touched u32[P] 7 point_offsets u32[P]
// 8 point_list u32[R]    9 point_list_keys u64[R]
// 10 accum_alpha f32[HW] 11 n_contrib u32[HW]  12 ranges u32[2*HW] (first 2*tiles used)
int gsref_state_ptrs(void* h, unsigned long long* out)
{
	Ctx* c = (Ctx*)h;
	if (!c || !c->geom.ptr) return -1;
	char* p = c->geom.ptr;
	auto g = CudaRasterizer::GeometryState::fromChunk(p, c->P);
	out[0] = (unsigned long long)g.depths; out[1] = (unsigned long long)g.clamped;
	out[2] = (unsigned long long)g.means2D; out[3] = (unsigned long long)g.cov3D;
	out[4] = (unsigned long long)g.conic_opacity; out[5] = (unsigned long long)g.rgb;
	out[6] = (unsigned long long)g.tiles_touched; out[7] = (unsigned long long)g.point_offsets;
	out[8] = out[9] = 0;
	if (c->binning.ptr) {
		char* q = c->binning.ptr;
		auto b = CudaRasterizer::BinningState::fromChunk(q, c->R);
		out[8] = (unsigned long long)b.point_list; out[9] = (unsigned long long)b.point_list_keys;
	}
	char* r = c->img.ptr;
	auto im = CudaRasterizer::ImageState::fromChunk(r, (size_t)c->W * c->H);
	out[10] = (unsigned long long)im.accum_alpha; out[11] = (unsigned long long)im.n_contrib;
	out[12] = (unsigned long long)im.ranges;
	return 0;
}

int gsref_sync() { return (int)cudaDeviceSynchronize(); }
const char* gsref_last_error() { return cudaGetErrorString(cudaGetLastError()); }

}  // extern "C"
