// Forward per-Gaussian preprocess: EWA projection of the 3D covariance to a 2D conic, SH -> RGB,
// radius and tile bounds.  Replaces reference preprocessCUDA (cuda_rasterizer/forward.cu:157-401)
// and checkFrustum (rasterizer_impl.cu:54-66).
//
// Arithmetic contract: radii, tile rectangles and depth keys must be bit-identical to the reference,
// so the fp32 expression trees that feed them (view/clip transform, cov3D = (S R)^T (S R),
// cov2D = (W J)^T Vrk^T (W J), eigenvalue radius, double-precision ndc2Pix, getRect) are written
// with the same association order as the reference + glm 0.9.9 (type_mat3x3.inl:486-518), and are
// compiled with the same default -fmad contraction.
#include <cstdlib>
#include "gsr_params.h"

namespace gsr {

namespace {

struct M3 {            // column-major 3x3: m[c][r], same indexing as glm::mat3
	float m[3][3];
};
__device__ __forceinline__ M3 m3_mul(const M3& A, const M3& B)
{
	M3 R;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
	return R;
}
__device__ __forceinline__ M3 m3_transpose(const M3& A)
{
	M3 R;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++) R.m[c][r] = A.m[r][c];
	return R;
}

__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* __restrict__ sh, float3 pos, float3 campos, unsigned& clamped)
{
	float3 dir = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
	float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
	dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
	float res[3];
#pragma unroll
	for (int ch = 0; ch < 3; ch++) res[ch] = kSH_C0 * sh[ch];
	if (deg > 0) {
		const float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
		for (int ch = 0; ch < 3; ch++)
			res[ch] = res[ch] - kSH_C1 * y * sh[3 + ch] + kSH_C1 * z * sh[6 + ch] - kSH_C1 * x * sh[9 + ch];
		if (deg > 1) {
			const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
			for (int ch = 0; ch < 3; ch++)
				res[ch] = res[ch] + kSH_C2[0] * xy * sh[12 + ch] + kSH_C2[1] * yz * sh[15 + ch] +
				          kSH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + ch] + kSH_C2[3] * xz * sh[21 + ch] +
				          kSH_C2[4] * (xx - yy) * sh[24 + ch];
			if (deg > 2) {
#pragma unroll
				for (int ch = 0; ch < 3; ch++)
					res[ch] = res[ch] + kSH_C3[0] * y * (3.0f * xx - yy) * sh[27 + ch] + kSH_C3[1] * xy * z * sh[30 + ch] +
					          kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + ch] +
					          kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + ch] +
					          kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + ch] + kSH_C3[5] * z * (xx - yy) * sh[42 + ch] +
					          kSH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + ch];
			}
		}
	}
	clamped = 0;
#pragma unroll
	for (int ch = 0; ch < 3; ch++) {
		res[ch] += 0.5f;
		if (res[ch] < 0.0f) clamped |= 1u << ch;
		res[ch] = fmaxf(res[ch], 0.0f);
	}
	return make_float3(res[0], res[1], res[2]);
}

// FUSED_SCATTER (cooperative launch, every CTA resident): the kernel goes on, behind ONE grid-wide barrier, to build the
// per-tile segments itself -- the Gaussian's rectangle / depth / id are still in registers and the CTA's tile histogram
// is still in shared memory, so the separate scatter kernel's reload, its recount and the serial last-CTA scan fall away:
// every CTA scans the (complete) tile counters redundantly, claims its slice of every touched tile with one global
// atomic, and stores its (depth bits, id) pairs.
constexpr int kFusedMaxTiles = 4096;      // tile counters one CTA scans / claims in registers (16 per thread)
struct FusedScatterArgs {
	uint2* pairs;          // binning workspace: [capacity] (depth bits, id)
	unsigned capacity;
};

template <bool FUSED_SCATTER>
__global__ void __launch_bounds__(256) preprocess_forward_kernel(Scene s, GeomView g, int* __restrict__ radii,
                                                                 int* __restrict__ n_touched, int vec_mask, int hist_smem,
                                                                 FusedScatterArgs fa)
{
	__shared__ __align__(16) float s_mean[768];
	__shared__ __align__(16) float s_scale[768];
	__shared__ __align__(16) float s_col[768];
	__shared__ float s_view[16], s_proj[16];
	__shared__ unsigned s_red[16];
	__shared__ bool s_last;
	extern __shared__ uint32_t s_hist[];   // [tiles] CTA-private tile histogram (hist_smem != 0)

	GSR_PROBE(0, 0);
	if (!FUSED_SCATTER) pdl_launch_dependents();      // two-kernel path: the scatter (launched as programmatic dependent) moves into this grid's tail
	const int row0 = blockIdx.x * 256;
	const int idx = row0 + threadIdx.x;
	const int n_tiles = s.grid_x * s.grid_y;
	if (hist_smem)
		for (int t = threadIdx.x; t < n_tiles; t += 256) s_hist[t] = 0;
	if (threadIdx.x < 16) {
		s_view[threadIdx.x] = s.viewmatrix[threadIdx.x];
		s_proj[threadIdx.x] = s.projmatrix[threadIdx.x];
	}
	const bool dc_only = (s.colors_precomp != nullptr) || (s.M == 1);
	// The rows of this CTA's 256 Gaussians are ONE contiguous block per input array: a full CTA with 16-byte aligned arrays
	// has them fetched by the TMA unit (one elected thread, cp.async.bulk, completion on an mbarrier); the ragged last CTA and
	// unaligned arrays take the vector / scalar loads (vec_mask bit 3: bulk copies allowed)
	__shared__ __align__(8) uint64_t s_bar;
	const bool bulk = (vec_mask & 8) && row0 + 256 <= s.P && (vec_mask & 1) && (!s.scales || (vec_mask & 2)) && (!dc_only || (vec_mask & 4));
	if (bulk) {
		if (threadIdx.x == 0) mbar_init(&s_bar, 1);
	} else {
		load_rows3(s.means3D, row0, s.P, s_mean, vec_mask & 1);
		if (s.scales) load_rows3(s.scales, row0, s.P, s_scale, vec_mask & 2);
		if (dc_only) load_rows3(s.colors_precomp ? s.colors_precomp : s.shs, row0, s.P, s_col, vec_mask & 4);
	}
	__syncthreads();
	if (bulk) {
		if (threadIdx.x == 0) {
			constexpr unsigned kBytes = 256 * 3 * sizeof(float);
			mbar_arrive_expect_tx(&s_bar, kBytes * (1u + (s.scales ? 1u : 0u) + (dc_only ? 1u : 0u)));
			bulk_copy_g2s(s_mean, s.means3D + (size_t)row0 * 3, kBytes, &s_bar);
			if (s.scales) bulk_copy_g2s(s_scale, s.scales + (size_t)row0 * 3, kBytes, &s_bar);
			if (dc_only) bulk_copy_g2s(s_col, (s.colors_precomp ? s.colors_precomp : s.shs) + (size_t)row0 * 3, kBytes, &s_bar);
		}
		// every thread polls the barrier's phase 0 (bounded: a copy that never lands must not hang the GPU)
		unsigned spins = 0;
		while (!mbar_try_wait(&s_bar, 0u)) {
			if (++spins > (1u << 22)) { g.hdr->spin_timeout = 3; break; }
		}
	}
	GSR_PROBE(0, 1);

	unsigned my_tiles = 0, my_vis = 0, rect_lo = 0, rect_hi = 0, depth_bits = 0;
	if (idx < s.P) {
		int out_radius = 0;
		GaussRec rec;
		rec.q0 = make_float4(0.f, 0.f, 0.f, 0.f); rec.q1 = rec.q0; rec.q2 = rec.q0;
		unsigned clamped = 0;
		const float3 p_orig = {s_mean[3 * threadIdx.x], s_mean[3 * threadIdx.x + 1], s_mean[3 * threadIdx.x + 2]};
		const float* vm = s_view;
		const float* pm = s_proj;
		// transformPoint4x4 / 4x3 (auxiliary.h:58-77)
		const float hx = pm[0] * p_orig.x + pm[4] * p_orig.y + pm[8] * p_orig.z + pm[12];
		const float hy = pm[1] * p_orig.x + pm[5] * p_orig.y + pm[9] * p_orig.z + pm[13];
		const float hw = pm[3] * p_orig.x + pm[7] * p_orig.y + pm[11] * p_orig.z + pm[15];
		const float p_w = 1.0f / (hw + 0.0000001f);
		const float projx = hx * p_w, projy = hy * p_w;
		float3 t;
		t.x = vm[0] * p_orig.x + vm[4] * p_orig.y + vm[8] * p_orig.z + vm[12];
		t.y = vm[1] * p_orig.x + vm[5] * p_orig.y + vm[9] * p_orig.z + vm[13];
		t.z = vm[2] * p_orig.x + vm[6] * p_orig.y + vm[10] * p_orig.z + vm[14];
		const float depth = t.z;
		if (!(depth <= 0.2f)) {   // near-plane cull only (auxiliary.h:154)
			float cov3D[6];
			if (s.cov3D_precomp) {
#pragma unroll
				for (int i = 0; i < 6; i++) cov3D[i] = __ldg(s.cov3D_precomp + (size_t)idx * 6 + i);
			} else {
				// computeCov3D (forward.cu:120-154): quaternion used un-normalised
				const float mod = s.scale_modifier;
				M3 S;
#pragma unroll
				for (int c = 0; c < 3; c++)
#pragma unroll
					for (int r = 0; r < 3; r++) S.m[c][r] = (c == r) ? 1.0f : 0.0f;
				S.m[0][0] = mod * s_scale[3 * threadIdx.x];
				S.m[1][1] = mod * s_scale[3 * threadIdx.x + 1];
				S.m[2][2] = mod * s_scale[3 * threadIdx.x + 2];
				const float4 q = __ldg(reinterpret_cast<const float4*>(s.rotations) + idx);
				const float r = q.x, x = q.y, y = q.z, z = q.w;
				M3 R;
				// Which product of a sum  a*b +- c*d  is rounded and which is fused is the compiler's choice and depends on the
				// surrounding code; the reference build (forward.cu:131-140 under nvcc 12.9 for sm_100a, read off its SASS with
				// tools/sass_symbolic.py) rounds r*z, r*x, x*z, y*y, z*z and fuses the other factor pair.  Pinned with intrinsics so
				// that cov3D -- and through it the conic, alpha and T -- is bit-identical in every instantiation of this kernel.
				const float rz = __fmul_rn(r, z), rx = __fmul_rn(r, x), xz = __fmul_rn(x, z), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
				R.m[0][0] = 1.f - 2.f * __fadd_rn(yy, zz);
				R.m[0][1] = 2.f * __fmaf_rn(x, y, -rz);
				R.m[0][2] = 2.f * __fmaf_rn(r, y, xz);
				R.m[1][0] = 2.f * __fmaf_rn(x, y, rz);
				R.m[1][1] = 1.f - 2.f * __fmaf_rn(x, x, zz);
				R.m[1][2] = 2.f * __fmaf_rn(y, z, -rx);
				R.m[2][0] = 2.f * __fmaf_rn(-r, y, xz);
				R.m[2][1] = 2.f * __fmaf_rn(y, z, rx);
				R.m[2][2] = 1.f - 2.f * __fmaf_rn(x, x, yy);
				const M3 Mm = m3_mul(S, R);
				const M3 Sigma = m3_mul(m3_transpose(Mm), Mm);
				cov3D[0] = Sigma.m[0][0]; cov3D[1] = Sigma.m[0][1]; cov3D[2] = Sigma.m[0][2];
				cov3D[3] = Sigma.m[1][1]; cov3D[4] = Sigma.m[1][2]; cov3D[5] = Sigma.m[2][2];
			}
			// computeCov2D (forward.cu:76-115)
			const float limx = 1.3f * s.tan_fovx;
			const float limy = 1.3f * s.tan_fovy;
			const float txtz = t.x / t.z;
			const float tytz = t.y / t.z;
			t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
			t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
			M3 J;
			J.m[0][0] = s.focal_x / t.z; J.m[0][1] = 0.0f; J.m[0][2] = -(s.focal_x * t.x) / (t.z * t.z);
			J.m[1][0] = 0.0f; J.m[1][1] = s.focal_y / t.z; J.m[1][2] = -(s.focal_y * t.y) / (t.z * t.z);
			J.m[2][0] = 0.0f; J.m[2][1] = 0.0f; J.m[2][2] = 0.0f;
			M3 Wm;
			Wm.m[0][0] = vm[0]; Wm.m[0][1] = vm[4]; Wm.m[0][2] = vm[8];
			Wm.m[1][0] = vm[1]; Wm.m[1][1] = vm[5]; Wm.m[1][2] = vm[9];
			Wm.m[2][0] = vm[2]; Wm.m[2][1] = vm[6]; Wm.m[2][2] = vm[10];
			const M3 T = m3_mul(Wm, J);
			M3 Vrk;
			Vrk.m[0][0] = cov3D[0]; Vrk.m[0][1] = cov3D[1]; Vrk.m[0][2] = cov3D[2];
			Vrk.m[1][0] = cov3D[1]; Vrk.m[1][1] = cov3D[3]; Vrk.m[1][2] = cov3D[4];
			Vrk.m[2][0] = cov3D[2]; Vrk.m[2][1] = cov3D[4]; Vrk.m[2][2] = cov3D[5];
			M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
			cov.m[0][0] += 0.3f;
			cov.m[1][1] += 0.3f;
			const float cx = cov.m[0][0], cy = cov.m[0][1], cz = cov.m[1][1];
			const float det = (cx * cz - cy * cy);
			if (det != 0.0f) {
				const float det_inv = 1.f / det;
				const float3 conic = {cz * det_inv, -cy * det_inv, cx * det_inv};
				const float mid = 0.5f * (cx + cz);
				const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
				const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
				const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
				// ndc2Pix in double (auxiliary.h:41-44)
				const float pix_x = ((projx + 1.0) * s.W - 1.0) * 0.5;
				const float pix_y = ((projy + 1.0) * s.H - 1.0) * 0.5;
				// getRect (auxiliary.h:46-56), int max_radius
				const int mr = (int)my_radius;
				const unsigned gx = (unsigned)s.grid_x, gy = (unsigned)s.grid_y;
				const unsigned rminx = min(gx, (unsigned)max((int)0, (int)((pix_x - mr) / GSR_TILE)));
				unsigned rminy = min(gy, (unsigned)max((int)0, (int)((pix_y - mr) / GSR_TILE)));
				const unsigned rmaxx = min(gx, (unsigned)max((int)0, (int)((pix_x + mr + GSR_TILE - 1) / GSR_TILE)));
				unsigned rmaxy = min(gy, (unsigned)max((int)0, (int)((pix_y + mr + GSR_TILE - 1) / GSR_TILE)));
				if (s.band_y1 > 0) {      // a band of tile rows: Gaussians that do not reach it are culled for this call (radii 0)
					rminy = min(max(rminy, (unsigned)s.band_y0), (unsigned)s.band_y1);
					rmaxy = min(max(rmaxy, (unsigned)s.band_y0), (unsigned)s.band_y1);
				}
				const unsigned area = (rmaxx - rminx) * (rmaxy - rminy);
				if (area != 0) {
					float3 rgb;
					if (s.colors_precomp) {
						rgb = make_float3(s_col[3 * threadIdx.x], s_col[3 * threadIdx.x + 1], s_col[3 * threadIdx.x + 2]);
					} else if (s.M == 1) {
						const float3 campos = {__ldg(s.campos), __ldg(s.campos + 1), __ldg(s.campos + 2)};
						rgb = sh_to_rgb(s.D, &s_col[3 * threadIdx.x], p_orig, campos, clamped);
					} else {
						const float3 campos = {__ldg(s.campos), __ldg(s.campos + 1), __ldg(s.campos + 2)};
						rgb = sh_to_rgb(s.D, s.shs + (size_t)idx * s.M * 3, p_orig, campos, clamped);
					}
					out_radius = mr;
					my_tiles = area;
					my_vis = 1;
					rect_lo = rminx | (rminy << 16);
					rect_hi = rmaxx | (rmaxy << 16);
					depth_bits = __float_as_uint(depth);
					rec.q0 = make_float4(pix_x, pix_y, conic.x, conic.y);
					rec.q1 = make_float4(conic.z, __ldg(s.opacities + idx), depth, rgb.x);
					rec.q2 = make_float4(rgb.y, rgb.z, __uint_as_float(rect_lo), __uint_as_float(rect_hi));
				}
			}
		} else if (s.prefiltered) {
			__trap();   // auxiliary.h:156-160
		}
		radii[idx] = out_radius;
		if (n_touched) n_touched[idx] = 0;
		g.rec[idx] = rec;
		GaussAcc z;
		z.a0 = make_float4(0.f, 0.f, 0.f, 0.f); z.a1 = z.a0; z.a2 = z.a0; z.a3 = z.a0;
		g.acc[idx] = z;
		g.tiles_touched[idx] = my_tiles;
		g.clamped[idx] = (uint8_t)clamped;
	}
	GSR_PROBE(0, 2);
	// per-tile instance counts (integer REDs): small rectangles lane-parallel, large ones warp-cooperative
	// (two copies of the walk: the choice of the counter array stays out of its loop)
	if (FUSED_SCATTER || hist_smem) for_each_tile(my_tiles, rect_lo, rect_hi, s.grid_x, 0u, 0u, [&](uint32_t tile, uint32_t, uint32_t) { atomicAdd(&s_hist[tile], 1u); });
	else for_each_tile(my_tiles, rect_lo, rect_hi, s.grid_x, 0u, 0u, [&](uint32_t tile, uint32_t, uint32_t) { atomicAdd(&g.tile_count[tile], 1u); });
	GSR_PROBE(0, 3);
	// block totals -> header (integer atomics: deterministic); issued BEFORE the histogram flush so that one fence
	// covers every global update of this CTA
	unsigned wt = my_tiles, wv = my_vis;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		wt += __shfl_xor_sync(0xffffffffu, wt, o);
		wv += __shfl_xor_sync(0xffffffffu, wv, o);
	}
	if (lane_id() == 0) { s_red[threadIdx.x >> 5] = wt; s_red[8 + (threadIdx.x >> 5)] = wv; }
	__syncthreads();      // also: the CTA-private histogram is complete
	if (threadIdx.x == 0) {
		unsigned a = 0, b = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) { a += s_red[w]; b += s_red[8 + w]; }
		if (a) atomicAdd(&g.hdr->num_rendered, a);
		if (b) atomicAdd(&g.hdr->num_visible, b);
	}
	if (hist_smem) {
		// one coalesced RED per touched tile and CTA instead of one scattered RED per instance
		for (int t = threadIdx.x; t < n_tiles; t += 256) {
			const unsigned c = s_hist[t];
			if (c) atomicAdd(&g.tile_count[t], c);
		}
	}
	GSR_PROBE(0, 4);
	if (FUSED_SCATTER) {
		// ---- grid-wide barrier (single use per launch: the counter is cleared by the memset in front of the kernel) ----
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();      // one cumulative release behind the CTA barrier (as in cooperative-groups' grid sync)
			atomicAdd(&g.hdr->fwd_blocks_done, 1u);
			while (*reinterpret_cast<volatile unsigned*>(&g.hdr->fwd_blocks_done) < gridDim.x) { }
			__threadfence();
		}
		__syncthreads();
		pdl_launch_dependents();      // a forward compositing kernel launched as programmatic dependent may take its first slots
		GSR_PROBE(0, 5);
		// ---- claim this CTA's slice of every tile it touches (tile_cursor counts from 0: cleared by the memset).  The
		// atomics need nothing of the scan below, so they are issued first and their round trip overlaps the scan's loads;
		// the fused path takes at most kFusedMaxTiles tiles (fused_scatter_fits), i.e. kPer per thread ----
		constexpr int kPer = kFusedMaxTiles / 256;
		unsigned got[kPer];
#pragma unroll
		for (int k = 0; k < kPer; k++) {
			const int t = (int)threadIdx.x + 256 * k;
			const unsigned c = t < n_tiles ? s_hist[t] : 0u;
			got[k] = c ? atomicAdd(&g.tile_cursor[t], c) : 0u;
		}
		// ---- every CTA: exclusive scan of the tile counters -> s_start[tile]; CTA 0 also publishes ranges etc. ----
		uint32_t* s_start = s_hist + n_tiles;
		{
			const int per = (n_tiles + 255) / 256;
			const int t0 = min(n_tiles, (int)threadIdx.x * per), t1 = min(n_tiles, t0 + per);
			unsigned sum = 0;
			unsigned cnt[kPer];      // this thread's counters stay in registers for the second sweep
#pragma unroll
			for (int k = 0; k < kPer; k++) {
				cnt[k] = (t0 + k < t1) ? __ldcg(&g.tile_count[t0 + k]) : 0u;
				sum += cnt[k];
			}
			unsigned inc = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane_id() >= o) inc += v;
			}
			if (lane_id() == 31) s_red[threadIdx.x >> 5] = inc;
			__syncthreads();
			unsigned run = inc - sum;
			for (int w = 0; w < (int)(threadIdx.x >> 5); w++) run += s_red[w];
			unsigned mx = 0;
#pragma unroll
			for (int k = 0; k < kPer; k++) {
				const int t = t0 + k;
				if (t >= t1) break;
				const unsigned c = cnt[k];
				s_start[t] = run;
				if (blockIdx.x == 0) {
					g.ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
					if (c > (unsigned)GSR_SORT_CHUNK) g.long_tiles[atomicAdd(&g.hdr->num_long_tiles, 1u)] = (uint32_t)t;
					mx = max(mx, c);
				}
				run += c;
			}
			if (blockIdx.x == 0 && mx) atomicMax(&g.hdr->max_tile_count, mx);
		}
		__syncthreads();
#pragma unroll
		for (int k = 0; k < kPer; k++) {
			const int t = (int)threadIdx.x + 256 * k;
			if (t < n_tiles) { s_start[t] += got[k]; s_hist[t] = 0; }
		}
		__syncthreads();
		GSR_PROBE(0, 6);
		// ---- store the pairs ----
		bool overflow = false;
		for_each_tile(my_tiles, rect_lo, rect_hi, s.grid_x, depth_bits, (uint32_t)idx, [&](uint32_t tile, uint32_t key, uint32_t id) {
			const uint32_t pos = s_start[tile] + atomicAdd(&s_hist[tile], 1u);
			if (pos < fa.capacity) fa.pairs[pos] = make_uint2(key, id);
			else overflow = true;
		});
		if (overflow) g.hdr->overflow = 1;
		GSR_PROBE(0, 7);
		return;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();      // one cumulative release behind the CTA barrier
		s_last = (atomicAdd(&g.hdr->fwd_blocks_done, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	GSR_PROBE(0, 5);
	if (!s_last) return;
	// The last CTA to finish turns the tile counts into list ranges + scatter cursors (exclusive scan over
	// the tiles).  Untouched tiles keep the range (0,0) like the reference's memset (rasterizer_impl.cu:360).
	__threadfence();
	{
		const int tiles = n_tiles;
		const int per = (tiles + 255) / 256;
		const int t0 = min(tiles, (int)threadIdx.x * per), t1 = min(tiles, t0 + per);
		constexpr int kHold = 8;      // counts of up to 8 tiles per thread stay in registers between the two sweeps
		unsigned held[kHold];
		unsigned sum = 0;
#pragma unroll
		for (int u = 0; u < kHold; u++) held[u] = (t0 + u < t1) ? __ldcg(&g.tile_count[t0 + u]) : 0u;
#pragma unroll
		for (int u = 0; u < kHold; u++) sum += held[u];
		for (int t = t0 + kHold; t < t1; t++) sum += __ldcg(&g.tile_count[t]);
		unsigned inc = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane_id() >= o) inc += v;
		}
		__syncthreads();
		if (lane_id() == 31) s_red[threadIdx.x >> 5] = inc;
		__syncthreads();
		unsigned run = inc - sum;
		for (int w = 0; w < (int)(threadIdx.x >> 5); w++) run += s_red[w];
		unsigned mx = 0;
		for (int t = t0; t < t1; t++) {
			const unsigned c = (t - t0 < kHold) ? held[t - t0] : __ldcg(&g.tile_count[t]);
			g.ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
			g.tile_cursor[t] = run;
			if (c > (unsigned)GSR_SORT_CHUNK) g.long_tiles[atomicAdd(&g.hdr->num_long_tiles, 1u)] = (uint32_t)t;
			run += c;
			mx = max(mx, c);
		}
		if (mx) atomicMax(&g.hdr->max_tile_count, mx);
		if (threadIdx.x == 0) g.hdr->fwd_blocks_done = 0;
	}
	GSR_PROBE(0, 6);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ vm,
                                    unsigned char* __restrict__ present)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	const float x = means[3 * idx], y = means[3 * idx + 1], z = means[3 * idx + 2];
	const float pz = vm[2] * x + vm[6] * y + vm[10] * z + vm[14];
	present[idx] = !(pz <= 0.2f);
}

}  // namespace

static inline bool aligned16(const void* p) { return ((size_t)p & 15) == 0; }

// Can the cooperative preprocess + scatter kernel be used for P Gaussians on `tiles` tiles on the current device?
bool fused_scatter_fits(int P, int tiles)
{
	static const bool no_fuse = getenv("GSR_NO_FUSED_SCATTER") != nullptr;      // A/B switch for measurements
	if (no_fuse || P <= 0 || tiles > kFusedMaxTiles) return false;
	int dev = 0, sms = 0, per_sm = 0, coop = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, preprocess_forward_kernel<true>, 256, 2 * (size_t)tiles * sizeof(uint32_t));
	return coop && (long long)per_sm * sms >= (P + 255) / 256;
}

// Returns true when the kernel also built the per-tile segments (fused scatter): only attempted when the caller hands in
// the binning workspace (no-sync path), the CTA-private histogram fits and every CTA of the grid can be resident at once.
bool launch_preprocess_forward(const Scene& s, const GeomView& g, int* radii, int* n_touched, cudaStream_t stream,
                               const BinView* bin, size_t R_capacity)
{
	const int tiles = s.grid_x * s.grid_y;
	// header, the per-tile counters, the claim cursors and the tile-done flags behind it are cleared together
	cudaMemsetAsync(g.hdr, 0, sizeof(GeomHeader) + 3 * align_up((size_t)tiles * sizeof(uint32_t)), stream);
	if (s.P == 0) {
		cudaMemsetAsync(g.ranges, 0, (size_t)tiles * sizeof(uint2), stream);
		return false;
	}
	static const bool no_tma = getenv("GSR_NO_TMA") != nullptr;      // A/B switch for measurements
	int vec_mask = (aligned16(s.means3D) ? 1 : 0) | (aligned16(s.scales) ? 2 : 0) |
	               (aligned16(s.colors_precomp ? s.colors_precomp : s.shs) ? 4 : 0) | (no_tma ? 0 : 8);
	// CTA-private tile histogram in shared memory while it fits next to the static arrays (<= 8192 tiles, e.g. 1920x1080)
	const int hist_smem = tiles <= 8192 ? 1 : 0;
	const int grid = (s.P + 255) / 256;
	FusedScatterArgs fa;
	fa.pairs = bin ? bin->pairs : nullptr;
	fa.capacity = (unsigned)R_capacity;
	if (bin && R_capacity > 0 && fused_scatter_fits(s.P, tiles)) {
		const size_t smem = 2 * (size_t)tiles * sizeof(uint32_t);
		Scene sc = s;
		GeomView gv = g;
		int hs = 1;
		void* args[] = {&sc, &gv, &radii, &n_touched, &vec_mask, &hs, &fa};
		if (cudaLaunchCooperativeKernel((const void*)preprocess_forward_kernel<true>, dim3(grid), dim3(256), args, smem, stream) == cudaSuccess)
			return true;
		cudaGetLastError();      // fall back to the two-kernel path
	}
	preprocess_forward_kernel<false><<<grid, 256, hist_smem ? tiles * sizeof(uint32_t) : 0, stream>>>(s, g, radii, n_touched, vec_mask,
	                                                                                                 hist_smem, fa);
	return false;
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present, cudaStream_t stream)
{
	if (P == 0) return;
	mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
}

GSR_PROBE_READER(probe_read_preprocess)

}  // namespace gsr
