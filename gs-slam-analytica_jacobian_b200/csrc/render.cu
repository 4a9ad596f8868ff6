// Per-tile compositing, forward and backward.
// Forward replaces reference renderCUDA (cuda_rasterizer/forward.cu:406-535); backward replaces
// renderCUDA (cuda_rasterizer/backward.cu:648-872).
//
// One CTA per 16x16 tile, one thread per pixel; each warp owns an 8x4 pixel block.  Batches of 256
// list entries are gathered as 48-byte records (3 x 16 B, cp.async / LDGSTS, double buffered) into
// shared memory.
//
// Work skipping that does not change results: a (pixel, Gaussian) pair only matters when
// alpha = min(0.99, o*exp(-q)) >= 1/255, i.e. q <= ln(255 o).  For every chunk of 32 list entries the
// 32 lanes test one entry each against the warp's pixel block (exact minimum of the convex quadratic q
// over the 8x4 rectangle, with a conservative margin) and ballot; the warp then evaluates only the
// survivors (typically < 1/3 of the entries) with the reference's per-pair decisions.
// Every per-pair decision is predicated, so the warp stays converged and can (a) vote its own early
// termination, (b) aggregate the n_touched integer atomics to one RED per chunk and warp, and in the
// backward (c) reduce the ten per-Gaussian gradient terms with a transposing butterfly of 12 register
// shuffles instead of the reference's 256-thread shared-memory tree (backward.cu:626-644, ~12
// __syncthreads per (tile, Gaussian)) and flush them with four 16-byte vector REDs per (warp, Gaussian)
// instead of ten scalar global atomics behind a CTA-wide tree (backward.cu:859-868).
#include "gsr_params.h"

namespace gsr {

namespace {

#ifndef GSR_ACCURATE_EXP
// ex2.approx(x * log2 e): relative error ~2^-21 + |x| * 2^-23 (|x| < 6 where it matters); accurate expf costs 8
// more instructions per (pixel, Gaussian) pair in issue-bound kernels.
__device__ __forceinline__ float gsr_exp(float x) { return __expf(x); }
#else
__device__ __forceinline__ float gsr_exp(float x) { return expf(x); }
#endif

__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int& px, int& py)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	px = tile_x * GSR_TILE + (warp & 1) * 8 + (lane & 7);
	py = tile_y * GSR_TILE + (warp >> 1) * 4 + (lane >> 3);
}

// Conservative test: can the Gaussian (mean q0.xy, conic q0.z q0.w q1.x, opacity q1.y) reach
// alpha >= 1/255 at any pixel centre inside [x0,x1] x [y0,y1] ?  Returns false only when the exact
// per-pixel test (fp32) is guaranteed to reject every pixel of the block.
__device__ __forceinline__ bool may_touch(const float4 q0, const float4 q1, float x0, float y0, float x1, float y1)
{
	const float mx = q0.x, my = q0.y, A = q0.z, B = q0.w, Cc = q1.x, o = q1.y;
	const float cxp = fminf(fmaxf(mx, x0), x1), cyp = fminf(fmaxf(my, y0), y1);
	const float dx = mx - cxp, dy = my - cyp;   // 0 along an axis where the mean lies within the block
	float qmin = 0.f, S = 0.f;
	if (dx != 0.f || dy != 0.f) {
		// the constrained minimum of the convex quadratic lies on an edge facing the mean
		float qx = 3.0e38f, qy = 3.0e38f, Sx = 0.f, Sy = 0.f;
		if (dx != 0.f) {
			const float py = fminf(fmaxf(my + __fdividef(B * dx, Cc), y0), y1);
			const float e = my - py;
			const float t0 = 0.5f * (A * dx * dx + Cc * e * e), t1 = B * dx * e;
			qx = t0 + t1; Sx = t0 + fabsf(t1);
		}
		if (dy != 0.f) {
			const float px = fminf(fmaxf(mx + __fdividef(B * dy, A), x0), x1);
			const float e = mx - px;
			const float t0 = 0.5f * (A * e * e + Cc * dy * dy), t1 = B * e * dy;
			qy = t0 + t1; Sy = t0 + fabsf(t1);
		}
		if (qx < qy) { qmin = qx; S = Sx; } else { qmin = qy; S = Sy; }
	}
	const float tau = __logf(255.0f * o);                  // alpha >= 1/255  <=>  q <= ln(255 o)
	const float thr = tau + 1e-3f * fabsf(tau) + 1e-2f + 1e-5f * S;
	const bool convex = (A > 0.f) && (Cc > 0.f) && (A * Cc > B * B);
	return !(qmin > thr) || !convex;                      // NaNs fall through to "keep"
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
render_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_opacity, int* __restrict__ n_touched)
{
	__shared__ GaussRec s_rec[2][256];
	__shared__ uint32_t s_id[2][256];

	const int tile = blockIdx.x;
	const int tile_y = tile / grid_x, tile_x = tile - tile_y * grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int px, py;
	pixel_of_thread(tile_x, tile_y, px, py);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float bx0 = (float)(tile_x * GSR_TILE + (warp & 1) * 8), by0 = (float)(tile_y * GSR_TILE + (warp >> 1) * 4);
	const float bx1 = bx0 + 7.f, by1 = by0 + 3.f;
	const uint2 range = ranges[tile];
	const int n = (int)(range.y - range.x);
	const int rounds = (n + 255) / 256;

	bool done = !inside;
	float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
	uint32_t last_contributor = 0;
	// n_touched counts pixels whose transmittance after the blend is still > 0.5 (forward.cu:511-514).
	// T only decreases, so once no live lane of the warp is above 0.5 the bookkeeping is skipped for good.
	bool warp_hi_T = __any_sync(0xffffffffu, inside);

	auto stage = [&](int b, int buf) {
		const int i = b * 256 + threadIdx.x;
		if (i < n) {
			const uint32_t id = __ldg(point_list + range.x + i);
			s_id[buf][threadIdx.x] = id;
			const GaussRec* r = rec + id;
			cp_async16(&s_rec[buf][threadIdx.x].q0, &r->q0);
			cp_async16(&s_rec[buf][threadIdx.x].q1, &r->q1);
			cp_async16(&s_rec[buf][threadIdx.x].q2, &r->q2);
		}
		cp_async_commit();
	};
	if (rounds > 0) stage(0, 0);

	for (int b = 0; b < rounds; b++) {
		const int buf = b & 1;
		if (__syncthreads_and(done)) break;   // also: everyone is past batch b-1, buffer buf^1 is free
		if (b + 1 < rounds) stage(b + 1, buf ^ 1);
		else cp_async_commit();
		cp_async_wait<1>();
		__syncthreads();
		const int cnt = min(256, n - b * 256);
		int last_local = -1;
		for (int c0 = 0; c0 < cnt; c0 += 32) {
			if (__all_sync(0xffffffffu, done)) break;
			// cull phase: one list entry per lane
			const int e = c0 + lane;
			bool keep = false;
			if (e < cnt) keep = may_touch(s_rec[buf][e].q0, s_rec[buf][e].q1, bx0, by0, bx1, by1);
			unsigned live = __ballot_sync(0xffffffffu, keep);
			int my_touched = 0;   // lane L accumulates the n_touched increment of entry c0 + L
			while (live) {
				const int jl = __ffs(live) - 1;
				live &= live - 1;
				const GaussRec* r = &s_rec[buf][c0 + jl];
				const float4 q0 = r->q0;
				const float4 q1 = r->q1;
				const float4 q2 = r->q2;
				const float dx = q0.x - pxf, dy = q0.y - pyf;
				const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
				const float alpha = fminf(0.99f, q1.y * gsr_exp(power));
				bool valid = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				const float test_T = T * (1 - alpha);
				if (valid && test_T < 0.0001f) {
					done = true;
					valid = false;
				}
				if (warp_hi_T) {
					const unsigned touched = __ballot_sync(0xffffffffu, valid && test_T > 0.5f);
					if (lane == jl) my_touched = __popc(touched);
				}
				if (valid) {
					const float w = alpha * T;
					C0 += q1.w * w;
					C1 += q2.x * w;
					C2 += q2.y * w;
					D += q1.z * w;
					T = test_T;
					last_local = c0 + jl;
				}
			}
			if (warp_hi_T) {
				if (my_touched) atomicAdd(&n_touched[s_id[buf][e]], my_touched);
				warp_hi_T = __any_sync(0xffffffffu, !done && T > 0.5f);
			}
		}
		if (last_local >= 0) last_contributor = b * 256 + last_local + 1;
	}
	cp_async_wait<0>();
	if (inside) {
		const size_t pix = (size_t)W * py + px, HW = (size_t)H * W;
		final_T[pix] = T;
		n_contrib[pix] = last_contributor;
		out_color[pix] = C0 + T * bg[0];
		out_color[HW + pix] = C1 + T * bg[1];
		out_color[2 * HW + pix] = C2 + T * bg[2];
		out_depth[pix] = D;
		out_opacity[pix] = 1 - T;
	}
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------

// Sum ten per-lane values over the warp with a transposing butterfly (12 shuffles instead of 50 for ten
// independent butterflies), then gather them so that lane 8*g (g = 0..3) holds the g-th 16-byte word of
// the Gaussian's accumulator: g0 = {v0,v1,v2,0}  g1 = {v3,v4,0,0}  g2 = {v5,v6,v7,0}  g3 = {v8,v9,0,0}.
__device__ __forceinline__ float4 reduce10(const float v[10], int lane)
{
	const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
	float w[5], x[3], y[2];
#pragma unroll
	for (int k = 0; k < 5; k++) {           // halves: lanes 0-15 keep v0..4, lanes 16-31 keep v5..9
		const float send = b4 ? v[k] : v[k + 5];
		const float keep = b4 ? v[k + 5] : v[k];
		w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
	{                                       // quarters: first three | last two (+ padding)
		const float s0 = b3 ? w[0] : w[3], k0 = b3 ? w[3] : w[0];
		const float s1 = b3 ? w[1] : w[4], k1 = b3 ? w[4] : w[1];
		const float s2 = b3 ? w[2] : 0.f, k2 = b3 ? 0.f : w[2];
		x[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
		x[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
		x[2] = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
	}
	{                                       // eighths: first two | last one (+ padding)
		const float s0 = b2 ? x[0] : x[2], k0 = b2 ? x[2] : x[0];
		const float s1 = b2 ? x[1] : 0.f, k1 = b2 ? 0.f : x[1];
		y[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
		y[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
	}
	const float s = b1 ? y[0] : y[1], k = b1 ? y[1] : y[0];
	float z = k + __shfl_xor_sync(0xffffffffu, s, 2);
	z += __shfl_xor_sync(0xffffffffu, z, 1);
	// within each group of 8 lanes: offsets 0,1 hold X0; 2,3 hold X1; 4,5 hold X2 (or padding 0); 6,7 padding
	const float x1 = __shfl_down_sync(0xffffffffu, z, 2);
	const float x2 = __shfl_down_sync(0xffffffffu, z, 4);
	return make_float4(z, x1, x2, 0.f);
}

__global__ void __launch_bounds__(256)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_depth,
                       GaussAcc* __restrict__ acc)
{
	__shared__ GaussRec s_rec[2][256];
	__shared__ uint32_t s_id[2][256];
	__shared__ uint32_t s_max[8];

	const int tile = blockIdx.x;
	const int tile_y = tile / grid_x, tile_x = tile - tile_y * grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int px, py;
	pixel_of_thread(tile_x, tile_y, px, py);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float bx0 = (float)(tile_x * GSR_TILE + (warp & 1) * 8), by0 = (float)(tile_y * GSR_TILE + (warp >> 1) * 4);
	const float bx1 = bx0 + 7.f, by1 = by0 + 3.f;
	const uint2 range = ranges[tile];
	const size_t pix = (size_t)W * py + px, HW = (size_t)H * W;

	const float T_final = inside ? final_T[pix] : 0.f;
	float T = T_final;
	const uint32_t last_contributor = inside ? n_contrib[pix] : 0;
	float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f, dpd = 0.f;
	if (inside) {
		dp0 = dL_dpix[pix]; dp1 = dL_dpix[HW + pix]; dp2 = dL_dpix[2 * HW + pix];
		dpd = dL_dpix_depth[pix];
	}
	const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
	// value order of v[]: 0 mean2D.x, 1 mean2D.y, 2 conic.xx | 3 conic.xy, 4 conic.yy | 5 opacity, 6 depth, 7 red | 8 green, 9 blue

	// entries behind the tile's deepest contributor can never contribute (backward.cu:763)
	uint32_t m = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	if (lane == 0) s_max[warp] = m;
	__syncthreads();
	uint32_t top = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) top = max(top, s_max[w]);
	const uint32_t warp_top = m;   // this warp's deepest contributor

	// The reference keeps, per channel, accum_rec = last_alpha*last_color + (1-last_alpha)*accum_rec and sums
	// (c - accum_rec)*dL/dC over colour and depth (backward.cu:799-813).  The sum is linear, so ONE scalar
	// recurrence on beta = <accum_rec, dL/dpixel> with s = <(rgb,depth), dL/dpixel> is the same quantity.
	float beta = 0.f, last_alpha = 0.f, last_s = 0.f;

	const int rounds = ((int)top + 255) / 256;
	// batch b covers list positions [hi_b - cnt_b, hi_b), hi_b = top - 256 b; smem slot t <-> position hi_b-1-t
	auto stage = [&](int b, int buf) {
		const int hi = (int)top - b * 256;
		const int t = threadIdx.x;
		if (t < hi) {
			const uint32_t id = __ldg(point_list + range.x + (hi - 1 - t));
			s_id[buf][t] = id;
			const GaussRec* r = rec + id;
			cp_async16(&s_rec[buf][t].q0, &r->q0);
			cp_async16(&s_rec[buf][t].q1, &r->q1);
			cp_async16(&s_rec[buf][t].q2, &r->q2);
		}
		cp_async_commit();
	};
	if (rounds > 0) stage(0, 0);

	for (int b = 0; b < rounds; b++) {
		const int buf = b & 1;
		const int hi = (int)top - b * 256;
		const int cnt = min(256, hi);
		__syncthreads();   // everyone is past batch b-1: buffer buf^1 is free
		if (b + 1 < rounds) stage(b + 1, buf ^ 1);
		else cp_async_commit();
		cp_async_wait<1>();
		__syncthreads();
		// positions >= warp_top are skipped by this warp
		const int first = max(0, hi - (int)warp_top);
		for (int c0 = first & ~31; c0 < cnt; c0 += 32) {
			const int t = c0 + lane;
			bool keep = false;
			if (t >= first && t < cnt) keep = may_touch(s_rec[buf][t].q0, s_rec[buf][t].q1, bx0, by0, bx1, by1);
			unsigned live = __ballot_sync(0xffffffffu, keep);
			while (live) {
				const int j = c0 + __ffs(live) - 1;
				live &= live - 1;
				const uint32_t e = (uint32_t)(hi - 1 - j);   // 0-based position in the tile list
				const GaussRec* r = &s_rec[buf][j];
				const float4 q0 = r->q0;
				const float4 q1 = r->q1;
				const float dx = q0.x - pxf, dy = q0.y - pyf;
				const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
				const float G = gsr_exp(power);
				const float alpha = fminf(0.99f, q1.y * G);
				const bool valid = (e < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				const unsigned vmask = __ballot_sync(0xffffffffu, valid);
				if (vmask == 0) continue;
				const float4 q2 = r->q2;
				float v[10];
#pragma unroll
				for (int q = 0; q < 10; q++) v[q] = 0.f;
				if (valid) {
					const float rcp = __fdividef(1.f, 1.f - alpha);
					T = T * rcp;
					const float w = alpha * T;     // d(channel)/d(colour)
					const float sdot = q1.w * dp0 + q2.x * dp1 + q2.y * dp2 + q1.z * dpd;
					beta = last_alpha * last_s + (1.f - last_alpha) * beta;
					last_s = sdot;
					last_alpha = alpha;
					const float dL_dalpha = (sdot - beta) * T + (-T_final * rcp) * bg_dot;
					const float dL_dG = q1.y * dL_dalpha;
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
					const float dG_ddely = -gdy * q1.x - gdx * q0.w;
					const float hg = -0.5f * dL_dG;
					v[0] = dL_dG * dG_ddelx * ddelx_dx;
					v[1] = dL_dG * dG_ddely * ddely_dy;
					v[2] = hg * gdx * dx;
					v[3] = hg * gdx * dy;
					v[4] = hg * gdy * dy;
					v[5] = G * dL_dalpha;
					v[6] = w * dpd;
					v[7] = w * dp0;
					v[8] = w * dp1;
					v[9] = w * dp2;
				}
				GaussAcc* dst = acc + s_id[buf][j];
				if (__popc(vmask) <= 2) {
					// sparse pair: the one or two contributing lanes add their terms directly
					if (valid) {
						red_add_v4(&dst->a0, make_float4(v[0], v[1], v[2], 0.f));
						red_add_v4(&dst->a1, make_float4(v[3], v[4], 0.f, 0.f));
						red_add_v4(&dst->a2, make_float4(v[5], v[6], v[7], 0.f));
						red_add_v4(&dst->a3, make_float4(v[8], v[9], 0.f, 0.f));
					}
				} else {
					const float4 z = reduce10(v, lane);
					if ((lane & 7) == 0) red_add_v4(&dst->a0 + (lane >> 3), z);
				}
			}
		}
	}
	cp_async_wait<0>();
}

}  // namespace

void launch_render_forward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, float* out_color,
                           float* out_depth, float* out_opacity, int* n_touched, cudaStream_t stream)
{
	const int tiles = s.grid_x * s.grid_y;
	if (tiles == 0) return;
	render_forward_kernel<<<tiles, 256, 0, stream>>>(im.ranges, b.point_list, g.rec, s.W, s.H, s.grid_x, s.background,
	                                                  im.final_T, im.n_contrib, out_color, out_depth, out_opacity, n_touched);
}

void launch_render_backward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im,
                            const float* dL_dpix, const float* dL_dpix_depth, cudaStream_t stream)
{
	const int tiles = s.grid_x * s.grid_y;
	if (tiles == 0) return;
	render_backward_kernel<<<tiles, 256, 0, stream>>>(im.ranges, b.point_list, g.rec, s.W, s.H, s.grid_x, s.background,
	                                                   im.final_T, im.n_contrib, dL_dpix, dL_dpix_depth, g.acc);
}

}  // namespace gsr
