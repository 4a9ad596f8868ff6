// Per-tile compositing, forward and backward.
// Forward replaces reference renderCUDA (cuda_rasterizer/forward.cu:406-535); backward replaces
// renderCUDA (cuda_rasterizer/backward.cu:648-872).
//
// One CTA per 16x16 tile, one thread per pixel; each warp owns an 8x4 pixel block.  Batches of 256
// list entries are gathered as 48-byte records (3 x 16 B, cp.async / LDGSTS) into shared memory.
//
// Work skipping that does not change results: a (pixel, Gaussian) pair only matters when
// alpha = min(0.99, o*exp(-q)) >= 1/255, i.e. q <= ln(255 o).  For every chunk of 32 list entries the
// 32 lanes test one entry each against the warp's pixel block (exact minimum of the convex quadratic q
// over the 8x4 rectangle, with a conservative margin) and ballot; the warp then evaluates only the
// survivors (typically < 1/3 of the entries), with exactly the reference's per-pair arithmetic.
// Every per-pair decision is predicated, so the warp stays converged and can (a) vote its own early
// termination, (b) aggregate the n_touched integer atomics to one RED per warp, and in the backward
// (c) reduce the ten per-Gaussian gradient terms with a transposing butterfly of 12 register shuffles
// instead of the reference's 256-thread shared-memory tree (backward.cu:626-644, ~12 __syncthreads per
// (tile, Gaussian)), accumulate them per CTA in shared memory and flush them with three 16-byte vector
// REDs per (tile, Gaussian) instead of ten scalar global atomics (backward.cu:859-868).
#include "gsr_params.h"

namespace gsr {

namespace {

__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int& px, int& py)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	px = tile_x * GSR_TILE + (warp & 1) * 8 + (lane & 7);
	py = tile_y * GSR_TILE + (warp >> 1) * 4 + (lane >> 3);
}

// Conservative test: can the Gaussian (mean q0.xy, conic q0.z q0.w q1.x, opacity q1.y) reach
// alpha >= 1/255 at any pixel centre inside [x0,x1] x [y0,y1] ?  Returns false only when the exact
// per-pixel test (fp32, reference arithmetic) is guaranteed to reject every pixel of the block.
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
	__shared__ float4 s_q0[2][256], s_q1[2][256], s_q2[2][256];
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

	auto stage = [&](int b, int buf) {
		const int i = b * 256 + threadIdx.x;
		if (i < n) {
			const uint32_t id = __ldg(point_list + range.x + i);
			s_id[buf][threadIdx.x] = id;
			const GaussRec* r = rec + id;
			cp_async16(&s_q0[buf][threadIdx.x], &r->q0);
			cp_async16(&s_q1[buf][threadIdx.x], &r->q1);
			cp_async16(&s_q2[buf][threadIdx.x], &r->q2);
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
		bool warp_done = __all_sync(0xffffffffu, done);
		for (int c0 = 0; c0 < cnt && !warp_done; c0 += 32) {
			// cull phase: one list entry per lane
			const int e = c0 + lane;
			bool keep = false;
			if (e < cnt) keep = may_touch(s_q0[buf][e], s_q1[buf][e], bx0, by0, bx1, by1);
			unsigned live = __ballot_sync(0xffffffffu, keep);
			while (live) {
				const int j = c0 + __ffs(live) - 1;
				live &= live - 1;
				const float4 q0 = s_q0[buf][j];
				const float4 q1 = s_q1[buf][j];
				const float dx = q0.x - pxf, dy = q0.y - pyf;
				const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
				const float alpha = fminf(0.99f, q1.y * expf(power));
				bool valid = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				const float test_T = T * (1 - alpha);
				if (valid && test_T < 0.0001f) {
					done = true;
					valid = false;
				}
				if (__any_sync(0xffffffffu, valid)) {
					const float4 q2 = s_q2[buf][j];
					if (valid) {
						C0 += q1.w * alpha * T;
						C1 += q2.x * alpha * T;
						C2 += q2.y * alpha * T;
						D += q1.z * alpha * T;
						T = test_T;
						last_contributor = b * 256 + j + 1;
					}
					const unsigned touched = __ballot_sync(0xffffffffu, valid && test_T > 0.5f);
					if (touched && lane == 0) atomicAdd(&n_touched[s_id[buf][j]], __popc(touched));
				} else if (__all_sync(0xffffffffu, done)) {
					warp_done = true;
					break;
				}
			}
		}
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

// Sum ten per-lane values over the warp with a transposing butterfly: after the call lane L (even)
// holds the full sum of value index red_slot(L) (or nothing when red_slot(L) < 0).
// 12 shuffles instead of 50 for ten independent butterflies.
__device__ __forceinline__ int red_slot(int lane)
{
	// bit4: {0..4} | {5..9};  bit3: first three | last two (+pad);  bit2: first two | last one;  bit1: first | second
	const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
	int i3;                       // index among {0,1,2} after step B, -1 = padding
	if (!b2) i3 = b1;             // first two -> 0 or 1
	else i3 = b1 ? -1 : 2;        // last one -> 2 (second slot is padding)
	if (i3 < 0) return -1;
	int i5;                       // index among {0..4} after step A
	if (!b3) i5 = i3;             // first three: 0,1,2
	else { if (i3 == 2) return -1; i5 = 3 + i3; }   // last two: 3,4 (third slot is padding)
	return b4 * 5 + i5;
}

__device__ __forceinline__ float reduce10(const float v[10], int lane)
{
	const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
	float w[5], x[3], y[2];
#pragma unroll
	for (int k = 0; k < 5; k++) {
		const float send = b4 ? v[k] : v[k + 5];
		const float keep = b4 ? v[k + 5] : v[k];
		w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
	{
		const float s0 = b3 ? w[0] : w[3], k0 = b3 ? w[3] : w[0];
		const float s1 = b3 ? w[1] : w[4], k1 = b3 ? w[4] : w[1];
		const float s2 = b3 ? w[2] : 0.f, k2 = b3 ? 0.f : w[2];
		x[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
		x[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
		x[2] = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
	}
	{
		const float s0 = b2 ? x[0] : x[2], k0 = b2 ? x[2] : x[0];
		const float s1 = b2 ? x[1] : 0.f, k1 = b2 ? 0.f : x[1];
		y[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
		y[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
	}
	const float s = b1 ? y[0] : y[1], k = b1 ? y[1] : y[0];
	float z = k + __shfl_xor_sync(0xffffffffu, s, 2);
	z += __shfl_xor_sync(0xffffffffu, z, 1);
	return z;
}

__global__ void __launch_bounds__(256)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_depth,
                       GaussAcc* __restrict__ acc)
{
	__shared__ float4 s_q0[256], s_q1[256], s_q2[256];
	__shared__ uint32_t s_id[256];
	__shared__ float s_acc[256][12];
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
	const int my_slot = (lane & 1) ? -1 : red_slot(lane);
	// value order of v[]: 0 mean2D.x, 1 mean2D.y, 2 conic.xx, 3 conic.xy, 4 conic.yy, 5 opacity, 6 depth, 7..9 rgb

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

	float accum0 = 0.f, accum1 = 0.f, accum2 = 0.f, accumd = 0.f;
	float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_depth = 0.f;

	for (int hi = (int)top; hi > 0; hi -= 256) {
		const int cnt = min(256, hi);
		__syncthreads();   // previous batch fully flushed
		{
			const int t = threadIdx.x;
			if (t < cnt) {
				const uint32_t id = __ldg(point_list + range.x + (hi - 1 - t));
				s_id[t] = id;
				const GaussRec* r = rec + id;
				s_q0[t] = r->q0; s_q1[t] = r->q1; s_q2[t] = r->q2;
			}
#pragma unroll
			for (int q = 0; q < 12; q++) s_acc[t][q] = 0.f;
		}
		__syncthreads();
		// smem slot t holds list position hi-1-t: positions >= warp_top are skipped by this warp
		const int first = max(0, hi - (int)warp_top);
		for (int c0 = first & ~31; c0 < cnt; c0 += 32) {
			const int t = c0 + lane;
			bool keep = false;
			if (t >= first && t < cnt) keep = may_touch(s_q0[t], s_q1[t], bx0, by0, bx1, by1);
			unsigned live = __ballot_sync(0xffffffffu, keep);
			while (live) {
				const int j = c0 + __ffs(live) - 1;
				live &= live - 1;
				const uint32_t e = (uint32_t)(hi - 1 - j);   // 0-based position in the tile list
				const float4 q0 = s_q0[j];
				const float4 q1 = s_q1[j];
				const float dx = q0.x - pxf, dy = q0.y - pyf;
				const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
				const float G = expf(power);
				const float alpha = fminf(0.99f, q1.y * G);
				const bool valid = (e < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, valid)) continue;
				const float4 q2 = s_q2[j];
				float v[10];
#pragma unroll
				for (int q = 0; q < 10; q++) v[q] = 0.f;
				if (valid) {
					const float rcp = 1.f / (1.f - alpha);
					T = T * rcp;
					const float dchannel_dcolor = alpha * T;
					float dL_dalpha = 0.0f;
					accum0 = last_alpha * lc0 + (1.f - last_alpha) * accum0; lc0 = q1.w;
					dL_dalpha += (q1.w - accum0) * dp0;
					accum1 = last_alpha * lc1 + (1.f - last_alpha) * accum1; lc1 = q2.x;
					dL_dalpha += (q2.x - accum1) * dp1;
					accum2 = last_alpha * lc2 + (1.f - last_alpha) * accum2; lc2 = q2.y;
					dL_dalpha += (q2.y - accum2) * dp2;
					accumd = last_alpha * last_depth + (1.f - last_alpha) * accumd; last_depth = q1.z;
					dL_dalpha += (q1.z - accumd) * dpd;
					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += (-T_final * rcp) * bg_dot;
					const float dL_dG = q1.y * dL_dalpha;
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
					const float dG_ddely = -gdy * q1.x - gdx * q0.w;
					v[0] = dL_dG * dG_ddelx * ddelx_dx;
					v[1] = dL_dG * dG_ddely * ddely_dy;
					v[2] = -0.5f * gdx * dx * dL_dG;
					v[3] = -0.5f * gdx * dy * dL_dG;
					v[4] = -0.5f * gdy * dy * dL_dG;
					v[5] = G * dL_dalpha;
					v[6] = dchannel_dcolor * dpd;
					v[7] = dchannel_dcolor * dp0;
					v[8] = dchannel_dcolor * dp1;
					v[9] = dchannel_dcolor * dp2;
				}
				const float z = reduce10(v, lane);
				if (my_slot >= 0) atomicAdd(&s_acc[j][my_slot], z);
			}
		}
		__syncthreads();
		{
			const int t = threadIdx.x;
			if (t < cnt) {
				const float4 a0 = make_float4(s_acc[t][0], s_acc[t][1], s_acc[t][2], s_acc[t][3]);
				const float4 a1 = make_float4(s_acc[t][4], s_acc[t][5], s_acc[t][6], s_acc[t][7]);
				const float4 a2 = make_float4(s_acc[t][8], s_acc[t][9], 0.f, 0.f);
				GaussAcc* dst = acc + s_id[t];
				if (a0.x != 0.f || a0.y != 0.f || a0.z != 0.f || a0.w != 0.f) red_add_v4(&dst->a0, a0);
				if (a1.x != 0.f || a1.y != 0.f || a1.z != 0.f || a1.w != 0.f) red_add_v4(&dst->a1, a1);
				if (a2.x != 0.f || a2.y != 0.f) red_add_v4(&dst->a2, a2);
			}
		}
	}
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
