// Per-tile back-to-front backward compositing.  Replaces reference renderCUDA
// (cuda_rasterizer/backward.cu:648-872).
//
// Same CTA / warp / pixel mapping, record staging and exact warp-level culling as the forward kernel.
// What differs from the reference is how the per-(pixel, Gaussian) gradient terms are summed over pixels.
// The reference reduces ten values per (tile, Gaussian) through a 256-thread shared-memory tree with ~12
// __syncthreads (backward.cu:626-644, 844-850) followed by ten scalar global atomics (:859-868).
// Here each warp:
//   1. evaluates a surviving pair with one thread per pixel and boils it down to TWO numbers per pixel,
//        ga = G * dL/dalpha      (every geometric gradient is ga times a polynomial in the pixel position)
//        w  = alpha * T          (every colour/depth gradient is w times the pixel's dL/dpixel)
//      which it drops into a warp-private shared-memory panel  [slot][pixel];
//   2. once NSLOT pairs are collected (or the batch ends) the lanes SWITCH ROLES: lane g now owns the
//      g-th collected Gaussian and walks the 32 pixels of the panel, accumulating in registers the six
//      moments  sum ga * {1, x, y, x^2, xy, y^2}  and the four sums  sum w * dL/d{r,g,b,depth}.
//      The per-pixel factors (pixel offsets: compile-time constants; dL/dpixel: one broadcast LDS.128) are
//      identical for all lanes, so this is a register-blocked [10 x 32] x [32 x NSLOT] product without any
//      cross-lane reduction: ~13 instructions per (pixel, 32 Gaussians) instead of ~60 instructions of
//      shuffles + selects per (warp, Gaussian);
//   3. the owner lane turns the moments into dL/dmean2D, dL/dconic, dL/dopacity (exact algebra, no
//      approximation) and issues four 16-byte vector REDs for its Gaussian.
// The reference's four per-channel "accum_rec" recurrences (backward.cu:799-813) are linear in dL/dpixel,
// so they are carried as ONE scalar recurrence on beta = <accum_rec, dL/dpixel>.
#include "render_common.cuh"

namespace gsr {

namespace {

#ifndef GSR_BWD_NSLOT
#define GSR_BWD_NSLOT 16
#endif
constexpr int kPanelStride = 33;   // floats per panel row: conflict-free for both access patterns

template <int NSLOT>
struct BwdSmem {
	GaussRec rec[2][256];
	uint32_t id[2][256];
	float4 dpix[8][32];                      // per warp: (dL/dr, dL/dg, dL/db, dL/ddepth) of its 32 pixels
	float panel_ga[8][NSLOT * kPanelStride];
	float panel_w[8][NSLOT * kPanelStride];
	uint32_t wmax[8];
};

// Role switch: lane (g, part) sums its Gaussian's panel row over 32/PARTS pixels, halves are combined,
// lane g < nslots converts moments to gradients and issues the REDs.
template <int NSLOT>
__device__ __forceinline__ void flush_panel(const float* __restrict__ pga, const float* __restrict__ pw,
                                            const float4* __restrict__ dpix, int nslots, int lane, const GaussRec* recs,
                                            const uint32_t* ids, int my_j, float bx0, float by0, float ddelx_dx,
                                            float ddely_dy, GaussAcc* __restrict__ acc)
{
	constexpr int PARTS = 32 / NSLOT, PPL = 32 / PARTS;
	const int g = lane % NSLOT, part = lane / NSLOT;
	__syncwarp();
	float H0 = 0.f, Hx = 0.f, Hy = 0.f, Hxx = 0.f, Hxy = 0.f, Hyy = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, cd = 0.f;
	const float* rga = pga + g * kPanelStride + part * PPL;
	const float* rw = pw + g * kPanelStride + part * PPL;
	const float4* rdp = dpix + part * PPL;
	const float ry0 = (float)(part * (PPL / 8));
#pragma unroll
	for (int k = 0; k < PPL; k++) {
		const float h = rga[k], w = rw[k];
		const float4 dp = rdp[k];
		const float rx = (float)(k & 7);             // pixel offset inside the warp's 8x4 block
		const float ry = ry0 + (float)(k >> 3);
		const float hx = h * rx, hy = h * ry;
		H0 += h; Hx += hx; Hy += hy;
		Hxx += hx * rx; Hxy += hx * ry; Hyy += hy * ry;
		c0 += w * dp.x; c1 += w * dp.y; c2 += w * dp.z; cd += w * dp.w;
	}
	if (PARTS == 2) {
		H0 += __shfl_xor_sync(0xffffffffu, H0, 16); Hx += __shfl_xor_sync(0xffffffffu, Hx, 16);
		Hy += __shfl_xor_sync(0xffffffffu, Hy, 16); Hxx += __shfl_xor_sync(0xffffffffu, Hxx, 16);
		Hxy += __shfl_xor_sync(0xffffffffu, Hxy, 16); Hyy += __shfl_xor_sync(0xffffffffu, Hyy, 16);
		c0 += __shfl_xor_sync(0xffffffffu, c0, 16); c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
		c2 += __shfl_xor_sync(0xffffffffu, c2, 16); cd += __shfl_xor_sync(0xffffffffu, cd, 16);
	}
	__syncwarp();   // panel fully consumed before the next collection overwrites it
	if (lane < nslots) {
		const GaussRec* r = recs + my_j;
		const float4 q0 = r->q0;
		const float4 q1 = r->q1;
		const float A = q0.z, B = q0.w, Cc = q1.x, o = q1.y;
		const float ux = q0.x - bx0, uy = q0.y - by0;    // mean relative to the block origin; d = u - r
		// sums of ga*dx, ga*dy, ga*dx^2, ga*dx*dy, ga*dy^2 from the moments
		const float Sx = ux * H0 - Hx, Sy = uy * H0 - Hy;
		const float Sxx = ux * (ux * H0 - 2.f * Hx) + Hxx;
		const float Syy = uy * (uy * H0 - 2.f * Hy) + Hyy;
		const float Sxy = ux * (uy * H0 - Hy) - uy * Hx + Hxy;
		// dL/dG = o * dL/dalpha; dG/dd = -G (A dx + B dy, C dy + B dx)   (backward.cu:831-842)
		const float m0 = -o * ddelx_dx * (A * Sx + B * Sy);
		const float m1 = -o * ddely_dy * (Cc * Sy + B * Sx);
		const float h = -0.5f * o;
		GaussAcc* dst = acc + ids[my_j];
		red_add_v4(&dst->a0, make_float4(m0, m1, h * Sxx, 0.f));
		red_add_v4(&dst->a1, make_float4(h * Sxy, h * Syy, 0.f, 0.f));
		red_add_v4(&dst->a2, make_float4(H0, cd, c0, 0.f));
		red_add_v4(&dst->a3, make_float4(c1, c2, 0.f, 0.f));
	}
}

template <int NSLOT>
__global__ void __launch_bounds__(256, 3)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_depth,
                       GaussAcc* __restrict__ acc)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	BwdSmem<NSLOT>& sm = *reinterpret_cast<BwdSmem<NSLOT>*>(smem_raw);

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
	sm.dpix[warp][lane] = make_float4(dp0, dp1, dp2, dpd);
	const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
	float* pga = sm.panel_ga[warp];
	float* pw = sm.panel_w[warp];

	// entries behind the tile's deepest contributor can never contribute (backward.cu:763)
	uint32_t m = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	if (lane == 0) sm.wmax[warp] = m;
	__syncthreads();
	uint32_t top = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) top = max(top, sm.wmax[w]);
	const uint32_t warp_top = m;   // this warp's deepest contributor

	float beta = 0.f, last_alpha = 0.f, last_s = 0.f;

	const int rounds = ((int)top + 255) / 256;
	// batch b covers list positions [hi_b - cnt_b, hi_b), hi_b = top - 256 b; smem slot t <-> position hi_b-1-t
	auto stage = [&](int b, int buf) {
		const int hi = (int)top - b * 256;
		const int t = threadIdx.x;
		if (t < hi) {
			const uint32_t id = __ldg(point_list + range.x + (hi - 1 - t));
			sm.id[buf][t] = id;
			const GaussRec* r = rec + id;
			cp_async16(&sm.rec[buf][t].q0, &r->q0);
			cp_async16(&sm.rec[buf][t].q1, &r->q1);
			cp_async16(&sm.rec[buf][t].q2, &r->q2);
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
		const int first = max(0, hi - (int)warp_top);   // positions >= warp_top are skipped by this warp
		int slot = 0, my_j = 0;
		for (int c0 = first & ~31; c0 < cnt; c0 += 32) {
			const int t = c0 + lane;
			bool keep = false;
			if (t >= first && t < cnt) keep = may_touch(sm.rec[buf][t].q0, sm.rec[buf][t].q1, bx0, by0, bx1, by1);
			unsigned live = __ballot_sync(0xffffffffu, keep);
			while (live) {
				const int j = c0 + __ffs(live) - 1;
				live &= live - 1;
				const uint32_t e = (uint32_t)(hi - 1 - j);   // 0-based position in the tile list
				const GaussRec* r = &sm.rec[buf][j];
				const float4 q0 = r->q0;
				const float4 q1 = r->q1;
				const float dx = q0.x - pxf, dy = q0.y - pyf;
				const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
				const float G = gsr_exp(power);
				const float alpha = fminf(0.99f, q1.y * G);
				const bool valid = (e < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, valid)) continue;
				float ga = 0.f, w = 0.f;
				if (valid) {
					const float4 q2 = r->q2;
					const float rcp = __fdividef(1.f, 1.f - alpha);
					T = T * rcp;
					w = alpha * T;     // d(channel)/d(colour)
					const float sdot = q1.w * dp0 + q2.x * dp1 + q2.y * dp2 + q1.z * dpd;
					beta = last_alpha * last_s + (1.f - last_alpha) * beta;
					last_s = sdot;
					last_alpha = alpha;
					const float dL_dalpha = (sdot - beta) * T + (-T_final * rcp) * bg_dot;
					ga = G * dL_dalpha;
				}
				pga[slot * kPanelStride + lane] = ga;
				pw[slot * kPanelStride + lane] = w;
				if ((lane % NSLOT) == slot) my_j = j;
				if (++slot == NSLOT) {
					flush_panel<NSLOT>(pga, pw, sm.dpix[warp], NSLOT, lane, sm.rec[buf], sm.id[buf], my_j, bx0, by0, ddelx_dx,
					                   ddely_dy, acc);
					slot = 0;
				}
			}
		}
		if (slot > 0)
			flush_panel<NSLOT>(pga, pw, sm.dpix[warp], slot, lane, sm.rec[buf], sm.id[buf], my_j, bx0, by0, ddelx_dx, ddely_dy, acc);
	}
	cp_async_wait<0>();
}

}  // namespace

void launch_render_backward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im,
                            const float* dL_dpix, const float* dL_dpix_depth, cudaStream_t stream)
{
	const int tiles = s.grid_x * s.grid_y;
	if (tiles == 0) return;
	constexpr int NSLOT = GSR_BWD_NSLOT;
	const size_t smem = sizeof(BwdSmem<NSLOT>);
	static bool configured = false;   // idempotent attribute, set once per process
	if (!configured) {
		cudaFuncSetAttribute(render_backward_kernel<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = true;
	}
	render_backward_kernel<NSLOT><<<tiles, 256, smem, stream>>>(g.ranges, b.point_list, g.rec, s.W, s.H, s.grid_x, s.background,
	                                                            im.final_T, im.n_contrib, dL_dpix, dL_dpix_depth, g.acc);
}

}  // namespace gsr
