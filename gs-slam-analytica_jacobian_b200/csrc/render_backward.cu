// Per-tile back-to-front backward compositing.  Replaces reference renderCUDA
// (cuda_rasterizer/backward.cu:648-872).
//
// Same CTA / warp / pixel mapping, record staging and exact warp-level culling as the forward kernel.
// What differs from the reference is how the per-(pixel, Gaussian) gradient terms are summed over pixels.
// The reference reduces ten values per (tile, Gaussian) through a 256-thread shared-memory tree with ~12
// __syncthreads (backward.cu:626-644, 844-850) followed by ten scalar global atomics (:859-868).
// Here each warp:
//   0. replays the forward's cull ballots (one word per warp and 32 list positions: the exact test of an entry
//      against the warp's 8x4 pixel block is not repeated) and appends the survivors, in back-to-front order, to a
//      warp-private ring queue in shared memory;
//   1. whenever 16 entries are queued, evaluates them with one thread per pixel in a branch-free two-way
//      unrolled loop (three broadcast LDS.128 per entry at immediate offsets) and boils
//      every pair down to TWO numbers per pixel,
//        ga = G * dL/dalpha      (every geometric gradient is ga times a polynomial in the pixel position)
//        w  = alpha * T          (every colour/depth gradient is w times the pixel's dL/dpixel)
//      which it drops into a warp-private shared-memory panel  [slot][pixel];
//   2. then the lanes SWITCH ROLES: lane g now owns the g-th queued Gaussian and walks the pixels of the
//      panel, accumulating in registers the six moments  sum ga * {1, x, y, x^2, xy, y^2}  and the four sums
//      sum w * dL/d{r,g,b,depth}.  The per-pixel factors (pixel offsets: compile-time constants; dL/dpixel:
//      one broadcast LDS.128) are identical for all lanes, so this is a register-blocked [10 x 32] x [32 x 16]
//      product without any cross-lane reduction;
//   3. the owner lane turns the moments into dL/dmean2D, dL/dconic, dL/dopacity (exact algebra, no
//      approximation) and issues four 16-byte vector REDs for its Gaussian.
// The reference's four per-channel "accum_rec" recurrences (backward.cu:799-813) are linear in dL/dpixel,
// so they are carried as ONE scalar recurrence on beta = <accum_rec, dL/dpixel>.
#include "render_common.cuh"

namespace gsr {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSlots = 16;         // Gaussians per role switch
constexpr int kPanelStride = 33;   // floats per panel row: conflict-free for both access patterns
constexpr int kBatch = 128;        // list entries staged per round
constexpr int kQueueCap = 48;      // ring: <= 15 pending + <= 32 appended per cull step

struct BwdSmem {
	GaussRec rec[2][kBatch];
	uint32_t id[2][kBatch];
	float4 dpix[8][32];                       // per warp: (dL/dr, dL/dg, dL/db, dL/ddepth) of its 32 pixels
	float2 panel[8][kSlots * kPanelStride];   // per warp: [slot][pixel] (ga, w)
	QueueRec queue[8][kQueueCap];
	uint32_t wmax[8];
};

struct PixelState {
	float pxf, pyf;
	float dp0, dp1, dp2, dpd, c_bg;
	int last_contributor;
	float T, beta;
};

// One thread per pixel: entries grp[0 .. cnt) (cnt even, a padding record with opacity 0 at the end if needed)
// EXACT: G from the reference's expf (backward.cu:772), so that alpha -- and with it every skip decision -- is bit-identical to
// the (exact) forward's and T = T / (1 - alpha) retraces the forward's transmittances to an ulp or two per step (the quotient
// itself as T * rcp.approx(1 - alpha): a correctly rounded division, -DGSR_BWD_DIV=1, was measured -- gradients 4e-7 instead
// of 7e-7 from the reference at 300 k Gaussians, compositing backward 126 -> 157 us at C1 -- and is not worth it).
// Otherwise ex2.approx (6 instructions less per pair, alpha within ~5e-7 relative, which the chain of divisions amplifies:
// gradients 1e-4 .. 2e-4 from the reference at 300 k Gaussians instead of 1e-6, tools/grad_noise.py).
template <bool EXACT>
__device__ __forceinline__ void eval_group(const QueueRec* __restrict__ grp, int cnt, float2* __restrict__ panel, PixelState& s)
{
	for (int k = 0; k < cnt; k += 2) {
#pragma unroll
		for (int u = 0; u < 2; u++) {
			const QueueRec* r = grp + k + u;
			const float4 w0 = r->w0;
			const float4 w1 = r->w1;
			const float4 w2 = r->w2;
			const float dx = w0.x - s.pxf, dy = w0.y - s.pyf;
			const float power = falloff_power(w0.z, w0.w, w1.x, dx, dy);
			const float G = EXACT ? expf(power) : gsr_exp(power);
			const float alpha = fminf(0.99f, w1.y * G);
			// backward.cu:763-783: only entries in front of the pixel's last contributor, same skips as the forward
			const bool valid = (__float_as_int(w2.z) < s.last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
			float ga = 0.f, w = 0.f;
			if (valid) {
#ifndef GSR_BWD_DIV
#define GSR_BWD_DIV 0
#endif
				const float rcp = (EXACT && GSR_BWD_DIV) ? __frcp_rn(1.f - alpha) : rcp_approx(1.f - alpha);   // 1 - alpha >= 0.01
				s.T = (EXACT && GSR_BWD_DIV) ? __fdiv_rn(s.T, 1.f - alpha) : s.T * rcp;
				w = alpha * s.T;     // d(channel)/d(colour)
				const float sdot = w1.z * s.dp0 + w1.w * s.dp1 + w2.x * s.dp2 + w2.y * s.dpd;
				// beta = <accum_rec, dL/dpixel> of the entries behind this one.  The reference updates accum_rec when it
				// reaches the NEXT entry, as last_alpha * last_color + (1 - last_alpha) * accum_rec (backward.cu:799-813);
				// the same value is beta + alpha * (sdot - beta), and (sdot - beta) is needed for dL/dalpha anyway.
				const float d = sdot - s.beta;
				const float dL_dalpha = d * s.T + s.c_bg * rcp;       // c_bg = -T_final * <bg, dL/dcolour>   (:823-828)
				s.beta = fmaf(alpha, d, s.beta);
				ga = G * dL_dalpha;
			}
			panel[(k + u) * kPanelStride] = make_float2(ga, w);
		}
	}
}

// Role switch: lane (g, part) sums its Gaussian's panel row over 16 pixels (two rows of the 8x4 block), the halves
// are combined, lane g < nslots converts moments to gradients and issues the REDs.  The moments are accumulated
// separably: per row  R0 = sum h, R1 = sum h x, R2 = sum h x^2  (x = 0..7 immediates), then folded with the row's y.
__device__ __forceinline__ void flush_panel(const float2* __restrict__ panel, const float4* __restrict__ dpix, int nslots,
                                            int lane, const QueueRec* __restrict__ grp, float bx0, float by0, float ddelx_dx,
                                            float ddely_dy, GaussAcc* __restrict__ acc)
{
	static_assert(kSlots == 16, "two lanes per Gaussian");
	const int g = lane & 15, part = lane >> 4;
	float H0 = 0.f, Hx = 0.f, Hy = 0.f, Hxx = 0.f, Hxy = 0.f, Hyy = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, cd = 0.f;
	const float2* row = panel + g * kPanelStride + part * 16;
	const float4* rdp = dpix + part * 16;
#pragma unroll
	for (int r = 0; r < 2; r++) {
		float R0 = 0.f, R1 = 0.f, R2 = 0.f;
#pragma unroll
		for (int x = 0; x < 8; x++) {
			const float2 hw = row[r * 8 + x];
			const float4 dp = rdp[r * 8 + x];
			R0 += hw.x;
			R1 = fmaf(hw.x, (float)x, R1);
			R2 = fmaf(hw.x, (float)(x * x), R2);
			c0 = fmaf(hw.y, dp.x, c0); c1 = fmaf(hw.y, dp.y, c1); c2 = fmaf(hw.y, dp.z, c2); cd = fmaf(hw.y, dp.w, cd);
		}
		const float ry = (float)(part * 2 + r);   // row inside the warp's 8x4 block
		H0 += R0; Hx += R1; Hxx += R2;
		Hy = fmaf(R0, ry, Hy); Hxy = fmaf(R1, ry, Hxy); Hyy = fmaf(R0, ry * ry, Hyy);
	}
	H0 += __shfl_xor_sync(kFull, H0, 16); Hx += __shfl_xor_sync(kFull, Hx, 16);
	Hy += __shfl_xor_sync(kFull, Hy, 16); Hxx += __shfl_xor_sync(kFull, Hxx, 16);
	Hxy += __shfl_xor_sync(kFull, Hxy, 16); Hyy += __shfl_xor_sync(kFull, Hyy, 16);
	c0 += __shfl_xor_sync(kFull, c0, 16); c1 += __shfl_xor_sync(kFull, c1, 16);
	c2 += __shfl_xor_sync(kFull, c2, 16); cd += __shfl_xor_sync(kFull, cd, 16);
	if (lane < nslots) {
		const QueueRec* r = grp + lane;
		const float4 w0 = r->w0;
		const float4 w1 = r->w1;
		const float A = w0.z, B = w0.w, Cc = w1.x, o = w1.y;
		const float ux = w0.x - bx0, uy = w0.y - by0;         // mean relative to the block origin; d = u - r
		// sums of ga*dx, ga*dy, ga*dx^2, ga*dx*dy, ga*dy^2 from the moments
		const float Sx = ux * H0 - Hx, Sy = uy * H0 - Hy;
		const float Sxx = ux * (ux * H0 - 2.f * Hx) + Hxx;
		const float Syy = uy * (uy * H0 - 2.f * Hy) + Hyy;
		const float Sxy = ux * (uy * H0 - Hy) - uy * Hx + Hxy;
		// dL/dG = o * dL/dalpha; dG/dd = -G (A dx + B dy, C dy + B dx)   (backward.cu:831-842)
		const float m0 = -o * ddelx_dx * (A * Sx + B * Sy);
		const float m1 = -o * ddely_dy * (Cc * Sy + B * Sx);
		const float h = -0.5f * o;
		GaussAcc* dst = acc + __float_as_uint(r->w2.w);
		red_add_v4(&dst->a0, make_float4(m0, m1, h * Sxx, 0.f));
		red_add_v4(&dst->a1, make_float4(h * Sxy, h * Syy, 0.f, 0.f));
		red_add_v4(&dst->a2, make_float4(H0, cd, c0, 0.f));
		red_add_v4(&dst->a3, make_float4(c1, c2, 0.f, 0.f));
	}
}

template <bool EXACT>
__global__ void __launch_bounds__(256, 3)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_depth,
                       GaussAcc* __restrict__ acc, const uint32_t* __restrict__ cull_masks, const uint32_t* __restrict__ tile_done,
                       const uint32_t* __restrict__ upstream_ready, GeomHeader* __restrict__ hdr, int tile0)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);

	const int tile = tile0 + (int)blockIdx.x;      // tile0: first tile of the band this launch covers
	GSR_PROBE(4, 0);
	pdl_launch_dependents();      // the per-Gaussian backward may move in (and fetch its inputs) during this kernel's tail
	// Launched as a programmatic dependent of the forward compositing kernel (RasterEngine.step) this CTA may be running
	// while forward CTAs of other tiles still are: wait for ITS tile (a flag the forward CTA releases behind its last
	// write; always set in a normally ordered launch) and read what the forward wrote with L2-coherent loads only.
	// upstream_ready (optional): a word that whoever delivers dL/dpixel sets behind the data (e.g. the last of the host-to-device
	// copies of a step on another stream): the dependency on it then need not be a full edge in front of this kernel.
	// The spins are bounded (~1 s each, separate budgets): a flag that never comes is reported through the header's
	// spin_timeout word (1 = tile flag, 2 = upstream word; gsr_step_status / GSR_ERR_TIMEOUT), not by hanging the GPU; the
	// tile then contributes no gradients.
	// The flag's value is the tile's deepest contributor + 1 (the forward CTA knows it): how far the list has to be walked
	// (backward.cu:763) is known here, and the first gather is on its way before this CTA's own n_contrib loads are back.
	if (threadIdx.x == 0) {
		int spins = 0;
		uint32_t v;
		while ((v = ld_acquire_u32(tile_done + tile)) == 0 && ++spins < (1 << 22)) __nanosleep(200);
		if (v == 0) hdr->spin_timeout = 1;
		if (upstream_ready) {
			spins = 0;
			uint32_t u;
			while ((u = ld_acquire_u32(upstream_ready)) == 0 && ++spins < (1 << 22)) __nanosleep(200);
			if (u == 0) { hdr->spin_timeout = 2; v = 0; }
		}
		sm.wmax[0] = v ? v - 1u : 0u;
	}
	__syncthreads();
	GSR_PROBE(4, 1);      // the tile's flag (and the upstream word) are in
	const uint32_t top = sm.wmax[0];
	const int tile_y = tile / grid_x, tile_x = tile - tile_y * grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	int px, py;
	pixel_of_thread(tile_x, tile_y, px, py);
	const bool inside = px < W && py < H;
	const float bx0 = (float)(tile_x * GSR_TILE + (warp & 1) * 8), by0 = (float)(tile_y * GSR_TILE + (warp >> 1) * 4);
	const uint2 range = __ldcg(ranges + tile);
	const size_t pix = (size_t)W * py + px, HW = (size_t)H * W;

	PixelState s;
	s.pxf = (float)px; s.pyf = (float)py;
	const float T_final = inside ? __ldcg(final_T + pix) : 0.f;
	s.T = T_final;
	s.last_contributor = inside ? (int)__ldcg(n_contrib + pix) : 0;
	s.dp0 = s.dp1 = s.dp2 = s.dpd = 0.f;
	if (inside) {
		// (L2-coherent: with the loss fused into the forward these are products of the forward CTA of this tile too)
		s.dp0 = __ldcg(dL_dpix + pix); s.dp1 = __ldcg(dL_dpix + HW + pix); s.dp2 = __ldcg(dL_dpix + 2 * HW + pix);
		s.dpd = __ldcg(dL_dpix_depth + pix);
	}
	// The list is walked back to front in groups of 32 positions aligned like the forward's cull chunks, so that the
	// forward's ballots (cull_masks) can be replayed: batch b covers positions [hi_b - kBatch, hi_b) with
	// hi_b = top32 - kBatch b (a multiple of 32); smem slot t <-> position hi_b - 1 - t.
	const int top32 = ((int)top + 31) & ~31;
	const int rounds = (top32 + kBatch - 1) / kBatch;
	const uint32_t* my_masks = cull_masks + cull_mask_base(range.x, (uint32_t)tile) + warp;
	auto stage = [&](int b, int buf) {
		const int hi = top32 - b * kBatch;
		for (int i = threadIdx.x; i < 3 * kBatch; i += 256) {
			const int t = i / 3, part = i - 3 * t;
			const int pos = hi - 1 - t;
			if (pos >= 0 && pos < (int)top) {
				const uint32_t id = __ldcg(point_list + range.x + pos);
				if (part == 0) sm.id[buf][t] = id;
				cp_async16(&sm.rec[buf][t].q0 + part, &rec[id].q0 + part);
			}
		}
		cp_async_commit();
	};
	if (rounds > 0) stage(0, 0);
	// (the first consumers of the per-pixel loads come behind the first gather's requests)
	sm.dpix[warp][lane] = make_float4(s.dp0, s.dp1, s.dp2, s.dpd);
	s.c_bg = -T_final * (bg[0] * s.dp0 + bg[1] * s.dp1 + bg[2] * s.dp2);
	s.beta = 0.f;
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
	float2* panel = sm.panel[warp];
	QueueRec* wq = sm.queue[warp];

	// entries behind the deepest contributor of the tile (top) / of this warp's pixels can never contribute (backward.cu:763)
	const uint32_t warp_top = __reduce_max_sync(kFull, (uint32_t)s.last_contributor);


	int head = 0, qn = 0;   // ring queue: qn entries starting at slot head (head is a multiple of kSlots)
	for (int b = 0; b < rounds; b++) {
		const int buf = b & 1;
		const int hi = top32 - b * kBatch;
		const int cnt = min(kBatch, hi);
		__syncthreads();   // everyone is past batch b-1: buffer buf^1 is free
		if (b + 1 < rounds) stage(b + 1, buf ^ 1);
		else cp_async_commit();
		// this warp's cull ballots of the (up to) four groups of the batch: lane j fetches group hi/32 - 1 - j
		uint32_t fetched = 0;
		{
			const int g = (hi >> 5) - 1 - lane;
			if (lane < kBatch / 32 && g >= 0 && g * 32 < (int)warp_top) fetched = __ldcg(my_masks + (size_t)g * 8);
		}
		cp_async_wait<1>();
		__syncthreads();
		if (b == 0) GSR_PROBE(4, 2);      // per-pixel state loaded, first batch of records staged: the prologue a fused kernel would not have
		for (int c0 = 0; c0 < cnt; c0 += 32) {
			// lane L <-> slot c0 + L <-> position p = hi - 1 - c0 - L = 32 g + (31 - L)
			const uint32_t fm = __shfl_sync(kFull, fetched, c0 >> 5);
			const int p = hi - 1 - c0 - lane;
			const bool keep = ((fm >> (31 - lane)) & 1u) && p < (int)warp_top;   // positions >= warp_top cannot contribute
			const unsigned mask = __ballot_sync(kFull, keep);
			if (mask == 0) continue;
			if (keep) {
				const int t = c0 + lane;
				const float4 q0 = sm.rec[buf][t].q0;
				const float4 q1 = sm.rec[buf][t].q1;
				const float4 q2 = sm.rec[buf][t].q2;
				int slot = head + qn + __popc(mask & lt);
				if (slot >= kQueueCap) slot -= kQueueCap;
				QueueRec* dst = wq + slot;
				dst->w0 = q0;
				dst->w1 = make_float4(q1.x, q1.y, q1.w, q2.x);
				dst->w2 = make_float4(q2.y, q1.z, __int_as_float(p), __uint_as_float(sm.id[buf][t]));
			}
			qn += __popc(mask);
			__syncwarp();
			while (qn >= kSlots) {
				eval_group<EXACT>(wq + head, kSlots, panel + lane, s);
				__syncwarp();
				flush_panel(panel, sm.dpix[warp], kSlots, lane, wq + head, bx0, by0, ddelx_dx, ddely_dy, acc);
				__syncwarp();   // panel and queue group consumed
				head = (head + kSlots == kQueueCap) ? 0 : head + kSlots;
				qn -= kSlots;
			}
		}
	}
	if (qn > 0) {
		if ((qn & 1) && lane == 0) {   // padding record: opacity 0 -> alpha 0 -> contributes nothing
			wq[head + qn].w0 = make_float4(0.f, 0.f, 0.f, 0.f);
			wq[head + qn].w1 = make_float4(0.f, 0.f, 0.f, 0.f);
			wq[head + qn].w2 = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		__syncwarp();
		eval_group<EXACT>(wq + head, (qn + 1) & ~1, panel + lane, s);
		__syncwarp();
		flush_panel(panel, sm.dpix[warp], qn, lane, wq + head, bx0, by0, ddelx_dx, ddely_dy, acc);
	}
	cp_async_wait<0>();
	GSR_PROBE(4, 3);
}

}  // namespace

void launch_render_backward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im,
                            const float* dL_dpix, const float* dL_dpix_depth, bool overlap_forward, cudaStream_t stream)
{
	const bool band = s.band_y1 > 0;
	const int tiles = band ? s.grid_x * (s.band_y1 - s.band_y0) : s.grid_x * s.grid_y;
	if (tiles <= 0) return;
	const size_t smem = sizeof(BwdSmem);
	static SmemAttrCache attr[2];
	ensure_dynamic_smem(render_backward_kernel<false>, smem, attr[0]);
	ensure_dynamic_smem(render_backward_kernel<true>, smem, attr[1]);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(tiles); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = overlap_forward ? 1 : 0;
	auto kernel = s.exact_exp_bwd ? render_backward_kernel<true> : render_backward_kernel<false>;
	cudaLaunchKernelEx(&cfg, kernel, (const uint2*)g.ranges, (const uint32_t*)b.point_list, (const GaussRec*)g.rec, s.W, s.H,
	                   s.grid_x, s.background, (const float*)im.final_T, (const uint32_t*)im.n_contrib, dL_dpix, dL_dpix_depth, g.acc,
	                   (const uint32_t*)b.cull_masks, (const uint32_t*)g.tile_done, (const uint32_t*)s.upstream_ready, g.hdr,
	                   band ? s.band_y0 * s.grid_x : 0);
}

GSR_PROBE_READER(probe_read_render_backward)

}  // namespace gsr
