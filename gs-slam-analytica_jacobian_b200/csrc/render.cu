// Per-tile front-to-back compositing of colour / depth / opacity.
// Replaces reference renderCUDA (cuda_rasterizer/forward.cu:406-535).
//
// One CTA per 16x16 tile, one thread per pixel; each warp owns an 8x4 pixel block.  Batches of 256
// list entries are gathered as 48-byte records (3 x 16 B, cp.async / LDGSTS, double buffered) into
// shared memory.
//
// Work skipping that does not change results: a (pixel, Gaussian) pair only matters when
// alpha = min(0.99, o*exp(-q)) >= 1/255, i.e. q <= ln(255 o).  For every chunk of 32 list entries the
// 32 lanes test one entry each against the warp's pixel block (exact minimum of the convex quadratic q
// over the 8x4 rectangle, with a conservative margin), ballot, and COMPACT the survivors (typically < 1/3
// of the entries) in list order into a warp-private shared-memory queue.  The pixel threads then run a
// branch-free, two-way unrolled loop over the queue with the reference's per-pair decisions:
//   * every record is three broadcast LDS.128 at immediate offsets (no per-entry bit scan / address math),
//   * the exponent keeps the reference's expression tree (bit-identical power), exp is one ex2.approx.ftz,
//   * "done" is carried in the sign of T (T < 0 <=> this pixel stopped; |T| is its final transmittance), so the
//     three per-pair tests of the reference collapse into compares whose results are used as predicates,
//   * n_touched (pixels whose transmittance after the blend is still > 0.5, forward.cu:511-514) costs one vote +
//     one predicated store per entry while any pixel of the warp is above 0.5, nothing afterwards, and ONE
//     integer RED per (warp, Gaussian).
// The cull ballots are also written out (one word per warp and 32 list positions): the backward kernel replays them
// instead of repeating the test.
#include "render_common.cuh"
#include "tile_sort.cuh"

namespace gsr {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct FwdSmem {
	GaussRec rec[2][256];
	uint32_t id[2][256];
	QueueRec queue[8][34];      // per warp: <= 32 survivors of a chunk + one padding record for the 2-way unroll
	uint32_t tmask[8][32];      // per warp and queue slot: ballot of "still above 0.5 after this blend"
	uint64_t bar[2];            // GSR_TMA_STAGE: one mbarrier per staging buffer
};

template <bool COUNT_TOUCHED>
__device__ __forceinline__ void blend_queue(const QueueRec* __restrict__ q, int n, float pxf, float pyf, float& T, float& C0,
                                            float& C1, float& C2, float& D, int& last, uint32_t* __restrict__ tmask, int lane)
{
	for (int k = 0; k < n; k += 2) {
#pragma unroll
		for (int u = 0; u < 2; u++) {
			const QueueRec* r = q + k + u;
			const float4 w0 = r->w0;
			const float4 w1 = r->w1;
			const float4 w2 = r->w2;
			const float dx = w0.x - pxf, dy = w0.y - pyf;
			const float power = falloff_power(w0.z, w0.w, w1.x, dx, dy);
			const float alpha = fminf(0.99f, w1.y * gsr_exp(power));
			const float test_T = T * (1.0f - alpha);
			// reference order (forward.cu:481-507): skip if power > 0, skip if alpha < 1/255, stop if test_T < 1e-4.
			// A stopped pixel has T < 0, hence test_T < 0: it can only re-enter the "stop" arm, which is idempotent.
			const bool live = !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
			const bool stop = live && (test_T < 0.0001f);
			const bool valid = live && !(test_T < 0.0001f);
			if (COUNT_TOUCHED) {
				const unsigned touched = __ballot_sync(kFull, valid && test_T > 0.5f);
				if (lane == 0) tmask[k + u] = touched;
			}
			const float w = alpha * T;
			if (valid) {
				C0 += w1.z * w;
				C1 += w1.w * w;
				C2 += w2.x * w;
				D += w2.y * w;
				T = test_T;
				last = __float_as_int(w2.z);
			}
			if (stop) T = -fabsf(T);
		}
	}
}

// FUSED_SORT: the CTA first sorts its tile's scattered (depth, id) segment (tile_sort.cuh; lists of at most 2048 entries
// in shared memory, anything else through the general path) and composites straight from the sorted ids it keeps in
// shared memory.  One launch less, no point_list round trip before the first gather, and the latency-bound sort phases
// of one CTA overlap the issue-bound blending of the other CTAs on the SM.
struct FusedSortArgs {
	uint2* ranges;
	uint2* pairs;
	uint2* pairs_alt;
	uint32_t* point_list;
	unsigned capacity;
	int id_bits;
	GeomHeader* hdr;
};
constexpr int kFusedIdsOffset = 40960;    // bytes: behind the FwdSmem overlay, inside the sort's counter scratch
constexpr int kFusedIdsCap = GSR_SORT_CHUNK;

template <bool FUSED_SORT>
__global__ void __launch_bounds__(256, 4)
render_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* point_list,
                      const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_opacity, int* __restrict__ n_touched,
                      uint32_t* __restrict__ cull_masks, FusedSortArgs fs)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	static_assert(sizeof(FwdSmem) <= kFusedIdsOffset, "sorted ids must sit behind the compositing overlay");
	FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
	const uint32_t* s_ids = reinterpret_cast<const uint32_t*>(smem_raw + kFusedIdsOffset);
	if (FUSED_SORT) {
		sort_tile<256>((int)blockIdx.x, fs.ranges, fs.pairs, fs.pairs_alt, fs.point_list, fs.capacity, kSmallChunk, fs.id_bits,
		               fs.hdr, reinterpret_cast<uint32_t*>(smem_raw), reinterpret_cast<uint32_t*>(smem_raw + kFusedIdsOffset),
		               kFusedIdsCap);
		__threadfence_block();
		__syncthreads();      // sorted ids (shared + global) and a possibly clamped range are visible to the whole CTA
	}

	const int tile = blockIdx.x;
	const int tile_y = tile / grid_x, tile_x = tile - tile_y * grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	int px, py;
	pixel_of_thread(tile_x, tile_y, px, py);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float bx0 = (float)(tile_x * GSR_TILE + (warp & 1) * 8), by0 = (float)(tile_y * GSR_TILE + (warp >> 1) * 4);
	const float bx1 = bx0 + 7.f, by1 = by0 + 3.f;
	const uint2 range = FUSED_SORT ? __ldcg(&fs.ranges[tile]) : ranges[tile];
	const int n = (int)(range.y - range.x);
	const int rounds = (n + 255) / 256;
	const bool ids_in_smem = FUSED_SORT && n <= kFusedIdsCap;
	QueueRec* wq = sm.queue[warp];
	uint32_t* wmask = sm.tmask[warp];
	uint32_t* my_masks = cull_masks + cull_mask_base(range.x, (uint32_t)tile) + warp;   // [group of 32 positions][warp]

	float T = inside ? 1.0f : -1.0f;   // sign bit = "done"
	float C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
	int last = -1;                     // list position of the last blended entry
	// T only decreases, so once no live pixel of the warp is above 0.5 the n_touched bookkeeping is skipped for good.
	bool warp_hi_T = __any_sync(kFull, inside);

#ifdef GSR_TMA_STAGE
	// Staging through the TMA unit: every thread issues ONE 48-byte bulk copy (UBLKCP) for its record and arrives on the
	// buffer's mbarrier (thread 0 also registers the expected byte count); consumers wait on the barrier's phase, which
	// also publishes the ids written with ordinary stores -- no cp.async groups, one CTA barrier less per batch.
	if (threadIdx.x == 0) { mbar_init(&sm.bar[0], 256); mbar_init(&sm.bar[1], 256); mbar_fence_init(); }
	__syncthreads();
	auto stage = [&](int b, int buf) {
		const int i = b * 256 + threadIdx.x;
		const int cnt_b = min(256, n - b * 256);
		fence_proxy_async();      // earlier generic-proxy reads of this buffer are ordered before the async writes
		if (i < n) {
			const uint32_t id = ids_in_smem ? s_ids[i] : (FUSED_SORT ? __ldcg(point_list + range.x + i) : __ldg(point_list + range.x + i));
			sm.id[buf][threadIdx.x] = id;
			tma_bulk_g2s(&sm.rec[buf][threadIdx.x], rec + id, (unsigned)sizeof(GaussRec), &sm.bar[buf]);
		}
		if (threadIdx.x == 0) mbar_arrive_expect_tx(&sm.bar[buf], (unsigned)cnt_b * (unsigned)sizeof(GaussRec));
		else mbar_arrive(&sm.bar[buf]);
	};
	if (rounds > 0) stage(0, 0);

	int b = 0;
	for (; b < rounds; b++) {
		const int buf = b & 1;
		if (__syncthreads_and(T < 0.f)) break;   // also: everyone is past batch b-1, buffer buf^1 is free
		if (b + 1 < rounds) stage(b + 1, buf ^ 1);
		mbar_wait(&sm.bar[buf], (unsigned)((b >> 1) & 1));
#else
	auto stage = [&](int b, int buf) {
		const int i = b * 256 + threadIdx.x;
		if (i < n) {
			// fused: ids come from shared memory; a list too long for it was written by this CTA -> coherent load
			const uint32_t id = ids_in_smem ? s_ids[i] : (FUSED_SORT ? __ldcg(point_list + range.x + i) : __ldg(point_list + range.x + i));
			sm.id[buf][threadIdx.x] = id;
			const GaussRec* r = rec + id;
			cp_async16(&sm.rec[buf][threadIdx.x].q0, &r->q0);
			cp_async16(&sm.rec[buf][threadIdx.x].q1, &r->q1);
			cp_async16(&sm.rec[buf][threadIdx.x].q2, &r->q2);
		}
		cp_async_commit();
	};
	if (rounds > 0) stage(0, 0);

	for (int b = 0; b < rounds; b++) {
		const int buf = b & 1;
		if (__syncthreads_and(T < 0.f)) break;   // also: everyone is past batch b-1, buffer buf^1 is free
		if (b + 1 < rounds) stage(b + 1, buf ^ 1);
		else cp_async_commit();
		cp_async_wait<1>();
		__syncthreads();
#endif
		const int cnt = min(256, n - b * 256);
		uint32_t held_mask = 0;
		for (int c0 = 0; c0 < cnt; c0 += 32) {
			if (__all_sync(kFull, T < 0.f)) break;
			// cull phase: one list entry per lane
			const int e = c0 + lane;
			bool keep = false;
			float4 q0, q1;
			if (e < cnt) {
				q0 = sm.rec[buf][e].q0;
				q1 = sm.rec[buf][e].q1;
				keep = may_touch(q0, q1, bx0, by0, bx1, by1);
			}
			const unsigned mask = __ballot_sync(kFull, keep);
			if (lane == (c0 >> 5)) held_mask = mask;   // lane j keeps the ballot of chunk j; stored once per batch
			if (mask == 0) continue;
			const int nq = __popc(mask);
			const int pos = __popc(mask & lt);
			if (keep) {
				const float4 q2 = sm.rec[buf][e].q2;
				QueueRec* dst = wq + pos;
				dst->w0 = q0;
				dst->w1 = make_float4(q1.x, q1.y, q1.w, q2.x);
				dst->w2 = make_float4(q2.y, q1.z, __int_as_float(b * 256 + e), __uint_as_float(sm.id[buf][e]));
			}
			if (lane == 0) {   // padding record for an odd survivor count: opacity 0 -> alpha 0 -> skipped
				wq[nq].w0 = make_float4(0.f, 0.f, 0.f, 0.f);
				wq[nq].w1 = make_float4(0.f, 0.f, 0.f, 0.f);
			}
			__syncwarp();
			if (warp_hi_T) {
				blend_queue<true>(wq, nq, pxf, pyf, T, C0, C1, C2, D, last, wmask, lane);
				__syncwarp();
				if (keep) {
					const unsigned m = wmask[pos];
					if (m) atomicAdd(&n_touched[__float_as_uint(wq[pos].w2.w)], __popc(m));
				}
				warp_hi_T = __any_sync(kFull, T > 0.5f);
			} else {
				blend_queue<false>(wq, nq, pxf, pyf, T, C0, C1, C2, D, last, wmask, lane);
			}
			__syncwarp();   // queue fully consumed before the next chunk overwrites it
		}
		// the backward replays these ballots instead of repeating the cull (chunks this warp skipped stay 0: they lie
		// behind its last contributor and are never read)
		if (lane < 8 && (b * 8 + lane) * 32 < n) my_masks[(b * 8 + lane) * 8] = held_mask;   // only this tile's own groups
	}
#ifdef GSR_TMA_STAGE
	if (b < rounds) mbar_wait(&sm.bar[b & 1], (unsigned)((b >> 1) & 1));   // a staged batch nobody consumed: let it land before exit
#else
	cp_async_wait<0>();
#endif
	if (inside) {
		const size_t pix = (size_t)W * py + px, HW = (size_t)H * W;
		T = fabsf(T);
		final_T[pix] = T;
		n_contrib[pix] = (uint32_t)(last + 1);
		out_color[pix] = C0 + T * bg[0];
		out_color[HW + pix] = C1 + T * bg[1];
		out_color[2 * HW + pix] = C2 + T * bg[2];
		out_depth[pix] = D;
		out_opacity[pix] = 1 - T;
	}
}

}  // namespace

void launch_render_forward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, float* out_color,
                           float* out_depth, float* out_opacity, int* n_touched, bool fused_sort, size_t R_capacity,
                           cudaStream_t stream)
{
	const int tiles = s.grid_x * s.grid_y;
	if (tiles == 0) return;
	FusedSortArgs fs;
	fs.ranges = g.ranges; fs.pairs = b.pairs; fs.pairs_alt = b.pairs_alt; fs.point_list = b.point_list;
	fs.capacity = (unsigned)R_capacity; fs.hdr = g.hdr;
	fs.id_bits = 1;
	while (fs.id_bits < 32 && (1ll << fs.id_bits) < (long long)s.P) fs.id_bits++;
	const size_t smem_plain = sizeof(FwdSmem);
	const size_t smem_fused = sort_smem_bytes(kSmallChunk, 256) > (size_t)kFusedIdsOffset + kFusedIdsCap * 4
	                              ? sort_smem_bytes(kSmallChunk, 256) : (size_t)kFusedIdsOffset + kFusedIdsCap * 4;
	static SmemAttrCache fused_attr, plain_attr;
	ensure_dynamic_smem(render_forward_kernel<true>, smem_fused, fused_attr);
	ensure_dynamic_smem(render_forward_kernel<false>, smem_plain, plain_attr);
	if (fused_sort && s.P > 0 && R_capacity > 0)
		render_forward_kernel<true><<<tiles, 256, smem_fused, stream>>>(g.ranges, b.point_list, g.rec, s.W, s.H, s.grid_x, s.background,
		                                                                im.final_T, im.n_contrib, out_color, out_depth, out_opacity,
		                                                                n_touched, b.cull_masks, fs);
	else
		render_forward_kernel<false><<<tiles, 256, smem_plain, stream>>>(g.ranges, b.point_list, g.rec, s.W, s.H, s.grid_x, s.background,
		                                                                 im.final_T, im.n_contrib, out_color, out_depth, out_opacity,
		                                                                 n_touched, b.cull_masks, fs);
}

}  // namespace gsr
