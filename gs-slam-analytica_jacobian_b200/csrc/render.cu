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
//   * the exponent keeps the reference's expression tree (bit-identical power); exp is the reference's own expf by default
//     (EXACT: alpha, T and every threshold decision bit-identical to the reference's) or one ex2.approx.ftz (gsr_scene.exact_exp < 0),
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
};

// EXACT: alpha from the reference's own exp (expf without fast-math, forward.cu:496; same nvcc, same instruction sequence)
// instead of one ex2.approx: with power and the conic bit-identical, alpha and hence T are then BIT-identical to the
// reference's for every pair, so that every threshold decision -- and with them n_contrib and n_touched -- is the reference's.
template <bool COUNT_TOUCHED, bool EXACT>
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
			const float alpha = fminf(0.99f, w1.y * (EXACT ? expf(power) : gsr_exp(power)));
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

// MODE 1 / 2: the CTA sorts its tile's scattered (depth, id) segment itself (tile_sort.cuh) and composites straight from
// the sorted ids it keeps in shared memory.  One launch less, no point_list round trip before the first gather, and the
// latency-bound sort phases of one CTA overlap the issue-bound blending of the other CTAs on the SM.
//
// ON-DEMAND ORDER (lists above lazy_min entries): front-to-back compositing stops once every pixel of the tile is opaque,
// typically after a few hundred entries, however long the list is (measured: 37 % of the list at 640x480 / 100 k
// Gaussians, 13 % at 1200x680 / 500 k, 2 % at 1920x1080 / 3 M).  So the CTA does not sort the list; it SELECTS: a
// 1024-bin histogram over the leading bits of the tile's depth-key range gives, for any bin boundary, the exact list
// position where that depth slab starts; the entries of the first slab (the fewest bins holding >= kLazyTarget entries)
// are placed bin by bin in shared memory (the bin offsets are known: one MSD step) and every entry is ranked against the
// handful of entries of its own bin on (depth, id) -- barrier-free and exact; rank + bin offset is the entry's position in
// the reference's list.  The ids go to point_list and are composited; only if pixels are still open is the next slab
// selected and ordered.  Positions behind the tile's deepest contributor are never read by the forward or the backward
// (backward.cu:763), and point_list holds the exact list up to there.  A slab that cannot be ordered this way (a bin above
// kRankMax entries: masses of equal depths) makes the CTA sort the whole list the general way and carry on.
struct FusedSortArgs {
	uint2* ranges;
	uint2* pairs;
	uint2* pairs_alt;
	uint32_t* point_list;
	unsigned capacity;
	int id_bits;
	GeomHeader* hdr;
	uint32_t* tile_done;
	int lazy_min;        // lists longer than this are ordered on demand; <= 0: every list is sorted completely
	int tile0;           // first tile of the band this launch renders (CTA b <-> tile tile0 + b); 0 for a whole view
	// depth partition (binning.cu): the scatter put the pairs with depth key <= depth_cut[tile] at the front of the segment
	// (tile_cursor[tile] = where the front part ends) and the others at its back.  Sorted front ++ sorted back IS the sorted
	// list, so the front part is ordered first and the back part only if the tile is still open behind it.
	const uint32_t* tile_cursor;
	uint32_t* depth_cut;     // [tiles] in/out, or null: this kernel writes the hint of the next iteration
	int use_front;
};
constexpr int kFusedIdsOffset = 40960;    // bytes: behind the FwdSmem overlay, inside the sort's counter scratch
constexpr int kFusedIdsCap = GSR_SORT_CHUNK;
constexpr int kLazyBins = 1024;           // depth bins over the tile's key range: four per thread
#ifndef GSR_LAZY_TARGET
#define GSR_LAZY_TARGET 512
#endif
constexpr int kLazyTarget = GSR_LAZY_TARGET;   // entries a slab should hold at least: two compositing batches (experiments: -DGSR_LAZY_TARGET=n)
constexpr int kRankMax = 256;             // longest depth bin ordered by all-pairs ranking
struct LazySmem {                         // lives behind the sort scratch: survives sorts and compositing
	uint32_t excl[kLazyBins + 1];         // list position at which bin b starts
	uint32_t wsum[8];
	uint32_t red[16];
	uint32_t b_hi;
	uint32_t big;                         // the slab holds a bin above kRankMax entries
	// CTA-uniform list state, kept here rather than in registers of the compositing loop
	uint32_t kmin;                        // smallest depth key of the tile
	int shift;                            // bin = (key - kmin) >> shift
	int b_lo;                             // first bin not ordered yet
	int sorted_end;                       // list positions < sorted_end are final (point_list)
	int ids_base, ids_cnt;                // s_ids[i] = id at list position ids_base + i, i < ids_cnt
	// depth partition: the part of the list the fields above refer to (positions relative to part_off)
	int part_off;                         // list position at which the part starts (0: the front part or the whole list)
	int part_n;                           // entries of the part
	int n_front;                          // entries of the front part (== n: nothing behind it)
};

constexpr int kLazyOffset = 51456;        // bytes: behind the sort scratch of one 2048-entry chunk (sort_smem_bytes)
constexpr int kLazyCursorOffset = 32768;  // bytes: per-bin placement cursors of the slab being built (scratch, inside the overlay)

// The helpers below are deliberately NOT inlined: they run once per slab, need the register file for themselves and must
// not raise the register pressure of the compositing loop around them.  They find the shared-memory carve-up themselves (a
// pointer parameter would turn every LDS/STS into a generic access).

// The slab that starts at bin b_lo / list position base: up to the first bin boundary with at least kLazyTarget entries in
// front of it (excl is non-decreasing), one bin less if that overshoots a shared-memory chunk.  Needs lz->b_hi == kLazyBins
// and lz->big == 0 on entry (behind a barrier); one barrier inside, lz->big is valid behind the caller's next one.
__device__ __forceinline__ void lazy_choose_slab(LazySmem* lz, int b_lo, uint32_t base, int& b_hi, int& m)
{
	uint32_t e[5];
#pragma unroll
	for (int u = 0; u < 5; u++) e[u] = lz->excl[4 * threadIdx.x + u];
#pragma unroll
	for (int u = 1; u < 5; u++) {
		const int j = 4 * (int)threadIdx.x + u;
		if (j > b_lo && e[u] - base >= (uint32_t)kLazyTarget && e[u - 1] - base < (uint32_t)kLazyTarget) lz->b_hi = (uint32_t)j;
	}
	__syncthreads();
	b_hi = (int)lz->b_hi;
	m = (int)(lz->excl[b_hi] - base);
	if (m > kSmallChunk && b_hi - 1 > b_lo && lz->excl[b_hi - 1] - base > 0) { b_hi--; m = (int)(lz->excl[b_hi] - base); }
#pragma unroll
	for (int u = 1; u < 5; u++) {
		const int j = 4 * (int)threadIdx.x + u;
		if (j > b_lo && j <= b_hi && e[u] - e[u - 1] > (uint32_t)kRankMax) lz->big = 1;
	}
}

// Orders a slab whose m entries sit in s_keys / s_vals grouped by depth bin (bin b at [excl[b] - base, excl[b + 1] - base),
// unordered inside): every thread ranks its entries against the other entries of their bin on (depth, id) -- exact, no
// barriers, bins hold a handful of entries.  The rank is the final list position: ids -> out (global) and s_ids.
__device__ __forceinline__ void rank_sort_bins(const LazySmem* lz, int m, uint32_t base, uint32_t kmin, int shift,
                                               const uint32_t* s_keys, const uint32_t* s_vals, uint32_t* __restrict__ out,
                                               uint32_t* s_ids)
{
	for (int p = threadIdx.x; p < m; p += 256) {
		const uint32_t k = s_keys[p], v = s_vals[p];
		const uint32_t bin = (k - kmin) >> shift;
		const uint32_t s0 = lz->excl[bin] - base, s1 = lz->excl[bin + 1] - base;
		uint32_t rank = s0;
		for (uint32_t j = s0; j < s1; j++) {
			const uint32_t kj = s_keys[j], vj = s_vals[j];
			rank += (kj < k || (kj == k && vj < v)) ? 1u : 0u;
		}
		out[rank] = v;
		s_ids[rank] = v;
	}
}

// First visit of a list: min / max of its depth keys, 1024-bin histogram over the leading bits of their range, exclusive
// scan (lz->excl[b] = list position at which depth bin b starts); then the first slab is selected, placed bin by bin and
// ordered.  Lists of up to kLazyRegs * 256 entries are read from global memory ONCE and held in registers for the three
// sweeps (range, histogram, placement); longer ones are re-read (L2).  Returns false when the slab cannot be ordered this
// way (a bin above kRankMax entries: masses of equal depths).  All 256 threads.
// SITE: one instantiation per call site -- the one in the kernel's prologue has nothing live around it and keeps the whole
// register file for its 16 pairs per thread; the one behind a used-up front part (rare) sits inside the compositing state,
// whose registers the callee would otherwise have to save around EVERY call (measured: list ready 8.4 -> 14.0 us at C2).
constexpr int kLazyRegs = 16;
template <int SITE>
__device__ __noinline__ bool lazy_first_slab(const uint2* seg, int n, uint32_t* __restrict__ list)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem_raw);
	uint32_t* s_vals = s_keys + 2 * kSmallChunk;
	uint32_t* s_cursor = reinterpret_cast<uint32_t*>(smem_raw + kLazyCursorOffset);
	uint32_t* s_ids = reinterpret_cast<uint32_t*>(smem_raw + kFusedIdsOffset);
	LazySmem* lz = reinterpret_cast<LazySmem*>(smem_raw + kLazyOffset);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool in_regs = SITE == 0 && n <= kLazyRegs * 256;      // (SITE 1 re-reads the pairs from L2: no register array, a light call)
	uint2 kv[kLazyRegs];
	uint32_t kmin = 0xffffffffu, kmax = 0;
	if (in_regs) {
#pragma unroll
		for (int u = 0; u < kLazyRegs; u++) {
			const int i = u * 256 + tid;
			kv[u] = (i < n) ? __ldcg(&seg[i]) : make_uint2(0xffffffffu, 0u);
		}
#pragma unroll
		for (int u = 0; u < kLazyRegs; u++)
			if (u * 256 + tid < n) { kmin = min(kmin, kv[u].x); kmax = max(kmax, kv[u].x); }
	} else {
		int i = tid;
		for (; i + 3 * 256 < n; i += 4 * 256) {
			uint32_t k[4];
#pragma unroll
			for (int u = 0; u < 4; u++) k[u] = __ldcg(&seg[i + u * 256].x);
#pragma unroll
			for (int u = 0; u < 4; u++) { kmin = min(kmin, k[u]); kmax = max(kmax, k[u]); }
		}
		for (; i < n; i += 256) { const uint32_t k = __ldcg(&seg[i].x); kmin = min(kmin, k); kmax = max(kmax, k); }
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		kmin = min(kmin, __shfl_xor_sync(kFull, kmin, o));
		kmax = max(kmax, __shfl_xor_sync(kFull, kmax, o));
	}
	if (lane == 0) { lz->red[warp] = kmin; lz->red[8 + warp] = kmax; }
	reinterpret_cast<uint4*>(lz->excl)[tid] = make_uint4(0, 0, 0, 0);
	reinterpret_cast<uint4*>(s_cursor)[tid] = make_uint4(0, 0, 0, 0);
	if (tid == 0) { lz->b_hi = kLazyBins; lz->big = 0; }
	__syncthreads();
	kmin = lz->red[0]; kmax = lz->red[8];
#pragma unroll
	for (int w = 1; w < 8; w++) { kmin = min(kmin, lz->red[w]); kmax = max(kmax, lz->red[8 + w]); }
	const int sig = (kmax == kmin) ? 0 : 32 - __clz(kmax - kmin);
	const int shift = max(0, sig - 10);
	if (in_regs) {
#pragma unroll
		for (int u = 0; u < kLazyRegs; u++)
			if (u * 256 + tid < n) atomicAdd(&lz->excl[(kv[u].x - kmin) >> shift], 1u);
	} else {
		int i = tid;
		for (; i + 3 * 256 < n; i += 4 * 256) {
			uint32_t k[4];
#pragma unroll
			for (int u = 0; u < 4; u++) k[u] = __ldcg(&seg[i + u * 256].x);
#pragma unroll
			for (int u = 0; u < 4; u++) atomicAdd(&lz->excl[(k[u] - kmin) >> shift], 1u);
		}
		for (; i < n; i += 256) atomicAdd(&lz->excl[(__ldcg(&seg[i].x) - kmin) >> shift], 1u);
	}
	__syncthreads();
	{   // exclusive scan of the bin counts: four consecutive bins per thread
		const uint4 c = reinterpret_cast<uint4*>(lz->excl)[tid];
		const uint32_t sum = c.x + c.y + c.z + c.w;
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(kFull, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) lz->wsum[warp] = incl;
		__syncthreads();
		uint32_t before = incl - sum;
		for (int w = 0; w < warp; w++) before += lz->wsum[w];
		reinterpret_cast<uint4*>(lz->excl)[tid] = make_uint4(before, before + c.x, before + c.x + c.y, before + c.x + c.y + c.z);
		if (tid == 255) lz->excl[kLazyBins] = before + sum;
		if (tid == 0) {
			lz->kmin = kmin; lz->shift = shift; lz->b_lo = 0;
			lz->sorted_end = 0; lz->ids_base = 0; lz->ids_cnt = 0;
		}
	}
	__syncthreads();
	// ---- first slab ----
	int b_hi, m;
	lazy_choose_slab(lz, 0, 0u, b_hi, m);
	__syncthreads();
	if (m > kSmallChunk || lz->big) return false;
	const uint32_t width = (uint32_t)b_hi;
	if (in_regs) {
#pragma unroll
		for (int u = 0; u < kLazyRegs; u++) {
			const uint32_t bin = (kv[u].x - kmin) >> shift;
			if (u * 256 + tid < n && bin < width) {
				const uint32_t pos = lz->excl[bin] + atomicAdd(&s_cursor[bin], 1u);
				s_keys[pos] = kv[u].x; s_vals[pos] = kv[u].y;
			}
		}
	} else {
		for (int i0 = 0; i0 < n; i0 += 4 * 256) {
			uint2 e[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int i = i0 + u * 256 + tid;
				e[u] = (i < n) ? __ldcg(&seg[i]) : make_uint2(0xffffffffu, 0u);
			}
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const uint32_t bin = (e[u].x - kmin) >> shift;
				if (i0 + u * 256 + tid < n && bin < width) {
					const uint32_t pos = lz->excl[bin] + atomicAdd(&s_cursor[bin], 1u);
					s_keys[pos] = e[u].x; s_vals[pos] = e[u].y;
				}
			}
		}
	}
	__syncthreads();
	rank_sort_bins(lz, m, 0u, kmin, shift, s_keys, s_vals, list, s_ids);
	if (tid == 0) { lz->b_lo = b_hi; lz->ids_base = 0; lz->ids_cnt = m; lz->sorted_end = m; }
	return true;
}

// Later visits: selects the next depth slab (bins [b_lo, b_hi), starting at list position sorted_end), places its entries
// bin by bin in the scratch and orders them: ids -> list[sorted_end ...) (global) and s_ids.  Advances the list state in
// LazySmem and returns true; false when the slab cannot be ordered this way (nothing written).  All 256 threads.
__device__ __noinline__ bool lazy_next_slab(const uint2* seg, int n, uint32_t* __restrict__ list)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem_raw);
	uint32_t* s_vals = s_keys + 2 * kSmallChunk;
	uint32_t* s_cursor = reinterpret_cast<uint32_t*>(smem_raw + kLazyCursorOffset);
	uint32_t* s_ids = reinterpret_cast<uint32_t*>(smem_raw + kFusedIdsOffset);
	LazySmem* lz = reinterpret_cast<LazySmem*>(smem_raw + kLazyOffset);
	const int tid = threadIdx.x;
	const int b_lo = lz->b_lo, shift = lz->shift;
	const uint32_t base = (uint32_t)lz->sorted_end, kmin = lz->kmin;
	__syncthreads();      // everyone holds the state before thread 0 touches the fields next to it
	if (tid == 0) { lz->b_hi = kLazyBins; lz->big = 0; lz->ids_cnt = 0; }
	reinterpret_cast<uint4*>(s_cursor)[tid] = make_uint4(0, 0, 0, 0);
	__syncthreads();
	int b_hi, m;
	lazy_choose_slab(lz, b_lo, base, b_hi, m);
	__syncthreads();
	if (m > kSmallChunk || lz->big) return false;
	const uint32_t lo = (uint32_t)b_lo, width = (uint32_t)(b_hi - b_lo);
	for (int i0 = 0; i0 < n; i0 += 4 * 256) {
		uint2 e[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int i = i0 + u * 256 + tid;
			e[u] = (i < n) ? __ldcg(&seg[i]) : make_uint2(0xffffffffu, 0u);
		}
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const uint32_t bin = (e[u].x - kmin) >> shift;
			if (i0 + u * 256 + tid < n && bin - lo < width) {
				const uint32_t pos = lz->excl[bin] - base + atomicAdd(&s_cursor[bin], 1u);
				s_keys[pos] = e[u].x; s_vals[pos] = e[u].y;
			}
		}
	}
	__syncthreads();
	rank_sort_bins(lz, m, base, kmin, shift, s_keys, s_vals, list + base, s_ids);
	if (tid == 0) { lz->b_lo = b_hi; lz->ids_base = (int)base; lz->ids_cnt = m; lz->sorted_end = (int)base + m; }
	return true;
}

// the whole list of the tile, sorted the general way (ids -> point_list and, up to ids_cap of them, s_ids)
__device__ __noinline__ void sort_whole_tile(int tile, FusedSortArgs fs, int ids_cap)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	sort_tile<256>(tile, fs.ranges, fs.pairs, fs.pairs_alt, fs.point_list, fs.capacity, kSmallChunk, fs.id_bits, fs.hdr,
	               reinterpret_cast<uint32_t*>(smem_raw), ids_cap > 0 ? reinterpret_cast<uint32_t*>(smem_raw + kFusedIdsOffset) : nullptr,
	               ids_cap);
}

// MODE 0: point_list holds the sorted lists.  MODE 1: the CTA sorts its whole list first (every list expected to fit one
// shared-memory chunk; anything longer takes the general path).  MODE 2: lists above fs.lazy_min are ordered on demand.
// LOSS: the view's SLAM loss (slam_ops.cu slam_loss_kernel, same arithmetic) is evaluated in the epilogue from the pixel values
// still in registers: dL/dcolor, dL/ddepth and the per-tile partial sums of {loss, dL/da, dL/db}; the last CTA adds the
// partials in tile order.
#ifndef GSR_FWD_MINB
#define GSR_FWD_MINB 4      // CTAs per SM the forward compositing kernel is compiled for (experiments: GSR_EXTRA_NVCC_FLAGS=-DGSR_FWD_MINB=3)
#endif
template <int MODE, bool LOSS, bool EXACT>
__global__ void __launch_bounds__(256, GSR_FWD_MINB)
render_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* point_list,
                      const GaussRec* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_opacity, int* __restrict__ n_touched,
                      uint32_t* __restrict__ cull_masks, FusedSortArgs fs, FusedLoss lf)
{
	pdl_launch_dependents();      // a backward launched as programmatic dependent may move in as soon as every tile has a CTA
	pdl_wait();                   // launched as programmatic dependent of the cooperative preprocess: its lists are complete from here
	extern __shared__ __align__(16) unsigned char smem_raw[];
	static_assert(sizeof(FwdSmem) <= kFusedIdsOffset, "sorted ids must sit behind the compositing overlay");
	FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
	const uint32_t* s_ids = reinterpret_cast<const uint32_t*>(smem_raw + kFusedIdsOffset);
	LazySmem* lz = reinterpret_cast<LazySmem*>(smem_raw + kLazyOffset);
	const int tile = fs.tile0 + (int)blockIdx.x;

	GSR_PROBE(3, 0);
	// ---- the tile's list: complete, or ordered on demand (positions < lz->sorted_end are final) ----
	uint2 range;
	int n;
	if (MODE == 2) {
		range = fs.ranges[tile];
		n = (int)(range.y - range.x);
		if (range.y <= fs.capacity && n > fs.lazy_min) {
			int nf = n;      // entries of the front part (depth partition), else the whole list
			if (fs.use_front) {
				const int f = (int)(__ldcg(&fs.tile_cursor[tile]) - range.x);
				if (f > 0 && f < n) nf = f;
			}
			if (threadIdx.x == 0) { lz->part_off = 0; lz->part_n = nf; lz->n_front = nf; }
			if (!lazy_first_slab<0>(fs.pairs + range.x, nf, fs.point_list + range.x)) {
				// the first depth bin alone is beyond the chunk length: sort the whole list the general way
				sort_whole_tile(tile, fs, 0);
				if (threadIdx.x == 0) { lz->sorted_end = n; lz->ids_cnt = 0; lz->part_off = 0; lz->part_n = n; lz->n_front = n; }
			}
			__threadfence_block();
			__syncthreads();
		} else {
			sort_whole_tile(tile, fs, kFusedIdsCap);
			__threadfence_block();
			__syncthreads();      // sorted ids (shared + global) and a possibly clamped range are visible to the whole CTA
			range = __ldcg(&fs.ranges[tile]);
			n = (int)(range.y - range.x);
			if (threadIdx.x == 0) {
				lz->sorted_end = n; lz->ids_base = 0; lz->ids_cnt = min(n, kFusedIdsCap);
				lz->part_off = 0; lz->part_n = n; lz->n_front = n;
			}
			__syncthreads();
		}
	} else if (MODE == 1) {
		sort_tile<256>(tile, fs.ranges, fs.pairs, fs.pairs_alt, fs.point_list, fs.capacity, kSmallChunk, fs.id_bits, fs.hdr,
		               reinterpret_cast<uint32_t*>(smem_raw), reinterpret_cast<uint32_t*>(smem_raw + kFusedIdsOffset), kFusedIdsCap);
		__threadfence_block();
		__syncthreads();
		range = __ldcg(&fs.ranges[tile]);
		n = (int)(range.y - range.x);
	} else {
		range = ranges[tile];
		n = (int)(range.y - range.x);
	}
	const bool ids_in_smem = MODE == 1 && n <= kFusedIdsCap;
	GSR_PROBE(3, 1);

	const int tile_y = tile / grid_x, tile_x = tile - tile_y * grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	int px, py;
	pixel_of_thread(tile_x, tile_y, px, py);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float bx0 = (float)(tile_x * GSR_TILE + (warp & 1) * 8), by0 = (float)(tile_y * GSR_TILE + (warp >> 1) * 4);
	const float bx1 = bx0 + 7.f, by1 = by0 + 3.f;
	const int rounds = (n + 255) / 256;
	QueueRec* wq = sm.queue[warp];
	uint32_t* wmask = sm.tmask[warp];
	uint32_t* my_masks = cull_masks + cull_mask_base(range.x, (uint32_t)tile) + warp;   // [group of 32 positions][warp]

	float T = inside ? 1.0f : -1.0f;   // sign bit = "done"
	float C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
	int last = -1;                     // list position of the last blended entry
	// T only decreases, so once no live pixel of the warp is above 0.5 the n_touched bookkeeping is skipped for good.
	bool warp_hi_T = __any_sync(kFull, inside);

	auto stage = [&](int b, int buf) {
		const int i = b * 256 + threadIdx.x;
		if (i < n) {
			// fused: ids of the chunk sorted last come from shared memory; older ones were written by this CTA -> coherent load
			uint32_t id;
			if (MODE == 2) {
				const int j = i - lz->part_off - lz->ids_base;
				id = ((unsigned)j < (unsigned)lz->ids_cnt) ? s_ids[j] : __ldcg(point_list + range.x + i);
			} else if (MODE == 1) id = ids_in_smem ? s_ids[i] : __ldcg(point_list + range.x + i);
			else id = __ldg(point_list + range.x + i);
			sm.id[buf][threadIdx.x] = id;
			const GaussRec* r = rec + id;
			cp_async16(&sm.rec[buf][threadIdx.x].q0, &r->q0);
			cp_async16(&sm.rec[buf][threadIdx.x].q1, &r->q1);
			cp_async16(&sm.rec[buf][threadIdx.x].q2, &r->q2);
		}
		cp_async_commit();
	};

	int b = 0;
	bool all_done = false;
	for (;;) {
		// batches of 256 positions that are completely ordered by now
		const int sorted_end = (MODE == 2) ? lz->part_off + lz->sorted_end : n;      // absolute list position
		const int b_end = (sorted_end >= n) ? rounds : (sorted_end >> 8);
		if (b < b_end) {
			stage(b, b & 1);
			for (; b < b_end; b++) {
				const int buf = b & 1;
				if (__syncthreads_and(T < 0.f)) { all_done = true; break; }   // also: everyone is past batch b-1, buffer buf^1 is free
				if (b + 1 < b_end) stage(b + 1, buf ^ 1);
				else cp_async_commit();
				cp_async_wait<1>();
				__syncthreads();
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
						blend_queue<true, EXACT>(wq, nq, pxf, pyf, T, C0, C1, C2, D, last, wmask, lane);
						__syncwarp();
						if (keep) {
							const unsigned m = wmask[pos];
							if (m) atomicAdd(&n_touched[__float_as_uint(wq[pos].w2.w)], __popc(m));
						}
						warp_hi_T = __any_sync(kFull, T > 0.5f);
					} else {
						blend_queue<false, EXACT>(wq, nq, pxf, pyf, T, C0, C1, C2, D, last, wmask, lane);
					}
					__syncwarp();   // queue fully consumed before the next chunk overwrites it
				}
				// the backward replays these ballots instead of repeating the cull (chunks this warp skipped stay 0: they lie
				// behind its last contributor and are never read)
				if (lane < 8 && (b * 8 + lane) * 32 < n) my_masks[(b * 8 + lane) * 8] = held_mask;   // only this tile's own groups
			}
			cp_async_wait<0>();
		}
		GSR_PROBE(3, 2);
		if (MODE != 2 || all_done || sorted_end >= n) break;
		// ---- more of the list is needed: order the next depth slab (the staging buffers are idle: scratch again) ----
		if (__syncthreads_and(T < 0.f)) break;
		{
			const int off = lz->part_off, pn = lz->part_n;
			bool ok;
			if (off == 0 && pn < n && lz->sorted_end >= pn) {
				// the front part is used up and the tile is still open: go on with the back part (a list of its own whose
				// positions follow the front part's)
				__syncthreads();      // everyone has read the old part before thread 0 replaces it
				if (threadIdx.x == 0) { lz->part_off = pn; lz->part_n = n - pn; }
				ok = lazy_first_slab<1>(fs.pairs + range.x + pn, n - pn, fs.point_list + range.x + pn);
			} else {
				ok = lazy_next_slab(fs.pairs + range.x + off, pn, fs.point_list + range.x + off);
			}
			if (!ok) {
				// a depth bin beyond the chunk length: sort the whole list the general way (same prefix) and carry on
				sort_whole_tile(tile, fs, 0);
				if (threadIdx.x == 0) { lz->sorted_end = n; lz->ids_cnt = 0; lz->part_off = 0; lz->part_n = n; }
			}
		}
		__threadfence_block();
		__syncthreads();
		GSR_PROBE(3, 3);
	}
	GSR_PROBE(3, 4);
	const bool open = inside && !(T < 0.f);      // this pixel never reached the stop condition: it read its whole list
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
	if (LOSS) {
		const int HW = H * W;
		float loss = 0.f, ga = 0.f, gb = 0.f;
		if (inside) {
			const int p = W * py + px;
			const float ea = lf.exposure ? __expf(lf.exposure[0]) : 1.f;       // image_ab = exp(a) * image + b
			const float eb = lf.exposure ? lf.exposure[1] : 0.f;
			const float inv_rgb = 1.f / (3.f * (float)HW), inv_d = 1.f / (float)HW;
			const float w_rgb = lf.use_depth ? lf.alpha : 1.f, w_d = 1.f - lf.alpha;
			const float gt[3] = {lf.gt_color[p], lf.gt_color[HW + p], lf.gt_color[2 * HW + p]};
			float m = (gt[0] + gt[1] + gt[2] > lf.rgb_boundary_threshold) ? 1.f : 0.f;
			if (lf.grad_mask) m *= lf.grad_mask[p] ? 1.f : 0.f;
			const float op = 1 - T;
			const float wgt = lf.opacity_weighted ? op : 1.f;
			const float c[3] = {C0 + T * bg[0], C1 + T * bg[1], C2 + T * bg[2]};
#pragma unroll
			for (int ch = 0; ch < 3; ch++) {
				const float iab = ea * c[ch] + eb;
				const float d = iab * m - gt[ch] * m;
				loss += w_rgb * inv_rgb * wgt * fabsf(d);
				const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
				const float g_iab = w_rgb * inv_rgb * wgt * m * sgn;
				lf.dL_dcolor[ch * HW + p] = g_iab * ea;
				ga += g_iab * ea * c[ch];
				gb += g_iab;
			}
			float gd = 0.f;
			if (lf.use_depth) {
				const float gtd = lf.gt_depth[p];
				float dm = (gtd > 0.01f) ? 1.f : 0.f;
				if (lf.opacity_weighted) dm *= (op > 0.95f) ? 1.f : 0.f;
				const float dd = D * dm - gtd * dm;
				loss += w_d * inv_d * fabsf(dd);
				gd = w_d * inv_d * dm * ((dd > 0.f) ? 1.f : ((dd < 0.f) ? -1.f : 0.f));
			}
			lf.dL_ddepth[p] = gd;
		}
		loss = warp_sum(loss); ga = warp_sum(ga); gb = warp_sum(gb);
		float* s_red = reinterpret_cast<float*>(sm.tmask);      // idle by now
		__syncthreads();
		if (lane == 0) { s_red[warp * 4] = loss; s_red[warp * 4 + 1] = ga; s_red[warp * 4 + 2] = gb; }
		__syncthreads();
		if (threadIdx.x < 3) {
			float v = 0.f;
#pragma unroll
			for (int w = 0; w < 8; w++) v += s_red[w * 4 + threadIdx.x];
			lf.partials[blockIdx.x * 4 + threadIdx.x] = v;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			s_red[32] = (atomicAdd(lf.ticket, 1u) == gridDim.x - 1) ? 1.f : 0.f;
		}
		__syncthreads();
		if (s_red[32] != 0.f) {      // last CTA: deterministic sum over the tiles, warp k takes component k
			__threadfence();
			if (threadIdx.x < 96) {
				float v = 0.f;
				for (unsigned t = lane; t < gridDim.x; t += 32) v += __ldcg(&lf.partials[t * 4 + warp]);
				v = warp_sum(v);
				if (lane == 0) lf.sums[warp] = v;
			}
			if (threadIdx.x == 0) *lf.ticket = 0;
		}
	}
	// publish the tile: everything the backward reads of it (final_T, n_contrib, point_list, cull_masks) is written.  The flag
	// carries the tile's deepest contributor (+ 1, so that it is never 0): the backward CTA knows how far to walk the list
	// the moment it has acquired the flag and starts its first gather without waiting for its own n_contrib loads.
	{
		const unsigned wtop = __reduce_max_sync(kFull, inside ? (unsigned)(last + 1) : 0u);
		const unsigned wopen = __any_sync(kFull, open) ? 1u : 0u;
		if (lane == 0) { sm.tmask[warp][31] = wtop; sm.tmask[warp][30] = wopen; }      // words 30, 31 of a warp's row: the loss scratch is done with
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned top = 0, any_open = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) { top = max(top, sm.tmask[w][31]); any_open |= sm.tmask[w][30]; }
		st_release_u32(fs.tile_done + tile, top + 1u);      // cumulative: orders the CTA's writes behind the barrier
		if (MODE == 2 && fs.depth_cut) {
			// the depth-partition hint of the NEXT iteration of this view (binning.cu): the depth key a little behind the batch
			// that holds the deepest contributor, so that the complete batches of the front part cover everything this tile
			// read.  A tile with an open pixel reads its whole list: no cut.  Only a hint -- any value gives the same results.
			uint32_t cut = 0x7f800000u;      // +inf: everything in front
			const int sorted_abs = lz->part_off + lz->sorted_end;
			if (!any_open && top > 0u && n > 0) {
				const int want = (int)((top + 255u) & ~255u) + 16;
				const int pos = min(min(want, sorted_abs - 1), n - 1);
				if (pos >= 0) cut = __float_as_uint(__ldcg(&rec[__ldcg(point_list + range.x + pos)].q1.z));
			}
			fs.depth_cut[tile] = cut;
		}
	}
}

}  // namespace

void launch_render_forward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, float* out_color,
                           float* out_depth, float* out_opacity, int* n_touched, bool fused_sort, int lazy_min, size_t R_capacity,
                           cudaStream_t stream, bool behind_preprocess, bool front_partition)
{
	const bool band = s.band_y1 > 0;
	const int tiles = band ? s.grid_x * (s.band_y1 - s.band_y0) : s.grid_x * s.grid_y;
	if (tiles <= 0) return;
	// behind_preprocess: the cooperative preprocess + scatter kernel was launched just before on this stream and releases its
	// dependents behind its grid barrier: the first forward CTAs are resident (and waiting) when it ends
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(tiles); cfg.blockDim = dim3(256); cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = behind_preprocess ? 1 : 0;
	FusedSortArgs fs;
	fs.ranges = g.ranges; fs.pairs = b.pairs; fs.pairs_alt = b.pairs_alt; fs.point_list = b.point_list;
	fs.capacity = (unsigned)R_capacity; fs.hdr = g.hdr; fs.tile_done = g.tile_done;
	fs.id_bits = 1;
	while (fs.id_bits < 32 && (1ll << fs.id_bits) < (long long)s.P) fs.id_bits++;
	fs.lazy_min = lazy_min;
	fs.tile0 = band ? s.band_y0 * s.grid_x : 0;
	fs.tile_cursor = g.tile_cursor;
	fs.use_front = front_partition ? 1 : 0;
	fs.depth_cut = (s.depth_cut && depth_partition_active(s)) ? s.depth_cut : nullptr;
	const size_t smem_plain = sizeof(FwdSmem);
	static_assert(kLazyOffset == (4 * kSmallChunk + 8 * kMaxBins + kMaxBins + 64) * 4, "LazySmem sits right behind the sort scratch");
	static_assert(kFusedIdsOffset + kFusedIdsCap * 4 <= kLazyOffset, "sorted ids end inside the sort scratch");
	const size_t smem_fused = (kLazyOffset + sizeof(LazySmem) + 15) / 16 * 16;
	const size_t smem_sort = sort_smem_bytes(kSmallChunk, 256);
	static SmemAttrCache attr[12];
	const int mode = (fused_sort && s.P > 0 && R_capacity > 0) ? (lazy_min > 0 ? 2 : 1) : 0;
	const size_t smem = mode == 2 ? smem_fused : (mode == 1 ? smem_sort : smem_plain);
	FusedLoss lf = s.loss;
#define GSR_FWD_LAUNCH_E(M, L, E)                                                                                                 \
	do {                                                                                                                          \
		ensure_dynamic_smem(render_forward_kernel<M, L, E>, smem, attr[4 * M + (L ? 2 : 0) + (E ? 1 : 0)]);                       \
		cfg.dynamicSmemBytes = smem;                                                                                              \
		cudaLaunchKernelEx(&cfg, render_forward_kernel<M, L, E>, (const uint2*)g.ranges, (const uint32_t*)b.point_list,           \
		                   (const GaussRec*)g.rec, s.W, s.H, s.grid_x, (const float*)s.background, im.final_T, im.n_contrib,      \
		                   out_color, out_depth, out_opacity, n_touched, b.cull_masks, fs, lf);                                   \
	} while (0)
#define GSR_FWD_LAUNCH(M, L)                                                                                                      \
	do {                                                                                                                          \
		if (s.exact_exp) GSR_FWD_LAUNCH_E(M, L, true);                                                                            \
		else GSR_FWD_LAUNCH_E(M, L, false);                                                                                       \
	} while (0)
	if (s.has_loss) {
		if (mode == 2) GSR_FWD_LAUNCH(2, true);
		else if (mode == 1) GSR_FWD_LAUNCH(1, true);
		else GSR_FWD_LAUNCH(0, true);
	} else {
		if (mode == 2) GSR_FWD_LAUNCH(2, false);
		else if (mode == 1) GSR_FWD_LAUNCH(1, false);
		else GSR_FWD_LAUNCH(0, false);
	}
#undef GSR_FWD_LAUNCH
#undef GSR_FWD_LAUNCH_E
}

GSR_PROBE_READER(probe_read_render)

}  // namespace gsr
