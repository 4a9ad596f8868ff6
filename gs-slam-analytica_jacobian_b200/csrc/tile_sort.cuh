// Per-tile sort of (depth bits, Gaussian id) pairs -- device code shared by the stand-alone sort kernels (binning.cu)
// and the fused sort + compositing forward kernel (render.cu).  See binning.cu for the algorithm.
#pragma once
#include "gsr_params.h"

namespace gsr {
namespace {

// ---- 2. per-tile sort -------------------------------------------------------------------------------
constexpr int kSortItems = 8;                       // keys per thread and chunk
constexpr int kMaxDigitBits = 9;
constexpr int kMaxBins = 1 << kMaxDigitBits;
constexpr int kSmallThreads = 256, kLongThreads = 512;
constexpr int kSmallChunk = kSmallThreads * kSortItems;   // 2048
static_assert(kSmallChunk == GSR_SORT_CHUNK, "the preprocess scan, the API and the sort kernels agree on the chunk length");
constexpr int kLongChunk = kLongThreads * kSortItems;     // 4096

struct Field {       // which 32-bit word of the pair a pass looks at
	int word;        // 0 = depth key (minus the tile minimum), 1 = Gaussian id
	int shift, bits;
};

__device__ __forceinline__ uint32_t digit_of(uint32_t k, uint32_t v, uint32_t kmin, Field f, uint32_t mask)
{
	return (((f.word == 0) ? (k - kmin) : v) >> f.shift) & mask;
}

// exclusive scan of s_base[0..nb) by warp 0 (nb <= 512: each lane scans nb/32 consecutive bins)
__device__ __forceinline__ void scan_bins(uint32_t* s_base, int nb, int lane)
{
	const int per = (nb + 31) / 32;
	uint32_t sum = 0;
	for (int j = 0; j < per; j++) { const int b = lane * per + j; if (b < nb) sum += s_base[b]; }
	uint32_t inc = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	uint32_t run = inc - sum;
	for (int j = 0; j < per; j++) {
		const int b = lane * per + j;
		if (b < nb) { const uint32_t c = s_base[b]; s_base[b] = run; run += c; }
	}
}

// General stable counting pass over n pairs held in (kin, vin) -> (kout, vout); all arrays may live in shared or
// global memory.  A counting sweep (shared-memory atomics) gives the digit bases; chunks of NT*8 pairs are then
// ranked with warp match.any + per-warp digit counters (stable) and scattered, bases advancing chunk by chunk.
template <int NT>
__device__ __forceinline__ void radix_pass(const uint32_t* kin, const uint32_t* vin, int in_stride, uint32_t* kout,
                                           uint32_t* vout, int out_stride, int n, uint32_t kmin, Field f,
                                           uint32_t* s_cnt /*[NT/32][kMaxBins]*/, uint32_t* s_base /*[kMaxBins]*/)
{
	constexpr int NW = NT / 32;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int nb = 1 << f.bits;
	const uint32_t mask = (uint32_t)nb - 1;
	for (int i = tid; i < nb; i += NT) s_base[i] = 0;
	__syncthreads();
	for (int i = tid; i < n; i += NT) atomicAdd(&s_base[digit_of(kin[(size_t)i * in_stride], vin[(size_t)i * in_stride], kmin, f, mask)], 1u);
	__syncthreads();
	if (warp == 0) scan_bins(s_base, nb, lane);
	__syncthreads();
	const uint32_t lt = (1u << lane) - 1;
	for (int c0 = 0; c0 < n; c0 += NT * kSortItems) {
		for (int i = tid; i < NW * kMaxBins / 4; i += NT) reinterpret_cast<uint4*>(s_cnt)[i] = make_uint4(0, 0, 0, 0);
		__syncthreads();
		uint32_t k[kSortItems], v[kSortItems], rank[kSortItems];
		const int wbase = c0 + warp * (32 * kSortItems) + lane;
#pragma unroll
		for (int i = 0; i < kSortItems; i++) {
			const int pos = wbase + i * 32;
			if (pos < n) { k[i] = kin[(size_t)pos * in_stride]; v[i] = vin[(size_t)pos * in_stride]; }
			else { k[i] = 0xffffffffu; v[i] = 0xffffffffu; }
		}
#pragma unroll
		for (int i = 0; i < kSortItems; i++) {
			const int pos = wbase + i * 32;
			const uint32_t d = (pos < n) ? digit_of(k[i], v[i], kmin, f, mask) : mask;   // padding ranks after every real key of its warp
			const unsigned peers = __match_any_sync(0xffffffffu, d);
			const uint32_t pre = s_cnt[warp * kMaxBins + d];
			__syncwarp();
			rank[i] = pre + __popc(peers & lt);
			if (lane == 31 - __clz(peers)) s_cnt[warp * kMaxBins + d] = pre + __popc(peers);
			__syncwarp();
		}
		__syncthreads();
		// per digit: exclusive scan over the warps; chunk total advances the base AFTER the scatter
		uint32_t tot[(kMaxBins + NT - 1) / NT];
		for (int j = 0, d = tid; d < nb; d += NT, j++) {
			uint32_t total = 0;
#pragma unroll 8
			for (int w = 0; w < NW; w++) {
				const uint32_t c = s_cnt[w * kMaxBins + d];
				s_cnt[w * kMaxBins + d] = total;
				total += c;
			}
			tot[j] = total;
		}
		__syncthreads();
#pragma unroll
		for (int i = 0; i < kSortItems; i++) {
			const int pos = wbase + i * 32;
			if (pos < n) {
				const uint32_t d = digit_of(k[i], v[i], kmin, f, mask);
				const uint32_t dst = s_base[d] + s_cnt[warp * kMaxBins + d] + rank[i];
				kout[(size_t)dst * out_stride] = k[i];
				vout[(size_t)dst * out_stride] = v[i];
			}
		}
		__syncthreads();
		for (int j = 0, d = tid; d < nb; d += NT, j++) s_base[d] += tot[j];
		// (the padding of the last chunk only inflates bin `mask` after its real keys: harmless)
		__syncthreads();
	}
}

// Single-chunk variant for segments of at most NT * ITEMS pairs held in shared memory: the digit histogram falls out of
// the ranking (per-warp counters), so there is no separate counting sweep and no atomics.  Digits are at most
// kFastDigitBits = 8 bits wide, so that (a) one thread owns one digit in the scan phases (NT >= 256) and (b) two passes
// use disjoint column halves of the counter rows (half = 0 / 1), which the caller zeroes ONCE before the first pass.
// Four barrier-separated phases: rank | per-digit totals + warp scan | cross-warp prefix | scatter.
constexpr int kFastDigitBits = 8;
template <int NT, int ITEMS>
__device__ __forceinline__ void radix_pass_small(const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout, int n,
                                                 uint32_t kmin, Field f, int half, uint32_t* s_cnt /*[NT/32][kMaxBins]*/,
                                                 uint32_t* s_base /*[kMaxBins]*/, uint32_t* s_wsum /*[32]*/)
{
	constexpr int NW = NT / 32;
	static_assert(NT >= (1 << kFastDigitBits), "one thread per digit");
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int nb = 1 << f.bits;
	const uint32_t mask = (uint32_t)nb - 1;
	const uint32_t lt = (1u << lane) - 1;
	uint32_t* cnt = s_cnt + half * (kMaxBins / 2);      // this pass's column half of every counter row
	uint32_t k[ITEMS], v[ITEMS], rank[ITEMS], dg[ITEMS];
	const int wbase = warp * (32 * ITEMS) + lane;
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const int pos = wbase + i * 32;
		if (pos < n) { k[i] = kin[pos]; v[i] = vin[pos]; }
		else { k[i] = 0xffffffffu; v[i] = 0xffffffffu; }
	}
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const int pos = wbase + i * 32;
		const uint32_t d = (pos < n) ? digit_of(k[i], v[i], kmin, f, mask) : mask;   // padding ranks last in its warp
		dg[i] = d;
		const unsigned peers = __match_any_sync(0xffffffffu, d);
		const uint32_t pre = cnt[warp * kMaxBins + d];
		__syncwarp();
		rank[i] = pre + __popc(peers & lt);
		if (lane == 31 - __clz(peers)) cnt[warp * kMaxBins + d] = pre + __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	// thread d: exclusive scan of digit d over the warps, then an inclusive warp scan of the digit totals
	uint32_t total = 0, incl = 0;
	if (tid < nb) {
#pragma unroll 8
		for (int w = 0; w < NW; w++) {
			const uint32_t c = cnt[w * kMaxBins + tid];
			cnt[w * kMaxBins + tid] = total;
			total += c;
		}
	}
	incl = total;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) s_wsum[warp] = incl;      // (the padding only inflates bin `mask`, behind every real key)
	__syncthreads();
	if (tid < nb) {
		uint32_t before = 0;
		for (int w = 0; w < warp; w++) before += s_wsum[w];
		s_base[tid] = before + incl - total;
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const int pos = wbase + i * 32;
		if (pos < n) {
			const uint32_t dst = s_base[dg[i]] + cnt[warp * kMaxBins + dg[i]] + rank[i];
			kout[dst] = k[i];
			vout[dst] = v[i];
		}
	}
	__syncthreads();
}

// Odd-even transposition sweeps on (key, id) until the segment is in (depth, id) order; cheap finisher for a
// segment that is already sorted on its leading key bits.  K / V may be shared (stride 1) or the two words of global
// uint2 pairs (stride 2).  Returns false if it did not converge in max_sweeps.
template <int NT>
__device__ __forceinline__ bool finish_by_transposition(uint32_t* K, uint32_t* V, int stride, int n, int max_sweeps)
{
	for (int sweep = 0; sweep < max_sweeps; sweep++) {
		int swapped = 0;
#pragma unroll
		for (int par = 0; par < 2; par++) {
			for (int i = 2 * (int)threadIdx.x + par; i + 1 < n; i += 2 * NT) {
				const size_t a = (size_t)i * stride, b = a + stride;
				const uint32_t k0 = K[a], k1 = K[b];
				if (k0 >= k1) {
					const uint32_t v0 = V[a], v1 = V[b];
					if (k0 > k1 || v0 > v1) {
						K[a] = k1; K[b] = k0; V[a] = v1; V[b] = v0;
						swapped = 1;
					}
				}
			}
			if (par == 0) __syncthreads();
		}
		if (!__syncthreads_or(swapped)) return true;      // also the barrier behind the odd half-round
	}
	return false;
}

__device__ __forceinline__ int plan_passes(int sigbits, int word, Field* out, int max_bits = kMaxDigitBits)
{
	if (sigbits <= 0) return 0;
	const int np = (sigbits + max_bits - 1) / max_bits;
	const int b = (sigbits + np - 1) / np;
	for (int p = 0; p < np; p++) {
		out[p].word = word;
		out[p].shift = p * b;
		out[p].bits = min(b, sigbits - p * b);
	}
	return np;
}

// Sorts n >= 2 pairs that already sit in shared memory (buffer 0 of the layout below; n <= cap_smem; the per-warp counters
// cleared if n <= NT * kSortItems; a CTA barrier behind both) into (depth, id) order and hands the ids to emit(i, id).
// Every key k satisfies 0 <= k - kmin < 2^sig.
// smem: keys[2][cap] | vals[2][cap] | cnt[NT/32][512] | base[512] | red[64]
template <int NT, typename Emit>
__device__ __forceinline__ void sort_loaded(int n, int cap_smem, uint32_t kmin, int sig, int id_bits, uint32_t* sm, Emit&& emit)
{
	constexpr int NW = NT / 32;
	uint32_t* s_keys = sm;
	uint32_t* s_vals = sm + 2 * (size_t)cap_smem;
	uint32_t* s_cnt = sm + 4 * (size_t)cap_smem;
	uint32_t* s_base = s_cnt + NW * kMaxBins;
	uint32_t* s_red = s_base + kMaxBins;
	const int tid = threadIdx.x;
	const bool one_chunk = n <= NT * kSortItems;
	int cur = 0;
	// Fast path: at most two stable passes over the LEADING key bits (16 in one shared-memory chunk, 18 otherwise), then
	// transposition sweeps settle the low bits and the id order of equal depths (adjacent by then).
	{
		const int digit_bits = one_chunk ? kFastDigitBits : kMaxDigitBits;
		const int top = min(sig, 2 * digit_bits);
		Field fp[2];
		const int np = plan_passes(top, 0, fp, digit_bits);
		for (int p = 0; p < np; p++) fp[p].shift += sig - top;
		for (int p = 0; p < np; p++) {
			uint32_t *ki = s_keys + cur * cap_smem, *vi = s_vals + cur * cap_smem;
			uint32_t *ko = s_keys + (cur ^ 1) * cap_smem, *vo = s_vals + (cur ^ 1) * cap_smem;
			if (!one_chunk) radix_pass<NT>(ki, vi, 1, ko, vo, 1, n, kmin, fp[p], s_cnt, s_base);
			else if (n <= NT * 2) radix_pass_small<NT, 2>(ki, vi, ko, vo, n, kmin, fp[p], p, s_cnt, s_base, s_red);
			else if (n <= NT * 3) radix_pass_small<NT, 3>(ki, vi, ko, vo, n, kmin, fp[p], p, s_cnt, s_base, s_red);
			else if (n <= NT * 4) radix_pass_small<NT, 4>(ki, vi, ko, vo, n, kmin, fp[p], p, s_cnt, s_base, s_red);
			else radix_pass_small<NT, kSortItems>(ki, vi, ko, vo, n, kmin, fp[p], p, s_cnt, s_base, s_red);
			cur ^= 1;
		}
		uint32_t *K = s_keys + cur * cap_smem, *V = s_vals + cur * cap_smem;
		if (finish_by_transposition<NT>(K, V, 1, n, 24)) {
			for (int i = tid; i < n; i += NT) emit(i, V[i]);
			return;
		}
		__syncthreads();
	}
	// The sweeps did not converge (masses of equal depths).  General sort from the permutation at hand: stable LSD passes,
	// id digits first, then every differing depth bit.
	Field depth_passes[4], id_passes[4];
	const int nd = plan_passes(sig, 0, depth_passes);
	const int ni = plan_passes(id_bits, 1, id_passes);
	for (int p = 0; p < ni + nd; p++) {
		const Field f = p < ni ? id_passes[p] : depth_passes[p - ni];
		radix_pass<NT>(s_keys + cur * cap_smem, s_vals + cur * cap_smem, 1, s_keys + (cur ^ 1) * cap_smem,
		               s_vals + (cur ^ 1) * cap_smem, 1, n, kmin, f, s_cnt, s_base);
		cur ^= 1;
	}
	const uint32_t* V = s_vals + cur * cap_smem;
	for (int i = tid; i < n; i += NT) emit(i, V[i]);
}

// Sorts one segment of n >= 2 (depth bits, id) pairs: seg[0..n) (global) -> out[0..n) ids in (depth, id) order; alt is a
// same-sized global scratch (ping-pong partner), both may be overwritten.  Segments up to cap_smem entries are sorted in
// shared memory (sort_loaded), longer ones by chunked passes through seg / alt.
template <int NT>
__device__ __forceinline__ void sort_segment(uint2* __restrict__ seg, uint2* __restrict__ alt, uint32_t* __restrict__ out, int n,
                                             int cap_smem, int id_bits, uint32_t* sm, uint32_t* ids_out, int ids_cap)
{
	// sorted ids go to `out` (global) and, when the caller wants to keep consuming them (fused sort + render), also to
	// ids_out[0 .. ids_cap) in shared memory (a region this function's scratch does not use while emitting)
	auto emit = [&](int i, uint32_t v) {
		out[i] = v;
		if (ids_out && i < ids_cap) ids_out[i] = v;
	};
	constexpr int NW = NT / 32;
	uint32_t* s_keys = sm;
	uint32_t* s_vals = sm + 2 * (size_t)cap_smem;
	uint32_t* s_cnt = sm + 4 * (size_t)cap_smem;
	uint32_t* s_base = s_cnt + NW * kMaxBins;
	uint32_t* s_red = s_base + kMaxBins;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool in_smem = n <= cap_smem;

	const bool one_chunk = in_smem && n <= NT * kSortItems;
	// the counters of both single-chunk passes (disjoint column halves) are cleared once, here
	if (one_chunk)
		for (int i = tid; i < NW * kMaxBins / 4; i += NT) reinterpret_cast<uint4*>(s_cnt)[i] = make_uint4(0, 0, 0, 0);
	// min / max depth key of the tile -> the digits that matter
	uint32_t kmin = 0xffffffffu, kmax = 0;
	for (int i = tid; i < n; i += NT) {
		const uint2 kv = seg[i];
		if (in_smem) { s_keys[i] = kv.x; s_vals[i] = kv.y; }
		kmin = min(kmin, kv.x);
		kmax = max(kmax, kv.x);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
		kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
	}
	if (lane == 0) { s_red[warp] = kmin; s_red[32 + warp] = kmax; }
	__syncthreads();
	kmin = s_red[0]; kmax = s_red[32];
	for (int w = 1; w < NW; w++) { kmin = min(kmin, s_red[w]); kmax = max(kmax, s_red[32 + w]); }
	const int sig = (kmax == kmin) ? 0 : 32 - __clz(kmax - kmin);
	if (in_smem) {
		sort_loaded<NT>(n, cap_smem, kmin, sig, id_bits, sm, emit);
		return;
	}

	// List beyond the shared-memory capacity: two stable passes over the leading 18 key bits through the global ping-pong
	// buffers, transposition sweeps on the global pairs; if they do not converge, the general LSD sort (id digits, then
	// every differing depth bit) from the permutation at hand.
	Field depth_passes[4], id_passes[4];
	const int nd = plan_passes(sig, 0, depth_passes);
	const int ni = plan_passes(id_bits, 1, id_passes);
	uint2* A = seg;
	uint2* B = alt;
	{
		const int top = min(sig, 2 * kMaxDigitBits);
		Field fp[2];
		const int np = plan_passes(top, 0, fp, kMaxDigitBits);
		for (int p = 0; p < np; p++) fp[p].shift += sig - top;
		for (int p = 0; p < np; p++) {
			radix_pass<NT>(&A->x, &A->y, 2, &B->x, &B->y, 2, n, kmin, fp[p], s_cnt, s_base);
			uint2* t = A; A = B; B = t;
		}
		if (finish_by_transposition<NT>(&A->x, &A->y, 2, n, 8)) {
			for (int i = tid; i < n; i += NT) emit(i, A[i].y);
			return;
		}
		__syncthreads();
	}
	for (int p = 0; p < ni + nd; p++) {
		const Field f = p < ni ? id_passes[p] : depth_passes[p - ni];
		radix_pass<NT>(&A->x, &A->y, 2, &B->x, &B->y, 2, n, kmin, f, s_cnt, s_base);
		uint2* t = A; A = B; B = t;
	}
	for (int i = tid; i < n; i += NT) emit(i, A[i].y);
}



// Sorts the segment of one tile (range handling + sort_segment).
// (A split of lists beyond the shared-memory capacity into depth sub-ranges sorted one after the other was measured on
// the 3 M-Gaussian shape, lists of ~12 k entries: no gain -- a sub-sort costs ~15 us of barrier-separated phases whatever
// its length, so sequential sub-sorts inside one CTA lose what the shared-memory passes win.)
template <int NT>
__device__ __forceinline__ void sort_tile(int tile, uint2* __restrict__ ranges, uint2* __restrict__ pairs,
                                          uint2* __restrict__ pairs_alt, uint32_t* __restrict__ point_list, unsigned capacity,
                                          int cap_smem, int id_bits, GeomHeader* hdr, uint32_t* sm,
                                          uint32_t* ids_out = nullptr, int ids_cap = 0)
{
	const int tid = threadIdx.x;
	const uint2 range = ranges[tile];
	unsigned start = range.x, end = range.y;
	if (end > capacity) {      // un-synchronised forward ran out of workspace: stay inside it, flag, caller re-runs
		end = capacity;
		if (tid == 0) {
			hdr->overflow = 1;
			ranges[tile] = make_uint2(min(start, capacity), capacity);   // the render kernels stay in bounds too
		}
		if (start >= end) return;
	}
	const int n = (int)(end - start);
	if (n == 0) return;
	uint2* seg = pairs + start;
	uint32_t* out = point_list + start;
	if (n == 1) {
		if (tid == 0) {
			const uint32_t v = seg[0].y;
			out[0] = v;
			if (ids_out && ids_cap > 0) ids_out[0] = v;
		}
		return;
	}
	sort_segment<NT>(seg, pairs_alt + start, out, n, cap_smem, id_bits, sm, ids_out, ids_cap);
}

__host__ inline size_t sort_smem_bytes(int cap_smem, int threads)
{
	return ((size_t)4 * cap_smem + (size_t)(threads / 32) * kMaxBins + kMaxBins + 64) * sizeof(uint32_t);
}

}  // namespace
}  // namespace gsr
