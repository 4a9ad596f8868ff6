// Tile binning.  Produces exactly the reference's per-tile lists -- Gaussian ids ordered by
// (tile id, float bits of view depth), ties in ascending Gaussian id -- and the per-tile ranges
// (reference: duplicateWithKeys + cub::DeviceRadixSort::SortPairs on 64-bit keys + identifyTileRanges,
// cuda_rasterizer/rasterizer_impl.cu:70-138, 339-368), but without ever materialising 64-bit keys:
//
//   1. depth sort   : stable LSD radix sort of the P Gaussians by their 32-bit depth key
//                     (4 one-sweep passes over P elements, culled Gaussians sink to the end);
//   2. scan + emit  : one kernel walks the Gaussians in depth order, turns tiles_touched into
//                     offsets with a decoupled look-back, and emits (tile id u16, Gaussian id u32)
//                     instances warp-cooperatively (coalesced) in (depth, tile) order;
//   3. tile sort    : stable LSD radix sort of the R instances by tile id only (<= 2 passes of 8 bits).
//                     Stability makes every tile's list depth ordered with id-ordered ties, i.e. the
//                     same permutation the reference's 43..45-bit sort yields.
//   4. ranges       : boundaries of equal tile ids.
//
// Traffic per instance: 6 B emit + 12 B/pass, vs 24 B x 6 passes for the 64-bit key sort.
// All kernels read the instance count R from device memory, so the whole stage can run without the
// reference's blocking D2H read of num_rendered (rasterizer_impl.cu:331).
#include "gsr_params.h"

namespace gsr {

namespace {

constexpr uint32_t LB_AGG = 1u << 30;
constexpr uint32_t LB_PREFIX = 2u << 30;
constexpr uint32_t LB_MASK = (1u << 30) - 1;

// ---- global digit histograms of the four depth passes -------------------------------------------
__global__ void __launch_bounds__(256) depth_hist_kernel(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ ghist)
{
	__shared__ uint32_t h[4 * 256];
	for (int i = threadIdx.x; i < 1024; i += 256) h[i] = 0;
	__syncthreads();
	for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
		const uint32_t k = keys[i];
		atomicAdd(&h[k & 255], 1u);
		atomicAdd(&h[256 + ((k >> 8) & 255)], 1u);
		atomicAdd(&h[512 + ((k >> 16) & 255)], 1u);
		atomicAdd(&h[768 + (k >> 24)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 1024; i += 256)
		if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// ---- one stable radix pass (8-bit digit), one-sweep with decoupled look-back -----------------------
template <typename KeyT, bool IOTA, bool WRITE_KEYS>
__global__ void __launch_bounds__(GSR_SORT_THREADS)
onesweep_kernel(const KeyT* __restrict__ kin, KeyT* __restrict__ kout, const uint32_t* __restrict__ vin,
                uint32_t* __restrict__ vout, const unsigned* __restrict__ n_dev, unsigned n_cap, int shift, unsigned mask,
                const uint32_t* __restrict__ ghist, uint32_t* __restrict__ lookback, unsigned* __restrict__ ticket)
{
	__shared__ uint32_t s_cnt[8][256];
	__shared__ uint32_t s_base[256];
	__shared__ uint32_t s_wsum[8];
	__shared__ unsigned s_tile;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = tid; i < 8 * 256; i += GSR_SORT_THREADS) (&s_cnt[0][0])[i] = 0;
	__syncthreads();
	const unsigned tile = s_tile;
	unsigned n = n_cap;
	if (n_dev) n = min(n, *n_dev);
	const unsigned base = tile * GSR_SORT_TILE;
	if (base >= n) return;

	KeyT key[GSR_SORT_ITEMS];
	uint32_t rank[GSR_SORT_ITEMS];
	const unsigned wbase = base + warp * (32 * GSR_SORT_ITEMS) + lane;
#pragma unroll
	for (int i = 0; i < GSR_SORT_ITEMS; i++) {
		const unsigned pos = wbase + i * 32;
		key[i] = pos < n ? kin[pos] : (KeyT)~(KeyT)0;
	}
	const unsigned lt = (1u << lane) - 1;
#pragma unroll
	for (int i = 0; i < GSR_SORT_ITEMS; i++) {
		const unsigned d = ((unsigned)key[i] >> shift) & mask;
		const unsigned peers = __match_any_sync(0xffffffffu, d);
		const uint32_t pre = s_cnt[warp][d];
		__syncwarp();
		rank[i] = pre + __popc(peers & lt);
		if (lane == 31 - __clz(peers)) s_cnt[warp][d] = pre + __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	// thread d: exclusive scan over warps of digit d, tile total, global digit base, look-back
	{
		const int d = tid;
		uint32_t total = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) {
			const uint32_t c = s_cnt[w][d];
			s_cnt[w][d] = total;
			total += c;
		}
		// exclusive scan of the global histogram across the 256 digits
		const uint32_t gh = ghist[d];
		uint32_t inc = gh;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += t;
		}
		if (lane == 31) s_wsum[warp] = inc;
		__syncthreads();
		uint32_t woff = 0;
#pragma unroll
		for (int w = 0; w < 8; w++)
			if (w < warp) woff += s_wsum[w];
		const uint32_t gbase = woff + inc - gh;

		// Decoupled look-back.  Flag and value share one 32-bit word, so relaxed gpu-scope accesses suffice;
		// four predecessors are polled per round trip to shorten the latency chain.
		uint32_t excl = 0;
		uint32_t* my = lookback + (size_t)tile * 256 + d;
		if (tile == 0) {
			st_relaxed_u32(my, LB_PREFIX | total);
		} else {
			st_relaxed_u32(my, LB_AGG | total);
			int t = (int)tile - 1;
			bool stop = false;
			while (!stop) {
				uint32_t w[4];
#pragma unroll
				for (int k = 0; k < 4; k++) w[k] = (t - k >= 0) ? ld_relaxed_u32(lookback + (size_t)(t - k) * 256 + d) : LB_PREFIX;
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (stop || (w[k] >> 30) == 0) break;
					excl += w[k] & LB_MASK;
					t--;
					if ((w[k] >> 30) == 2) stop = true;
				}
			}
			st_relaxed_u32(my, LB_PREFIX | (excl + total));
		}
		s_base[d] = gbase + excl;
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < GSR_SORT_ITEMS; i++) {
		const unsigned pos = wbase + i * 32;
		if (pos < n) {
			const unsigned d = ((unsigned)key[i] >> shift) & mask;
			const unsigned dst = s_base[d] + s_cnt[warp][d] + rank[i];
			if (WRITE_KEYS) kout[dst] = key[i];
			vout[dst] = IOTA ? pos : vin[pos];
		}
	}
}

// ---- scan + emit ----------------------------------------------------------------------------------
// order[k]: Gaussian ids in depth order.  One block = 256 consecutive ranks.
__global__ void __launch_bounds__(256)
scan_emit_kernel(const uint32_t* __restrict__ order, int P, const GaussRec* __restrict__ rec, int grid_x,
                 uint16_t* __restrict__ inst_tile, uint32_t* __restrict__ inst_val, unsigned R_capacity,
                 uint32_t* __restrict__ status, uint32_t* __restrict__ ghist_tile /*[2][256]*/, GeomHeader* hdr)
{
	__shared__ uint32_t s_h[512];
	__shared__ uint32_t s_w[8];
	__shared__ uint32_t s_blockbase;
	__shared__ unsigned s_blk;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (tid == 0) s_blk = atomicAdd(&hdr->ticket[6], 1u);
	s_h[tid] = 0; s_h[256 + tid] = 0;
	__syncthreads();
	const unsigned blk = s_blk;
	const int k = blk * 256 + tid;
	uint32_t id = 0, rmin = 0, rmax = 0, ntiles = 0;
	if (k < P) {
		id = order[k];
		const float4 q2 = rec[id].q2;
		rmin = __float_as_uint(q2.z);
		rmax = __float_as_uint(q2.w);
		ntiles = ((rmax & 0xffff) - (rmin & 0xffff)) * ((rmax >> 16) - (rmin >> 16));
	}
	// block exclusive scan of ntiles
	uint32_t inc = ntiles;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	uint32_t woff = 0, btotal = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) {
		if (w < warp) woff += s_w[w];
		btotal += s_w[w];
	}
	if (warp == 0) {
		// warp-parallel decoupled look-back: 32 predecessors per round trip
		uint32_t excl = 0;
		if (blk == 0) {
			if (lane == 0) st_relaxed_u32(status, LB_PREFIX | btotal);
		} else {
			if (lane == 0) st_relaxed_u32(status + blk, LB_AGG | btotal);
			int base = (int)blk - 1;
			while (true) {
				const int t = base - lane;
				const uint32_t w = (t >= 0) ? ld_relaxed_u32(status + t) : LB_PREFIX;
				const unsigned ready = __ballot_sync(0xffffffffu, (w >> 30) != 0);
				const unsigned pref = __ballot_sync(0xffffffffu, (w >> 30) == 2);
				const int p = pref ? __ffs(pref) - 1 : 31;          // nearest predecessor holding a prefix
				const unsigned need = (p >= 31) ? 0xffffffffu : ((2u << p) - 1u);
				if ((ready & need) != need) continue;                // someone in the window is not published yet
				uint32_t v = (lane <= p) ? (w & LB_MASK) : 0u;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
				excl += v;
				if (pref) break;
				base -= 32;
			}
			if (lane == 0) st_relaxed_u32(status + blk, LB_PREFIX | (excl + btotal));
		}
		if (lane == 0) s_blockbase = excl;
	}
	__syncthreads();
	const uint32_t off = s_blockbase + woff + inc - ntiles;
	// warp-cooperative emission: the warp walks its Gaussians in rank order, 32 tiles per step
	unsigned live = __ballot_sync(0xffffffffu, ntiles != 0);
	bool overflow = false;
	while (live) {
		const int src = __ffs(live) - 1;
		live &= live - 1;
		const uint32_t g_id = __shfl_sync(0xffffffffu, id, src);
		const uint32_t g_min = __shfl_sync(0xffffffffu, rmin, src);
		const uint32_t g_max = __shfl_sync(0xffffffffu, rmax, src);
		const uint32_t g_off = __shfl_sync(0xffffffffu, off, src);
		const uint32_t g_n = __shfl_sync(0xffffffffu, ntiles, src);
		const uint32_t x0 = g_min & 0xffff, y0 = g_min >> 16, w = (g_max & 0xffff) - x0;
		for (uint32_t i = lane; i < g_n; i += 32) {
			const uint32_t ty = i / w, tx = i - ty * w;
			const uint32_t tile = (y0 + ty) * grid_x + (x0 + tx);
			const uint32_t dst = g_off + i;
			if (dst < R_capacity) {
				inst_tile[dst] = (uint16_t)tile;
				inst_val[dst] = g_id;
				atomicAdd(&s_h[tile & 255], 1u);
				atomicAdd(&s_h[256 + (tile >> 8)], 1u);
			} else {
				overflow = true;
			}
		}
	}
	if (overflow) hdr->overflow = 1;
	__syncthreads();
	if (s_h[tid]) atomicAdd(&ghist_tile[tid], s_h[tid]);
	if (s_h[256 + tid]) atomicAdd(&ghist_tile[256 + tid], s_h[256 + tid]);
}

// ---- ranges ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint16_t* __restrict__ tiles, const unsigned* __restrict__ n_dev, unsigned n_cap, uint2* __restrict__ ranges)
{
	const unsigned n = min(*n_dev, n_cap);
	for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
		const uint32_t cur = tiles[i];
		if (i == 0) ranges[cur].x = 0;
		else {
			const uint32_t prev = tiles[i - 1];
			if (cur != prev) {
				ranges[prev].y = i;
				ranges[cur].x = i;
			}
		}
		if (i == n - 1) ranges[cur].y = n;
	}
}

}  // namespace

// R_capacity: instance capacity of the binning workspace (layout key);  R_bound: host-known upper bound of
// the instance count used to size grids (== R when the caller synchronised, == R_capacity otherwise).
int launch_binning(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, size_t R_capacity,
                   size_t R_bound, cudaStream_t stream)
{
	int launches = 0;
	const int tiles = s.grid_x * s.grid_y;
	cudaMemsetAsync(im.ranges, 0, (size_t)tiles * sizeof(uint2), stream);
	if (s.P == 0) return 0;
	cudaMemsetAsync(b.hist, 0, b.lookback_words * 4, stream);
	const int P = s.P;
	uint32_t* h = b.hist;
	// 1. depth sort
	{
		int hb = (P + 256 * 16 - 1) / (256 * 16);
		if (hb > 148 * 4) hb = 148 * 4;
		depth_hist_kernel<<<hb, 256, 0, stream>>>(g.depth_key, P, h);
		const unsigned nt = (unsigned)sort_tiles(P);
		uint32_t* lb = b.lookback;
		uint32_t* kA = g.depth_key;
		uint32_t* kB = b.gkey_alt;
		onesweep_kernel<uint32_t, true, true><<<nt, GSR_SORT_THREADS, 0, stream>>>(kA, kB, nullptr, b.order, nullptr, P, 0, 255, h, lb, &g.hdr->ticket[0]);
		lb += (size_t)nt * 256;
		onesweep_kernel<uint32_t, false, true><<<nt, GSR_SORT_THREADS, 0, stream>>>(kB, kA, b.order, b.order_alt, nullptr, P, 8, 255, h + 256, lb, &g.hdr->ticket[1]);
		lb += (size_t)nt * 256;
		onesweep_kernel<uint32_t, false, true><<<nt, GSR_SORT_THREADS, 0, stream>>>(kA, kB, b.order_alt, b.order, nullptr, P, 16, 255, h + 512, lb, &g.hdr->ticket[2]);
		lb += (size_t)nt * 256;
		onesweep_kernel<uint32_t, false, false><<<nt, GSR_SORT_THREADS, 0, stream>>>(kB, kA, b.order, b.order_alt, nullptr, P, 24, 255, h + 768, lb, &g.hdr->ticket[3]);
		launches += 5;
	}
	// 2. scan + emit (depth order -> instances)
	int tile_bits = 1;
	while ((1 << tile_bits) < tiles) tile_bits++;
	const int npass = tile_bits > 8 ? 2 : 1;
	uint32_t* V0 = npass == 2 ? b.point_list : b.inst_val_alt;
	uint32_t* V1 = npass == 2 ? b.inst_val_alt : b.point_list;
	scan_emit_kernel<<<(P + 255) / 256, 256, 0, stream>>>(b.order_alt, P, g.rec, s.grid_x, b.inst_tile, V0, (unsigned)R_capacity,
	                                                     b.emit_status, h + 1024, g.hdr);
	launches += 1;
	// 3. tile sort
	if (R_bound > 0) {
		const unsigned nt = (unsigned)sort_tiles(R_bound);
		uint32_t* lb = b.lookback + 4 * sort_tiles(P) * 256;
		const unsigned m0 = (1u << (tile_bits < 8 ? tile_bits : 8)) - 1;
		onesweep_kernel<uint16_t, false, true><<<nt, GSR_SORT_THREADS, 0, stream>>>(
			b.inst_tile, b.inst_tile_alt, V0, V1, &g.hdr->num_rendered, (unsigned)R_capacity, 0, m0, h + 1024, lb, &g.hdr->ticket[4]);
		const uint16_t* sorted_tiles = b.inst_tile_alt;
		if (npass == 2) {
			lb += (size_t)sort_tiles(R_capacity) * 256;
			onesweep_kernel<uint16_t, false, true><<<nt, GSR_SORT_THREADS, 0, stream>>>(
				b.inst_tile_alt, b.inst_tile, V1, V0, &g.hdr->num_rendered, (unsigned)R_capacity, 8, 255, h + 1280, lb, &g.hdr->ticket[5]);
			sorted_tiles = b.inst_tile;
		}
		int rb = (int)((R_bound + 256 * 8 - 1) / (256 * 8));
		if (rb > 148 * 8) rb = 148 * 8;
		tile_ranges_kernel<<<rb, 256, 0, stream>>>(sorted_tiles, &g.hdr->num_rendered, (unsigned)R_capacity, im.ranges);
		launches += npass + 1;
	}
	return launches;
}

}  // namespace gsr
