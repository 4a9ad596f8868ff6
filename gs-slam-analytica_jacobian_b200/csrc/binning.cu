// Tile binning.  Produces exactly the reference's per-tile lists -- Gaussian ids ordered by
// (tile id, float bits of view depth), ties in ascending Gaussian id -- and the per-tile ranges
// (reference: InclusiveSum + duplicateWithKeys + cub::DeviceRadixSort::SortPairs over 64-bit keys, 6 passes
// of 24 B/instance + identifyTileRanges, cuda_rasterizer/rasterizer_impl.cu:70-138, 327-368), with a
// tile-major pipeline:
//
//   (preprocess)  every visible Gaussian adds 1 to the counters of the tiles in its rectangle (integer
//                 REDs); the last preprocess CTA scans the counters into ranges[tile] and scatter cursors.
//   1. scatter    two lanes walk each Gaussian's rectangle (large rectangles: the whole warp by load-balanced
//                 expansion, for_each_tile in gsr_common.cuh), claim a slot in the tile's segment and store (depth
//                 bits, id).  Segments come out contiguous per tile but unordered inside.
//   2. order      BY DEFAULT the forward compositing kernel orders each tile's segment itself, on demand and only as
//                 far as compositing reads it (render.cu); the kernels below run when complete lists
//                 are asked for (gsr_sort_on_demand(0)) and some list is longer than one shared-memory chunk:
//                 one CTA per tile sorts its segment in shared memory.  Fast path (lists that fit one chunk of
//                 8 keys per thread): two stable single-chunk radix passes over the LEADING 16 of the depth bits
//                 that differ inside the tile (min/max), ranks from warp match.any + per-warp counters (no
//                 atomics, no counting sweep), then odd-even transposition sweeps on (depth, id) settle the low
//                 bits and the id order of equal depths.  256-thread CTAs take lists up to 2048 entries; longer
//                 lists are queued by the preprocess scan and sorted by a second launch of 512-thread CTAs (same
//                 code; one chunk up to 4096 entries, chunked passes through global ping-pong buffers beyond).
//                 If the finisher does not converge (masses of equal depths) the list takes the general path:
//                 stable LSD passes over the id digits, then every differing depth bit.
//                 The result is the unique (tile, depth, id) order, i.e. the reference's list.
//
// Every kernel reads num_rendered-dependent quantities from device memory, so the stage runs without the reference's
// blocking D2H read (rasterizer_impl.cu:331).
#include "tile_sort.cuh"

namespace gsr {

namespace {

// ---- 1. scatter -------------------------------------------------------------------------------------
// 256 Gaussians per CTA, TWO lanes per Gaussian (512 threads): the walk over a Gaussian's tile rectangle is a chain
// of dependent shared-memory atomics / stores per step, so its latency, not its instruction count, sets the pace --
// two lanes halve the chain (the largest rectangle among the warp's Gaussians) and double the warps in flight; four
// lanes are more than one wave of CTAs at P = 100 k.  One thread per Gaussian with the balanced walk for every
// rectangle was measured too: 5-8 % slower here (the kernel is bound by its partial-sector stores).
#ifndef GSR_SCATTER_LANES
#define GSR_SCATTER_LANES 2
#endif
constexpr int kScatterLanes = GSR_SCATTER_LANES;      // (experiments: -DGSR_SCATTER_LANES=1|4)
#ifndef GSR_SCATTER_GAUSS
#define GSR_SCATTER_GAUSS 256
#endif
constexpr int kScatterGauss = GSR_SCATTER_GAUSS;             // Gaussians per CTA
constexpr int kScatterThreads = kScatterGauss * kScatterLanes;

// visit the tiles of this thread's share of its Gaussian's rectangle: steps sub, sub + kScatterLanes, ... of a row-major
// walk (rectangles above kSoloTiles tiles: the whole warp, for the leading lane's Gaussian)
template <typename F>
__device__ __forceinline__ void for_each_tile_quad(uint32_t n, uint32_t lo, uint32_t hi, int grid_x, uint32_t key, uint32_t id,
                                                   int sub, F&& f)
{
	const uint32_t x0 = lo & 0xffff, y0 = lo >> 16, w = (hi & 0xffff) - x0;
	if (n != 0 && n <= kSoloTiles) {
		const float rcp = rect_rcp(w);
		for (uint32_t i = sub; i < n; i += kScatterLanes) {
			const uint32_t ty = rect_row(i, rcp), tx = i - ty * w;
			f((y0 + ty) * grid_x + (x0 + tx), key, id);
		}
	}
	for_each_tile((n > kSoloTiles && sub == 0) ? n : 0u, lo, hi, grid_x, key, id, f);
}

// ORDERED (gsr_scene.spatial_order): the CTA takes 256 consecutive entries of a screen-coherent permutation instead of 256
// consecutive Gaussians.  Their rectangles then cover a small BOX of tiles with dozens of instances per tile, so the CTA
// histograms in box-local shared memory (no sweep over every tile of the image), claims ONE slice per box tile (C2: ~100
// global atomics per CTA instead of ~2000) and its pairs land in runs of dozens instead of isolated 8-byte stores.  A box
// above kBoxMax tiles (an oversized Gaussian among the 256) sends the CTA down the per-instance path.
constexpr int kBoxMax = 2048;
// -DGSR_DEBUG_CHECKS (never the product; compute-sanitizer is closed on this pool): index checks of the ordered scatter that
// trap instead of writing out of bounds -- run the binning / config-scale tests with such a build (tools/debug_checks.sh)
#ifdef GSR_DEBUG_CHECKS
#define GSR_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define GSR_CHECK(cond) do { } while (0)
#endif

// DEPTH PARTITION (gsr_scene.depth_cut, ORDERED only): a tile's pairs whose depth key is <= cut[tile] ("front") fill its segment
// from the start upwards, the others ("back") from the end downwards -- same segment, same count, same ranges, so every
// result is unchanged; the forward compositing kernel orders the front part first and touches the back part only if the tile
// is still open behind it (render.cu).  The cut is a HINT (the depth a little behind the tile's deepest contributor in the
// previous iteration of the same view, written by the forward): any value is correct, a good one saves the forward the
// sweeps over the 85-98 % of each list that compositing never reads.  Front / back counts share one shared-memory word
// (16 bits each: a CTA holds at most 256 instances per tile), the back cursor is tile_count (= n after the preprocess).
struct DepthPartition {
	const uint32_t* cut;       // [tiles] depth-key bits, or null: no partition
	uint32_t* back_cursor;     // [tiles] starts at the tile's instance count n
	const uint2* ranges;       // [tiles]
};

template <bool ORDERED, bool PART>
__global__ void __launch_bounds__(kScatterThreads, 4)
scatter_kernel(int P, const GaussRec* __restrict__ rec, int grid_x,
               uint32_t* __restrict__ cursor, uint2* __restrict__ pairs, unsigned capacity, GeomHeader* hdr,
               int n_tiles, int use_smem, const uint32_t* __restrict__ order, DepthPartition dp)
{
	GSR_PROBE(1, 0);
	pdl_launch_dependents();      // the forward compositing kernel may take the slots this grid frees (it waits for the segments itself)
	pdl_wait();                   // launched as programmatic dependent of the preprocess: records, ranges and cursors are complete from here
	const int sub = threadIdx.x & (kScatterLanes - 1);
	const int slot = blockIdx.x * kScatterGauss + (threadIdx.x / kScatterLanes);
	int idx = slot;
	if (ORDERED) idx = slot < P ? (int)__ldg(&order[slot]) : P;
	GSR_CHECK(!ORDERED || (idx >= 0 && idx <= P));
	uint32_t n = 0, lo = 0, hi = 0, key = 0;
	if (idx < P) {
		// one round trip: the rectangle (all zero for a culled Gaussian) gives the tile count itself
		const float4 q2 = __ldg(&rec[idx].q2);
		key = __float_as_uint(__ldg(&rec[idx].q1.z));   // raw float bits of the view-space depth (rasterizer_impl.cu:104)
		lo = __float_as_uint(q2.z);
		hi = __float_as_uint(q2.w);
		n = ((hi & 0xffff) - (lo & 0xffff)) * ((hi >> 16) - (lo >> 16));
	}
	extern __shared__ uint32_t s_tile[];          // [2][tiles or kBoxMax]: CTA-local counts, then claimed bases (use_smem != 0)
	int span = n_tiles, gx = grid_x;              // the tile index space of the two passes: the image, or the CTA's box
	uint32_t bx0 = 0, by0 = 0;
	if (ORDERED) {
		__shared__ uint32_t s_box[4][kScatterThreads / 32];
		__shared__ uint32_t s_boxf[4];
		const unsigned kAll = 0xffffffffu;
		const uint32_t m0 = __reduce_min_sync(kAll, n ? (lo & 0xffff) : 0xffffu), m1 = __reduce_min_sync(kAll, n ? (lo >> 16) : 0xffffu);
		const uint32_t m2 = __reduce_max_sync(kAll, n ? (hi & 0xffff) : 0u), m3 = __reduce_max_sync(kAll, n ? (hi >> 16) : 0u);
		if ((threadIdx.x & 31) == 0) {
			const int w = threadIdx.x >> 5;
			s_box[0][w] = m0; s_box[1][w] = m1; s_box[2][w] = m2; s_box[3][w] = m3;
		}
		__syncthreads();
		if (threadIdx.x < 4) {
			uint32_t v = s_box[threadIdx.x][0];
			for (int w = 1; w < kScatterThreads / 32; w++) v = threadIdx.x < 2 ? min(v, s_box[threadIdx.x][w]) : max(v, s_box[threadIdx.x][w]);
			s_boxf[threadIdx.x] = v;
		}
		__syncthreads();
		bx0 = s_boxf[0]; by0 = s_boxf[1];
		const uint32_t bx1 = s_boxf[2], by1 = s_boxf[3];
		if (bx1 <= bx0 || by1 <= by0) return;      // nothing visible among the CTA's Gaussians
		const uint32_t bw = bx1 - bx0, bh = by1 - by0;
		if (bw * bh <= (uint32_t)kBoxMax) {
			// rectangles relative to the box: the walks below then yield box-local tile indices
			if (n) {
				lo = ((lo & 0xffff) - bx0) | (((lo >> 16) - by0) << 16);
				hi = ((hi & 0xffff) - bx0) | (((hi >> 16) - by0) << 16);
			}
			span = (int)(bw * bh);
			gx = (int)bw;
			use_smem = 1;
		} else {
			use_smem = 0;
			bx0 = by0 = 0;
		}
	}
	uint32_t* s_cnt = s_tile;
	uint32_t* s_base = s_tile + (ORDERED ? kBoxMax : n_tiles);
	constexpr bool part = ORDERED && PART;      // (its own instantiation: the plain kernels carry none of it)
	uint32_t* s_cut = s_tile + 2 * kBoxMax;        // [kBoxMax] cut of every box tile          (part only)
	uint32_t* s_base_b = s_tile + 3 * kBoxMax;     // [kBoxMax] last slot of the CTA's back slice (part only)
	auto global_tile = [&](uint32_t t) {
		if (!ORDERED) return t;
		const uint32_t ty = t / (uint32_t)gx;
		return (by0 + ty) * (uint32_t)grid_x + bx0 + (t - ty * (uint32_t)gx);
	};
	if (use_smem) {
		// pass 1: CTA-local tile histogram; then ONE global atomic per (CTA, touched tile) claims a contiguous
		// slice of the tile's segment (coalesced over consecutive tiles) instead of one atomic per instance
		for (int t = threadIdx.x; t < span; t += kScatterThreads) {
			s_cnt[t] = 0;
			if (part) s_cut[t] = __ldg(&dp.cut[global_tile((uint32_t)t)]);
		}
		__syncthreads();
		GSR_PROBE(1, 1);
		for_each_tile_quad(n, lo, hi, gx, part ? key : 0u, 0u, sub, [&](uint32_t tile, uint32_t g_key, uint32_t) {
			GSR_CHECK(tile < (uint32_t)span);
			atomicAdd(&s_cnt[tile], (part && g_key > s_cut[tile]) ? 0x10000u : 1u);
		});
		__syncthreads();
		GSR_PROBE(1, 2);
		for (int t = threadIdx.x; t < span; t += kScatterThreads) {
			const uint32_t c = s_cnt[t];
			if (c) {
				const uint32_t gt = global_tile((uint32_t)t);
				GSR_CHECK(gt < (uint32_t)n_tiles);
				const uint32_t cf = c & 0xffffu, cb = c >> 16;
				if (cf) s_base[t] = atomicAdd(&cursor[gt], cf);
				if (cb) {      // back slice: slots count down from the end of the segment
					const uint2 r = __ldg(&dp.ranges[gt]);
					const uint32_t k = atomicAdd(&dp.back_cursor[gt], cb) - (r.y - r.x);
					s_base_b[t] = r.y - 1u - k;
				}
			}
			s_cnt[t] = 0;
		}
		__syncthreads();
	}
	GSR_PROBE(1, 3);
	bool overflow = false;
	// pass 2: claim a slot per (Gaussian, tile) inside the CTA's slice and store the pair
	for_each_tile_quad(n, lo, hi, gx, key, (uint32_t)idx, sub, [&](uint32_t tile, uint32_t g_key, uint32_t g_id) {
		GSR_CHECK(tile < (uint32_t)(use_smem ? span : n_tiles));
		uint32_t pos;
		if (use_smem) {
			const bool back = part && g_key > s_cut[tile];
			const uint32_t k = atomicAdd(&s_cnt[tile], back ? 0x10000u : 1u);
			pos = back ? s_base_b[tile] - (k >> 16) : s_base[tile] + (k & 0xffffu);
		} else if (part && g_key > __ldg(&dp.cut[tile])) {      // oversized box: one atomic per instance (tile is a global index here)
			const uint2 r = __ldg(&dp.ranges[tile]);
			pos = r.y - 1u - (atomicAdd(&dp.back_cursor[tile], 1u) - (r.y - r.x));
		} else {
			pos = atomicAdd(&cursor[tile], 1u);
		}
		if (pos < capacity) pairs[pos] = make_uint2(g_key, g_id);
		else overflow = true;
	});
	if (overflow) hdr->overflow = 1;
	GSR_PROBE(1, 4);
}

// ---- screen-coherent permutation of the Gaussians (gsr_spatial_order): counting sort by home tile ----
__device__ __forceinline__ uint32_t home_bucket(const GaussRec* __restrict__ rec, int idx, int grid_x, int n_tiles)
{
	const float4 q2 = __ldg(&rec[idx].q2);
	const uint32_t lo = __float_as_uint(q2.z), hi = __float_as_uint(q2.w);
	const uint32_t x0 = lo & 0xffff, y0 = lo >> 16, x1 = hi & 0xffff, y1 = hi >> 16;
	if (x1 <= x0 || y1 <= y0) return 0u;      // culled: first bucket
	// the tile at the centre of the rectangle (clamped: records that no forward plan has written must not index out of bounds)
	return min(((y0 + y1 - 1) >> 1) * (uint32_t)grid_x + ((x0 + x1 - 1) >> 1), (uint32_t)n_tiles - 1u);
}
__global__ void __launch_bounds__(256) home_count_kernel(int P, const GaussRec* __restrict__ rec, int grid_x, int n_tiles,
                                                         uint32_t* __restrict__ count)
{
	const int idx = blockIdx.x * 256 + threadIdx.x;
	if (idx < P) atomicAdd(&count[home_bucket(rec, idx, grid_x, n_tiles)], 1u);
}
__global__ void __launch_bounds__(256) bucket_scan_kernel(uint32_t* __restrict__ count, int tiles)      // one CTA: exclusive scan in place
{
	__shared__ unsigned s_red[8];
	const int per = (tiles + 255) / 256;
	const int t0 = min(tiles, (int)threadIdx.x * per), t1 = min(tiles, t0 + per);
	unsigned sum = 0;
	for (int t = t0; t < t1; t++) sum += count[t];
	unsigned inc = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
		if ((threadIdx.x & 31) >= o) inc += v;
	}
	if ((threadIdx.x & 31) == 31) s_red[threadIdx.x >> 5] = inc;
	__syncthreads();
	unsigned run = inc - sum;
	for (int w = 0; w < (int)(threadIdx.x >> 5); w++) run += s_red[w];
	for (int t = t0; t < t1; t++) {
		const unsigned c = count[t];
		count[t] = run;
		run += c;
	}
}
__global__ void __launch_bounds__(256) home_order_kernel(int P, const GaussRec* __restrict__ rec, int grid_x, int n_tiles,
                                                         uint32_t* __restrict__ cursor, uint32_t* __restrict__ order)
{
	const int idx = blockIdx.x * 256 + threadIdx.x;
	if (idx < P) order[atomicAdd(&cursor[home_bucket(rec, idx, grid_x, n_tiles)], 1u)] = (uint32_t)idx;
}

// one CTA per tile; lists queued for the long-list kernel (use_long) are skipped here
__global__ void __launch_bounds__(kSmallThreads, 4)
tile_sort_kernel(uint2* __restrict__ ranges, uint2* __restrict__ pairs, uint2* __restrict__ pairs_alt,
                 uint32_t* __restrict__ point_list, unsigned capacity, int cap_smem, int id_bits, GeomHeader* hdr, int use_long)
{
	extern __shared__ __align__(16) uint32_t sm[];
	if (use_long) {
		const uint2 r = ranges[blockIdx.x];
		const unsigned n = r.y - r.x;
		if (n > (unsigned)kSmallChunk && r.y <= capacity) return;
	}
	sort_tile<kSmallThreads>(blockIdx.x, ranges, pairs, pairs_alt, point_list, capacity, cap_smem, id_bits, hdr, sm);
}

// lists of more than 2048 entries (queued by the scan at the end of the preprocess kernel): 512-thread CTAs, one
// chunk in shared memory up to 4096 entries, global ping-pong beyond
__global__ void __launch_bounds__(kLongThreads, 2)
tile_sort_long_kernel(uint2* __restrict__ ranges, uint2* __restrict__ pairs, uint2* __restrict__ pairs_alt,
                      uint32_t* __restrict__ point_list, unsigned capacity, int id_bits, GeomHeader* hdr,
                      const uint32_t* __restrict__ long_tiles)
{
	extern __shared__ __align__(16) uint32_t sm[];
	const unsigned num = hdr->num_long_tiles;
	for (unsigned i = blockIdx.x; i < num; i += gridDim.x) {
		const uint32_t tile = long_tiles[i];
		if (ranges[tile].y > capacity) continue;      // overflowing step: the 256-thread kernel clamps and flags it
		sort_tile<kLongThreads>((int)tile, ranges, pairs, pairs_alt, point_list, capacity, kLongChunk, id_bits, hdr, sm);
		__syncthreads();
	}
}

}  // namespace

size_t tile_sort_smem_bytes(int cap_smem) { return sort_smem_bytes(cap_smem, kSmallThreads); }

// the scatter partitions every tile's segment by depth and the forward compositing kernel orders the front part first
bool depth_partition_active(const Scene& s)
{
	static const bool off = getenv("GSR_NO_DEPTH_CUT") != nullptr || getenv("GSR_NO_SPATIAL_ORDER") != nullptr;      // A/B switches
	return !off && s.spatial_order != nullptr && s.depth_cut != nullptr;
}

// R_capacity: instance capacity of the binning workspace.  cap_smem: longest tile list sorted in shared memory by the
// 256-thread kernel; max_tile_hint: longest list expected (<= 0: unknown).
int launch_binning(const Scene& s, const GeomView& g, const BinView& b, size_t R_capacity, int cap_smem, long long max_tile_hint,
                   bool fuse_sort, cudaStream_t stream, bool scatter_done, bool behind_preprocess)
{
	if (s.P == 0 || R_capacity == 0) return 0;
	const int tiles = s.grid_x * s.grid_y;
	const int use_smem = tiles <= 16384 ? 1 : 0;     // 2 x tiles x 4 B of shared memory (1920x1080: 64 KB)
	const size_t scatter_smem = use_smem ? 2 * (size_t)tiles * sizeof(uint32_t) : 0;
	static SmemAttrCache scatter_attr;
	if (scatter_smem > 48 * 1024) ensure_dynamic_smem(scatter_kernel<false, false>, scatter_smem, scatter_attr);
	static const bool no_order = getenv("GSR_NO_SPATIAL_ORDER") != nullptr;      // A/B switch for measurements
	if (!scatter_done) {
		// behind_preprocess: the preprocess kernel was launched just before on this stream and releases its dependents at once:
		// the scatter's CTAs are resident when it ends (programmatic dependent launch; they wait for its results themselves)
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3((s.P + kScatterGauss - 1) / kScatterGauss); cfg.blockDim = dim3(kScatterThreads); cfg.stream = stream;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at;
		cfg.numAttrs = behind_preprocess ? 1 : 0;
		const unsigned cap = (unsigned)R_capacity;
		DepthPartition dp;
		dp.cut = nullptr; dp.back_cursor = g.tile_count; dp.ranges = g.ranges;
		if (s.spatial_order && !no_order) {
			if (depth_partition_active(s)) dp.cut = s.depth_cut;
			cfg.dynamicSmemBytes = (dp.cut ? 4 : 2) * kBoxMax * sizeof(uint32_t);
			if (dp.cut)
				cudaLaunchKernelEx(&cfg, scatter_kernel<true, true>, s.P, (const GaussRec*)g.rec, s.grid_x, g.tile_cursor, b.pairs, cap, g.hdr,
				                   tiles, 1, (const uint32_t*)s.spatial_order, dp);
			else
				cudaLaunchKernelEx(&cfg, scatter_kernel<true, false>, s.P, (const GaussRec*)g.rec, s.grid_x, g.tile_cursor, b.pairs, cap, g.hdr,
				                   tiles, 1, (const uint32_t*)s.spatial_order, dp);
		} else {
			cfg.dynamicSmemBytes = scatter_smem;
			cudaLaunchKernelEx(&cfg, scatter_kernel<false, false>, s.P, (const GaussRec*)g.rec, s.grid_x, g.tile_cursor, b.pairs, cap, g.hdr, tiles,
			                   use_smem, (const uint32_t*)nullptr, dp);
		}
	}
	if (fuse_sort) return scatter_done ? 0 : 1;      // the forward compositing kernel sorts its own tile
	int id_bits = 1;
	while (id_bits < 32 && (1ll << id_bits) < (long long)s.P) id_bits++;
	const int use_long = (max_tile_hint <= 0 || max_tile_hint > kSmallChunk) ? 1 : 0;
	const size_t smem = sort_smem_bytes(cap_smem, kSmallThreads);
	static SmemAttrCache sort_attr, long_attr;
	ensure_dynamic_smem(tile_sort_kernel, smem, sort_attr);
	tile_sort_kernel<<<tiles, kSmallThreads, smem, stream>>>(g.ranges, b.pairs, b.pairs_alt, b.point_list, (unsigned)R_capacity,
	                                                         cap_smem, id_bits, g.hdr, use_long);
	if (!use_long) return 2;
	const size_t lsmem = sort_smem_bytes(kLongChunk, kLongThreads);
	ensure_dynamic_smem(tile_sort_long_kernel, lsmem, long_attr);
	tile_sort_long_kernel<<<min(tiles, 2 * 148), kLongThreads, lsmem, stream>>>(g.ranges, b.pairs, b.pairs_alt, b.point_list,
	                                                                         (unsigned)R_capacity, id_bits, g.hdr, g.long_tiles);
	return 3;
}

void launch_spatial_order(const Scene& s, const GeomView& g, uint32_t* order_out, cudaStream_t stream)
{
	const int tiles = s.grid_x * s.grid_y, grid = (s.P + 255) / 256;
	cudaMemsetAsync(g.tile_cursor, 0, (size_t)tiles * sizeof(uint32_t), stream);
	home_count_kernel<<<grid, 256, 0, stream>>>(s.P, g.rec, s.grid_x, tiles, g.tile_cursor);
	bucket_scan_kernel<<<1, 256, 0, stream>>>(g.tile_cursor, tiles);
	home_order_kernel<<<grid, 256, 0, stream>>>(s.P, g.rec, s.grid_x, tiles, g.tile_cursor, order_out);
}

GSR_PROBE_READER(probe_read_scatter)

}  // namespace gsr
