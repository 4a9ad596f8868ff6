// gsr_b200: B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.
// Shared layouts and device helpers.  See DESIGN.md for the data layout in HBM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define GSR_TILE 16            // tile edge in pixels (reference config.h:16-17 BLOCK_X/BLOCK_Y)
#define GSR_TILE_PIX 256
#define GSR_ALIGN 256          // carve-up alignment inside workspaces
#define GSR_SORT_CHUNK 2048    // longest tile list the 256-thread sort handles as one shared-memory chunk (8 keys per thread);
                               // lists above it are queued for the long-list kernel, lists up to it can be sorted by the
                               // forward compositing kernel itself
#define GSR_EXACT_EXP_DEFAULT 2     // compositing kernels: alpha from the reference's expf in forward and backward (gsr_scene.exact_exp)
#define GSR_LAZY_MIN_DEFAULT 256    // per-tile lists longer than this are ordered on demand by the forward compositing kernel

namespace gsr {

// ---------------------------------------------------------------------------------------------
// Per-Gaussian render record: 48 B, three 16-byte words, written once by the forward preprocess,
// gathered (3 x LDG.128 / cp.async 16) by both render kernels and read by the backward preprocess.
//   q0 = { mean2D.x, mean2D.y, conic.xx, conic.xy }
//   q1 = { conic.yy, opacity, view-space depth, red }
//   q2 = { green, blue, rect (4 x u16: min.x, min.y, max.x, max.y packed in 2 words) }
// ---------------------------------------------------------------------------------------------
struct alignas(16) GaussRec {
	float4 q0, q1, q2;
};

// Per-Gaussian accumulator of the render backward (64 B = two 32-byte sectors): what the reference
// keeps in five zero-filled tensors (rasterize_points.cu:176-180).  Four 16-byte words, each the target
// of ONE vector RED (red.global.add.v4.f32) per (warp, Gaussian):
//   a0 = { dL/dmean2D.x, dL/dmean2D.y (NDC units), dL/dconic.xx, 0 }
//   a1 = { dL/dconic.xy, dL/dconic.yy, 0, 0 }
//   a2 = { dL/dopacity, dL/ddepth, dL/dred, 0 }
//   a3 = { dL/dgreen, dL/dblue, 0, 0 }
struct alignas(16) GaussAcc {
	float4 a0, a1, a2, a3;
};

// Header at the start of the geometry workspace (device memory, 256 B)
struct GeomHeader {
	unsigned int num_rendered;   // R = sum of tiles touched
	unsigned int overflow;       // set when R exceeded the binning capacity handed to stage B
	unsigned int num_visible;    // Gaussians with radii > 0
	unsigned int fwd_blocks_done;
	unsigned int bwd_blocks_done;
	unsigned int max_tile_count; // longest per-tile list
	unsigned int num_long_tiles; // tiles queued for the long-list sort kernel (> GSR_SORT_CHUNK entries)
	unsigned int spin_timeout;   // a bounded wait gave up: 1 = tile flag of the forward, 2 = upstream_ready word (compositing backward), 3 = a bulk copy of input rows (preprocess)
	unsigned int pad[64 - 8];
};
static_assert(sizeof(GeomHeader) == 256, "header must be 256 B");

__host__ __device__ inline size_t align_up(size_t v, size_t a = GSR_ALIGN) { return (v + a - 1) / a * a; }

// Geometry workspace (one per forward call, kept for the backward):
struct GeomView {
	GeomHeader* hdr;
	uint32_t* tile_count;    // [tiles]  instances per tile (integer REDs in the preprocess); zeroed with the header
	uint32_t* tile_cursor;   // [tiles]  write cursor of the scatter pass (starts at ranges[tile].x)
	uint32_t* tile_done;     // [tiles]  deepest contributor + 1 (never 0) once the forward compositing CTA of the tile has published its results (cleared with the
	                         //          header): lets a backward launched as a programmatic dependent start tile by tile
	uint2* ranges;           // [tiles]  (start, end) of every tile's list inside point_list
	uint32_t* long_tiles;    // [tiles]  ids of the tiles whose lists go to the long-list sort kernel
	GaussRec* rec;           // [P]
	GaussAcc* acc;           // [P]   zeroed by the forward preprocess, consumed+cleared by backward
	uint32_t* tiles_touched; // [P]
	uint8_t* clamped;        // [P]   bit ch set when SH colour channel ch was clamped at 0
	float* tau_partial;      // [ceil(P/256) * 8] per-block pose-gradient partials
};

__host__ __device__ inline size_t geom_bytes(size_t P, size_t tiles)
{
	size_t s = sizeof(GeomHeader);
	s += align_up(tiles * 4) * 4 + align_up(tiles * 8);
	s += align_up(P * sizeof(GaussRec));
	s += align_up(P * sizeof(GaussAcc));
	s += align_up(P * 4);
	s += align_up(P);
	s += align_up(((P + 255) / 256) * 8 * sizeof(float));
	return s + GSR_ALIGN;
}

__host__ __device__ inline GeomView geom_view(void* base, size_t P, size_t tiles)
{
	char* p = (char*)align_up((size_t)base);
	GeomView g;
	g.hdr = (GeomHeader*)p; p += sizeof(GeomHeader);
	g.tile_count = (uint32_t*)p; p += align_up(tiles * 4);      // directly behind the header: one memset clears both
	g.tile_cursor = (uint32_t*)p; p += align_up(tiles * 4);
	g.tile_done = (uint32_t*)p; p += align_up(tiles * 4);
	g.ranges = (uint2*)p; p += align_up(tiles * 8);
	g.long_tiles = (uint32_t*)p; p += align_up(tiles * 4);
	g.rec = (GaussRec*)p; p += align_up(P * sizeof(GaussRec));
	g.acc = (GaussAcc*)p; p += align_up(P * sizeof(GaussAcc));
	g.tiles_touched = (uint32_t*)p; p += align_up(P * 4);
	g.clamped = (uint8_t*)p; p += align_up(P);
	g.tau_partial = (float*)p;
	return g;
}

// Image workspace: final transmittance and last contributor of every pixel.
struct ImageView {
	float* final_T;      // [H*W]
	uint32_t* n_contrib; // [H*W]
};
__host__ __device__ inline size_t image_bytes(size_t W, size_t H) { return align_up(W * H * 4) * 2 + GSR_ALIGN; }
__host__ __device__ inline ImageView image_view(void* base, size_t W, size_t H)
{
	char* p = (char*)align_up((size_t)base);
	ImageView v;
	v.final_T = (float*)p; p += align_up(W * H * 4);
	v.n_contrib = (uint32_t*)p;
	return v;
}

// Binning workspace.  point_list[R] and cull_masks are the products (kept for the backward); the rest is scratch.
struct BinView {
	uint32_t* point_list;   // [R]  Gaussian ids sorted by (tile, depth, id)        <- kept
	uint32_t* cull_masks;   // [(R/32 + tiles) * 8]  per tile, group of 32 list positions and warp: the forward's cull
	                        //      ballot (bit i = entry 32 g + i may touch the warp's 8x4 pixel block)   <- kept
	uint2* pairs;           // [R]  (depth key, Gaussian id) scattered into tile segments, unordered inside a segment
	uint2* pairs_alt;       // [R]  ping-pong partner for tiles too long to sort in shared memory
};
__host__ __device__ inline size_t cull_mask_words(size_t R, size_t tiles) { return (R / 32 + tiles + 1) * 8; }
// first mask word of a tile: groups of different tiles never overlap because a tile with n entries owns
// ceil(n / 32) <= n / 32 + 1 groups starting at (range.x / 32 + tile)
__host__ __device__ inline size_t cull_mask_base(uint32_t range_start, uint32_t tile) { return ((size_t)(range_start >> 5) + tile) * 8; }
__host__ __device__ inline size_t binning_bytes(size_t R, size_t tiles)
{
	return align_up(R * 4) + align_up(cull_mask_words(R, tiles) * 4) + 2 * align_up(R * 8) + GSR_ALIGN;
}
__host__ __device__ inline BinView bin_view(void* base, size_t R, size_t tiles)
{
	char* p = (char*)align_up((size_t)base);
	BinView b;
	b.point_list = (uint32_t*)p; p += align_up(R * 4);
	b.cull_masks = (uint32_t*)p; p += align_up(cull_mask_words(R, tiles) * 4);
	b.pairs = (uint2*)p; p += align_up(R * 8);
	b.pairs_alt = (uint2*)p;
	return b;
}

// Opt-in to more than 48 KB of dynamic shared memory: a per-function, per-DEVICE attribute.  The sizes already granted
// are remembered per device (a process may drive several GPUs), so the driver call happens once per (kernel, device).
struct SmemAttrCache {
	size_t granted[64] = {};
};
template <typename Kernel>
inline void ensure_dynamic_smem(Kernel kernel, size_t bytes, SmemAttrCache& cache)
{
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& g = cache.granted[dev & 63];
	if (bytes > g) {
		cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
		g = bytes;
	}
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// ---- bulk asynchronous copies (TMA unit, cp.async.bulk) with transaction-counting mbarriers ----
// One elected thread programs the copy of a CONTIGUOUS block (16-byte aligned on both sides, a multiple of 16 bytes); the TMA
// unit moves it while the CTA does something else and signals the mbarrier with the byte count.  Used where a CTA's input
// really is one contiguous block (the rows of 256 consecutive Gaussians in the per-Gaussian kernels); the compositing kernels
// gather 48-byte records through an index list, which the TMA unit does no better than LDGSTS.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy before the first copy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
	             : "memory");
}
// true once the phase with the given parity has completed; bounded polling is the caller's business
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity)
{
	unsigned ok;
	asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
	             : "=r"(ok)
	             : "r"(smem_addr(bar)), "r"(parity)
	             : "memory");
	return ok != 0;
}

// Row index i / w inside a tile rectangle without a per-instance integer division and without special cases in the walk
// loops: floor((i + 0.5) / w) evaluated in fp32 as (float(i) + 0.5f) * fl(1 / w).  (i + 0.5) / w is at least 0.5 / w away
// from the next integer and the two roundings move it by less than (i / w) * 2^-22, so the truncation is exact while a
// rectangle holds fewer than 2^22 tiles (a 32 k x 32 k image has 2^22 tiles in all); checked exhaustively for w <= 512.
// The owning lane computes the reciprocal once per Gaussian.
__device__ __forceinline__ float rect_rcp(uint32_t w) { return __frcp_rn((float)w); }
__device__ __forceinline__ uint32_t rect_row(uint32_t i, float rcp)
{
	return (uint32_t)__fmul_rn(__fadd_rn((float)i, 0.5f), rcp);
}

// Visit every tile of the rectangles held by the lanes of a warp (n = tile count of this lane's Gaussian, 0 = none;
// lo/hi = packed rectangle).  f(tile, key, id) is called once per (Gaussian, tile) with the owner's key / id.  Must be
// called by all 32 lanes.
//
// Load-balanced expansion: the rectangles of the 32 lanes hold anything from 1 to thousands of tiles, so a walk "every
// lane its own rectangle" runs as long as the largest one while most lanes idle (measured at 3 M Gaussians / 1920x1080,
// 34 tiles per Gaussian: 117 warp instructions per Gaussian and pass).  Instead the non-empty rectangles are compacted to
// the low lanes, their tile counts prefix-summed, and the warp walks the concatenated instance range 32 instances per
// step: lane k of a step finds its segment from a bit mask of the segment starts inside the step's window (one
// redux.or + popc), fetches the segment's parameters by shuffle and derives its tile from the offset inside the segment.
constexpr uint32_t kSoloTiles = 32;
template <typename F>
__device__ __forceinline__ void for_each_tile(uint32_t n, uint32_t lo, uint32_t hi, int grid_x, uint32_t key, uint32_t id, F&& f)
{
	constexpr unsigned kAll = 0xffffffffu;
	const unsigned lane = threadIdx.x & 31;
	const unsigned E = __ballot_sync(kAll, n != 0);
	if (E == 0) return;
	// compact the non-empty rectangles to lanes 0 .. m-1 (dense lane r takes the r-th non-empty lane's values)
	const int m = __popc(E);
	const int src0 = (int)lane < m ? (int)__fns(E, 0, lane + 1) : 0;
	uint32_t d_n = __shfl_sync(kAll, n, src0);
	if ((int)lane >= m) d_n = 0;
	const uint32_t d_lo = __shfl_sync(kAll, lo, src0), d_hi = __shfl_sync(kAll, hi, src0);
	const uint32_t d_key = __shfl_sync(kAll, key, src0), d_id = __shfl_sync(kAll, id, src0);
	const uint32_t x0 = d_lo & 0xffff, y0 = d_lo >> 16, w = (d_hi & 0xffff) - x0;
	const float rcp = rect_rcp(w);
	const uint32_t first = y0 * (uint32_t)grid_x + x0;
	uint32_t inc = d_n;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(kAll, inc, o);
		if ((int)lane >= o) inc += t;
	}
	const uint32_t off = inc - d_n;                           // strictly increasing over the dense lanes
	const uint32_t total = __shfl_sync(kAll, inc, 31);
	unsigned before = 0;                                      // segments that start in front of the window
	for (uint32_t base = 0; base < total; base += 32) {
		const uint32_t rel = off - base;                      // wraps (>= 32) for segments that started earlier
		const unsigned starts = __reduce_or_sync(kAll, ((int)lane < m && rel < 32u) ? (1u << rel) : 0u);
		const unsigned c = before + __popc(starts & (kAll >> (31 - lane)));      // segments starting at or before position base + lane
		const int src = (int)c - 1;                           // >= 0: position 0 belongs to dense lane 0
		const uint32_t s_off = __shfl_sync(kAll, off, src);
		const uint32_t s_w = __shfl_sync(kAll, w, src);
		const float s_rcp = __shfl_sync(kAll, rcp, src);
		const uint32_t s_first = __shfl_sync(kAll, first, src);
		const uint32_t s_key = __shfl_sync(kAll, d_key, src), s_id = __shfl_sync(kAll, d_id, src);
		const uint32_t k = base + lane;
		if (k < total) {
			const uint32_t i = k - s_off;
			const uint32_t ty = rect_row(i, s_rcp), tx = i - ty * s_w;
			f(s_first + ty * (uint32_t)grid_x + tx, s_key, s_id);
		}
		before += __popc(starts);
	}
}

// Cooperative vectorised load of ROWS rows of a [P,3] fp32 array (from row0) into shared memory, by the ROWS threads that
// share threadIdx.x / ROWS: the whole 256-thread CTA (ROWS = 256) or one warp (ROWS = 32, synchronised by __syncwarp only).
template <int ROWS = 256>
__device__ __forceinline__ void load_rows3(const float* __restrict__ g, int row0, int P, float* s, bool vec_ok)
{
	const int t = threadIdx.x & (ROWS - 1);
	const int n = max(0, min(ROWS, P - row0)) * 3;
	const float* src = g + (size_t)row0 * 3;
	if (vec_ok) {
		const int n4 = n >> 2;
		const float4* src4 = reinterpret_cast<const float4*>(src);
		for (int i = t; i < n4; i += ROWS) reinterpret_cast<float4*>(s)[i] = __ldg(src4 + i);
		for (int i = (n4 << 2) + t; i < n; i += ROWS) s[i] = __ldg(src + i);
	} else {
		for (int i = t; i < n; i += ROWS) s[i] = __ldg(src + i);
	}
}

// The same for one warp's 32 rows in two halves, so that the loads can be issued together with the thread's other loads
// and the shared-memory deposit (which waits for them) comes behind all of them: fetch -> registers, deposit -> shared.
struct Rows3Regs {
	float4 v;      // float4 number `lane` of the slice (vector path)
	float t[3];    // scalar path / tail: elements lane, lane + 32, lane + 64 behind the vector part
};
__device__ __forceinline__ Rows3Regs fetch_rows3_warp(const float* __restrict__ g, int row0, int P, bool vec_ok)
{
	const int lane = threadIdx.x & 31;
	const int n = max(0, min(32, P - row0)) * 3;
	const float* src = g + (size_t)row0 * 3;
	const int n4 = vec_ok ? (n >> 2) : 0;
	Rows3Regs r;
	r.v = lane < n4 ? __ldg(reinterpret_cast<const float4*>(src) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const int i = (n4 << 2) + lane + 32 * k;
		r.t[k] = i < n ? __ldg(src + i) : 0.f;
	}
	return r;
}
__device__ __forceinline__ void deposit_rows3_warp(float* s, const Rows3Regs& r, int row0, int P, bool vec_ok)
{
	const int lane = threadIdx.x & 31;
	const int n = max(0, min(32, P - row0)) * 3;
	const int n4 = vec_ok ? (n >> 2) : 0;
	if (lane < n4) reinterpret_cast<float4*>(s)[lane] = r.v;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const int i = (n4 << 2) + lane + 32 * k;
		if (i < n) s[i] = r.t[k];
	}
}

// The mirror image: ROWS rows of a [P,3] fp32 array staged in shared memory (row r at s[3r..3r+2]) go out as 16-byte
// stores (full sectors) instead of three scalar stores per thread at a 12-byte stride.  accumulate: dst += rows (the
// group owns its rows; views of a window run in stream order).  The caller synchronises the group between filling s and this.
// accumulate: 0 = overwrite, 1 = read-modify-write (the caller's views run in stream order), 2 = vector / scalar REDs
// (views of a window running concurrently on several streams add into ONE buffer; red_add_v4 is defined below).
__device__ __forceinline__ void red_add_v4(float4* addr, float4 v);
template <int ROWS = 256>
__device__ __forceinline__ void store_rows3(float* __restrict__ g, int row0, int P, const float* s, bool vec_ok, int accumulate)
{
	const int t = threadIdx.x & (ROWS - 1);
	const int n = max(0, min(ROWS, P - row0)) * 3;
	float* dst = g + (size_t)row0 * 3;
	const int n4 = vec_ok ? (n >> 2) : 0;
	float4* dst4 = reinterpret_cast<float4*>(dst);
	for (int i = t; i < n4; i += ROWS) {
		float4 v = reinterpret_cast<const float4*>(s)[i];
		if (accumulate == 2) {
			if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(dst4 + i, v);
			continue;
		}
		if (accumulate) {
			const float4 o = dst4[i];
			v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
		}
		dst4[i] = v;
	}
	for (int i = (n4 << 2) + t; i < n; i += ROWS) {
		if (accumulate == 2) { if (s[i] != 0.f) atomicAdd(dst + i, s[i]); }
		else dst[i] = accumulate ? dst[i] + s[i] : s[i];
	}
}

// Optional phase probe (compile with -DGSR_PHASE_PROBE): thread 0 of every CTA records %globaltimer at phase
// boundaries; tools/phase_probe.py reads the table through gsr_debug_probe().  Not compiled into the product build.
#ifdef GSR_PHASE_PROBE
static __device__ unsigned long long g_probe[4096][8];      // one table per translation unit (= per probed kernel)
__device__ __forceinline__ void probe(int slot)
{
	if (threadIdx.x == 0 && blockIdx.x < 4096) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
		g_probe[blockIdx.x][slot] = t;
	}
}
#define GSR_PROBE(k, s) probe(s)
#define GSR_PROBE_READER(name) \
	int name(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_probe, sizeof(g_probe)); }
#else
#define GSR_PROBE(k, s)
#define GSR_PROBE_READER(name) \
	int name(unsigned long long*) { return -1; }
#endif

// vector reduction to global memory: one 16-byte RED instead of four scalar atomics (sm_90+)
__device__ __forceinline__ void red_add_v4(float4* addr, float4 v)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
	             : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// once every CTA of its predecessor has executed launch_dependents (or exited); grid_dependency_wait blocks until the
// predecessor has completed and its memory is visible.  Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// cp.async 16-byte global->shared (LDGSTS), used to stage gathered Gaussian records
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
	unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace gsr
