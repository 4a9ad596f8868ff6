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
// over the 8x4 rectangle, with a conservative margin) and ballot; the warp then evaluates only the
// survivors (typically < 1/3 of the entries) with the reference's per-pair decisions.
// Every per-pair decision is predicated, so the warp stays converged and can vote its own early
// termination (warp ballot) and aggregate the n_touched integer atomics to one RED per chunk and warp.
#include "render_common.cuh"

namespace gsr {

namespace {

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

}  // namespace

void launch_render_forward(const Scene& s, const GeomView& g, const BinView& b, const ImageView& im, float* out_color,
                           float* out_depth, float* out_opacity, int* n_touched, cudaStream_t stream)
{
	const int tiles = s.grid_x * s.grid_y;
	if (tiles == 0) return;
	render_forward_kernel<<<tiles, 256, 0, stream>>>(g.ranges, b.point_list, g.rec, s.W, s.H, s.grid_x, s.background,
	                                                  im.final_T, im.n_contrib, out_color, out_depth, out_opacity, n_touched);
}

}  // namespace gsr
