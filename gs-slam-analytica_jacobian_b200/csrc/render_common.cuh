// Helpers shared by the forward and backward compositing kernels.
#pragma once
#include "gsr_params.h"

namespace gsr {

// 2^x, one MUFU.EX2 (relative error ~2^-22)
__device__ __forceinline__ float gsr_exp2(float x)
{
	float r;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

__device__ __forceinline__ float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// thread -> pixel: each warp owns an 8x4 pixel block of the 16x16 tile (lane&7 -> x, lane>>3 -> y)
__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int& px, int& py)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	px = tile_x * GSR_TILE + (warp & 1) * 8 + (lane & 7);
	py = tile_y * GSR_TILE + (warp >> 1) * 4 + (lane >> 3);
}

// Conservative test: can the Gaussian (mean q0.xy, conic q0.z q0.w q1.x, opacity q1.y) reach
// alpha >= 1/255 at any pixel centre inside [x0,x1] x [y0,y1] ?  Returns false only when the exact
// per-pixel test (fp32) is guaranteed to reject every pixel of the block.
// alpha = min(0.99, o exp(-q)) >= 1/255  <=>  q <= ln(255 o), q = 0.5 (A dx^2 + C dy^2) + B dx dy convex;
// its minimum over the rectangle is 0 when the mean is inside, else it lies on an edge facing the mean
// (any segment from the mean into the rectangle crosses such an edge and q grows along it).
__device__ __forceinline__ bool may_touch(const float4 q0, const float4 q1, float x0, float y0, float x1, float y1)
{
	const float mx = q0.x, my = q0.y, A = q0.z, B = q0.w, Cc = q1.x, o = q1.y;
	const float cxp = fminf(fmaxf(mx, x0), x1), cyp = fminf(fmaxf(my, y0), y1);
	const float dx = mx - cxp, dy = my - cyp;   // 0 along an axis where the mean lies within the block
	float qmin = 0.f, S = 0.f;
	if (dx != 0.f || dy != 0.f) {
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
	const float tau = __logf(255.0f * o);
	const float thr = tau + 1e-3f * fabsf(tau) + 1e-2f + 1e-5f * S;   // margin >> fp32 rounding of either side
	const bool convex = (A > 0.f) && (Cc > 0.f) && (A * Cc > B * B);
	return !(qmin > thr) || !convex;                                  // NaNs fall through to "keep"
}

// The Gaussian falloff exponent, written with the reference's expression tree (forward.cu:478-484) so that it
// contracts to the same FMUL/FFMA sequence: power is then BIT-IDENTICAL to the reference's, and the three
// threshold decisions per pair (power > 0, alpha < 1/255, T < 1e-4) can only flip through the last bits of exp.
__device__ __forceinline__ float falloff_power(float A, float B, float Cc, float dx, float dy)
{
	return -0.5f * (A * dx * dx + Cc * dy * dy) - B * dx * dy;
}
// exp(x) as ex2.approx(x * log2 e), like __expf; .ftz: results below 2^-126 become 0, which the alpha >= 1/255
// test rejects anyway, and saves the three denormal-fix-up instructions of the non-ftz form.
__device__ __forceinline__ float gsr_exp(float x) { return gsr_exp2(x * 1.4426950408889634f); }

// Per-warp queue of the entries that survived the cull, in list order.  16-byte words, broadcast-read by the
// pixel threads:   w0 = { mean.x, mean.y, conic.xx, conic.xy }   w1 = { conic.yy, opacity, red, green }
//                  w2 = { blue, depth, list position (int bits), Gaussian id (int bits) }
struct alignas(16) QueueRec {
	float4 w0, w1, w2;
};

}  // namespace gsr
