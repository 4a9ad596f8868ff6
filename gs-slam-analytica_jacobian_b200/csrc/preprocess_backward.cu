// Fused per-Gaussian backward: conic -> cov2D -> (cov3D, mean, pose), NDC mean -> (mean, pose),
// depth -> (mean, pose), SH -> (coefficients, mean, pose), cov3D -> (scale, rotation), plus the
// analytic SE(3) pose gradient dL/dtau of the GS-SLAM paper reduced to ONE 6-vector per view.
// Replaces reference computeCov2DCUDA (cuda_rasterizer/backward.cu:150-345), preprocessCUDA
// (:494-624), computeColorFromSH (:21-145), computeCov3D (:426-489) and the torch.sum over the
// [P,6] buffer (diff_gaussian_rasterization/__init__.py:162-164): two kernels + eleven zero-filled
// tensors + a reduction there, one kernel and no memset here.  Every output row is written
// (zeros for culled Gaussians), so the caller allocates with torch.empty.
// dL/dtau: warp shuffle -> shared memory -> one partial per CTA -> the last CTA to finish sums the
// partials in index order, so the result is deterministic (the reference's is not needed to be:
// it sums a [P,6] tensor).
// Memory access: every global load of a warp's 32 Gaussians is issued before the first instruction that consumes one;
// [P,3] arrays cross the memory system as 16-byte accesses through warp-private shared-memory row slices; nothing is
// CTA-wide before the pose reduction.  3 M Gaussians: 0.21 ms, 3.3 TB/s of DRAM traffic (ncu, profiles/r1k_ncu_full_C4.md).
#include "gsr_params.h"

namespace gsr {

namespace {

__device__ const float bSH_C0 = 0.28209479177387814f;
__device__ const float bSH_C1 = 0.4886025119029199f;
__device__ const float bSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float bSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// SH backward for one Gaussian (backward.cu:21-145).  dRGB already masked by the clamp flags.
// Writes dL_dsh[0..M) (zeros above the active degree), returns dL/dmean through the view direction.
__device__ __forceinline__ float3 sh_backward(int deg, int M, const float* __restrict__ sh, float3 pos, float3 campos,
                                              const float dRGB[3], float* __restrict__ dL_dsh, int accumulate)
{
	const float3 dir_o = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
	const float len = sqrtf(dir_o.x * dir_o.x + dir_o.y * dir_o.y + dir_o.z * dir_o.z);
	const float x = dir_o.x / len, y = dir_o.y / len, z = dir_o.z / len;
	float ddx = 0.f, ddy = 0.f, ddz = 0.f;
	float w[16];
	w[0] = bSH_C0;
	int ncoef = 1;
	float xx = 0, yy = 0, zz = 0, xy = 0, yz = 0, xz = 0;
	if (deg > 0) {
		w[1] = -bSH_C1 * y; w[2] = bSH_C1 * z; w[3] = -bSH_C1 * x;
		ncoef = 4;
		if (deg > 1) {
			xx = x * x; yy = y * y; zz = z * z; xy = x * y; yz = y * z; xz = x * z;
			w[4] = bSH_C2[0] * xy; w[5] = bSH_C2[1] * yz; w[6] = bSH_C2[2] * (2.f * zz - xx - yy);
			w[7] = bSH_C2[3] * xz; w[8] = bSH_C2[4] * (xx - yy);
			ncoef = 9;
			if (deg > 2) {
				w[9] = bSH_C3[0] * y * (3.f * xx - yy); w[10] = bSH_C3[1] * xy * z;
				w[11] = bSH_C3[2] * y * (4.f * zz - xx - yy); w[12] = bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
				w[13] = bSH_C3[4] * x * (4.f * zz - xx - yy); w[14] = bSH_C3[5] * z * (xx - yy);
				w[15] = bSH_C3[6] * x * (xx - 3.f * yy);
				ncoef = 16;
			}
		}
	}
	for (int i = 0; i < M; i++) {
		const float wi = i < ncoef ? w[i] : 0.f;
		if (accumulate == 2) {
			atomicAdd(&dL_dsh[3 * i + 0], wi * dRGB[0]);
			atomicAdd(&dL_dsh[3 * i + 1], wi * dRGB[1]);
			atomicAdd(&dL_dsh[3 * i + 2], wi * dRGB[2]);
		} else if (accumulate) {
			dL_dsh[3 * i + 0] += wi * dRGB[0];
			dL_dsh[3 * i + 1] += wi * dRGB[1];
			dL_dsh[3 * i + 2] += wi * dRGB[2];
		} else {
			dL_dsh[3 * i + 0] = wi * dRGB[0];
			dL_dsh[3 * i + 1] = wi * dRGB[1];
			dL_dsh[3 * i + 2] = wi * dRGB[2];
		}
	}
	if (deg > 0) {
#pragma unroll
		for (int ch = 0; ch < 3; ch++) {
#define SHC(i) sh[3 * (i) + ch]
			float dx = -bSH_C1 * SHC(3), dy = -bSH_C1 * SHC(1), dz = bSH_C1 * SHC(2);
			if (deg > 1) {
				dx += bSH_C2[0] * y * SHC(4) + bSH_C2[2] * 2.f * -x * SHC(6) + bSH_C2[3] * z * SHC(7) + bSH_C2[4] * 2.f * x * SHC(8);
				dy += bSH_C2[0] * x * SHC(4) + bSH_C2[1] * z * SHC(5) + bSH_C2[2] * 2.f * -y * SHC(6) + bSH_C2[4] * 2.f * -y * SHC(8);
				dz += bSH_C2[1] * y * SHC(5) + bSH_C2[2] * 2.f * 2.f * z * SHC(6) + bSH_C2[3] * x * SHC(7);
				if (deg > 2) {
					dx += bSH_C3[0] * SHC(9) * 3.f * 2.f * xy + bSH_C3[1] * SHC(10) * yz + bSH_C3[2] * SHC(11) * -2.f * xy +
					      bSH_C3[3] * SHC(12) * -3.f * 2.f * xz + bSH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
					      bSH_C3[5] * SHC(14) * 2.f * xz + bSH_C3[6] * SHC(15) * 3.f * (xx - yy);
					dy += bSH_C3[0] * SHC(9) * 3.f * (xx - yy) + bSH_C3[1] * SHC(10) * xz +
					      bSH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + bSH_C3[3] * SHC(12) * -3.f * 2.f * yz +
					      bSH_C3[4] * SHC(13) * -2.f * xy + bSH_C3[5] * SHC(14) * -2.f * yz + bSH_C3[6] * SHC(15) * -3.f * 2.f * xy;
					dz += bSH_C3[1] * SHC(10) * xy + bSH_C3[2] * SHC(11) * 4.f * 2.f * yz +
					      bSH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + bSH_C3[4] * SHC(13) * 4.f * 2.f * xz +
					      bSH_C3[5] * SHC(14) * (xx - yy);
				}
			}
#undef SHC
			ddx += dx * dRGB[ch]; ddy += dy * dRGB[ch]; ddz += dz * dRGB[ch];
		}
	}
	// dnormvdv (auxiliary.h:107-117)
	const float sum2 = dir_o.x * dir_o.x + dir_o.y * dir_o.y + dir_o.z * dir_o.z;
	const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	float3 dm;
	dm.x = ((+sum2 - dir_o.x * dir_o.x) * ddx - dir_o.y * dir_o.x * ddy - dir_o.z * dir_o.x * ddz) * inv32;
	dm.y = (-dir_o.x * dir_o.y * ddx + (sum2 - dir_o.y * dir_o.y) * ddy - dir_o.z * dir_o.y * ddz) * inv32;
	dm.z = (-dir_o.x * dir_o.z * ddx - dir_o.y * dir_o.z * ddy + (sum2 - dir_o.z * dir_o.z) * ddz) * inv32;
	return dm;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB)
preprocess_backward_kernel(Scene s, GeomView g, const int* __restrict__ radii, float* __restrict__ dL_dmeans3D,
                           float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dsh, float* __restrict__ dL_dcolors,
                           float* __restrict__ dL_dopacity, float* __restrict__ dL_dscales, float* __restrict__ dL_drot,
                           float* __restrict__ dL_dcov3D_out, float* __restrict__ dL_dtau, int vec_mask)
{
	// [P,3] arrays pass through shared memory as 32-row slices, one per warp, and cross the memory system as 16-byte
	// accesses: rows[0] means3D in / dL_dmeans3D out, rows[1] scales in / dL_dscales out, rows[2] dL_dmeans2D, rows[3]
	// dL_dcolors or dL_dsh (one coefficient).  Three scalar stores per array at a 12-byte stride touch every sector three
	// times.  Everything up to the pose-gradient reduction is WARP-private (camera matrices included, __syncwarp only): the
	// warps of a CTA drift apart, one warp's loads overlap another's arithmetic, nobody waits at a CTA barrier for the
	// slowest warp (ncu at 3 M Gaussians with CTA-wide staging: 10 of 27 stall cycles per issue at barriers).
	__shared__ __align__(16) float s_rows[4][768];
	__shared__ float s_mat[8][48];
	__shared__ float s_tau[8][6];
	__shared__ bool s_last;
	GSR_PROBE(2, 0);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	float mat_v = 0.f, mat_p = 0.f, mat_r = 0.f;
	if (lane < 16) { mat_v = __ldg(s.viewmatrix + lane); mat_p = __ldg(s.projmatrix + lane); mat_r = __ldg(s.projmatrix_raw + lane); }
	const int row0 = blockIdx.x * 256 + warp * 32;      // first row of this warp
	float* const rows0 = s_rows[0] + 96 * warp;
	float* const rows1 = s_rows[1] + 96 * warp;
	float* const rows2 = s_rows[2] + 96 * warp;
	float* const rows3 = s_rows[3] + 96 * warp;
	// Every global input of this warp's Gaussians is requested up front, before visibility is known and before anything is
	// consumed: the loads are in flight together instead of forming a chain of dependent round trips (rows -> radii ->
	// accumulators), which is what bounds this kernel (few warps per SM, one long dependent computation per thread).
	const Rows3Regs in_mean = fetch_rows3_warp(s.means3D, row0, s.P, vec_mask & 1);
	Rows3Regs in_scale = {};
	if (!s.cov3D_precomp) in_scale = fetch_rows3_warp(s.scales, row0, s.P, vec_mask & 2);
	const float* vm = s_mat[warp];
	const float* pj = s_mat[warp] + 16;
	const float* s_raw = s_mat[warp] + 32;
	const int idx = row0 + lane;
	float tau[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	const bool sh_rows = s.shs && s.M == 1;      // degree 0 (every shipped config): dL_dsh is a [P,3] array too
	float dmean[3] = {0.f, 0.f, 0.f};
	float dm2x = 0.f, dm2y = 0.f;
	float dcol[3] = {0.f, 0.f, 0.f};             // dL_dcolors row, or the dL_dsh row when sh_rows
	float dscale[3] = {0.f, 0.f, 0.f};

	const bool in_range = idx < s.P;
	int radius = 0;
	float4 q_in = make_float4(0.f, 0.f, 0.f, 0.f);
	if (in_range) {
		radius = radii[idx];
		if (!s.cov3D_precomp) q_in = __ldg(reinterpret_cast<const float4*>(s.rotations) + idx);
	}
	// Launched as a programmatic dependent of the compositing backward, this CTA may be resident while that kernel still
	// runs: its own inputs are on their way; the accumulators are read behind the dependency (L2-coherent loads).
	// kEarlyMath (the one-wave build, MINB == 3: every CTA moves into the tail of the compositing backward): the part of the
	// arithmetic that needs no gradient -- covariance, projection, Jacobian -- runs BEFORE the dependency is waited for, so
	// only the gradient half sits behind the compositing backward.  Multi-wave builds keep all loads in one round trip: most
	// of their CTAs start after that kernel has finished, and a second round trip per CTA costs them more.
	constexpr bool kEarlyMath = (MINB == 3);
	GaussAcc a;
	a.a0 = make_float4(0.f, 0.f, 0.f, 0.f); a.a1 = a.a0; a.a2 = a.a0; a.a3 = a.a0;
	if (!kEarlyMath) {
		pdl_wait();
		if (in_range) {
			a.a0 = __ldcg(&g.acc[idx].a0); a.a1 = __ldcg(&g.acc[idx].a1); a.a2 = __ldcg(&g.acc[idx].a2); a.a3 = __ldcg(&g.acc[idx].a3);
		}
	}
	// the first consumers of any load: camera matrices and the [P,3] input rows go from registers to this warp's
	// shared-memory slices (one row per lane)
	if (lane < 16) { s_mat[warp][lane] = mat_v; s_mat[warp][16 + lane] = mat_p; s_mat[warp][32 + lane] = mat_r; }
	deposit_rows3_warp(rows0, in_mean, row0, s.P, vec_mask & 1);
	if (!s.cov3D_precomp) deposit_rows3_warp(rows1, in_scale, row0, s.P, vec_mask & 2);
	__syncwarp();

	if (in_range) {
		float dopac = 0.f;
		float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
		float drot[4] = {0.f, 0.f, 0.f, 0.f};
		const float mx = rows0[3 * lane], my = rows0[3 * lane + 1], mz = rows0[3 * lane + 2];
		float sc_in[3] = {0.f, 0.f, 0.f};
		if (!s.cov3D_precomp) {
#pragma unroll
			for (int i = 0; i < 3; i++) sc_in[i] = rows1[3 * lane + i];
		}
		const bool visible = radius > 0;
		if (visible) {
			// ---- 3D covariance (recomputed; reference re-reads geomState.cov3D) ----
			float c3[6];
			float sc[3] = {0.f, 0.f, 0.f};
			float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
			float Rq[3][3];
			if (s.cov3D_precomp) {
#pragma unroll
				for (int i = 0; i < 6; i++) c3[i] = s.cov3D_precomp[(size_t)idx * 6 + i];
			} else {
				q = q_in;
				const float r = q.x, x = q.y, y = q.z, zq = q.w;
				Rq[0][0] = 1.f - 2.f * (y * y + zq * zq); Rq[0][1] = 2.f * (x * y - r * zq); Rq[0][2] = 2.f * (x * zq + r * y);
				Rq[1][0] = 2.f * (x * y + r * zq); Rq[1][1] = 1.f - 2.f * (x * x + zq * zq); Rq[1][2] = 2.f * (y * zq - r * x);
				Rq[2][0] = 2.f * (x * zq - r * y); Rq[2][1] = 2.f * (y * zq + r * x); Rq[2][2] = 1.f - 2.f * (x * x + y * y);
#pragma unroll
				for (int i = 0; i < 3; i++) sc[i] = s.scale_modifier * sc_in[i];
				float Sg[3][3];
#pragma unroll
				for (int aa = 0; aa < 3; aa++)
#pragma unroll
					for (int bb = 0; bb < 3; bb++) {
						float accv = 0.f;
#pragma unroll
						for (int i = 0; i < 3; i++) accv += (sc[i] * Rq[aa][i]) * (sc[i] * Rq[bb][i]);
						Sg[aa][bb] = accv;
					}
				c3[0] = Sg[0][0]; c3[1] = Sg[0][1]; c3[2] = Sg[0][2]; c3[3] = Sg[1][1]; c3[4] = Sg[1][2]; c3[5] = Sg[2][2];
			}
			// ---- cov2D forward recompute (backward.cu:171-206) ----
			float t[3];
#pragma unroll
			for (int r = 0; r < 3; r++) t[r] = vm[r] * mx + vm[4 + r] * my + vm[8 + r] * mz + vm[12 + r];
			const float pCx = t[0], pCy = t[1];   // unclamped camera-space mean
			const float limx = 1.3f * s.tan_fovx, limy = 1.3f * s.tan_fovy;
			const float txtz = t[0] / t[2], tytz = t[1] / t[2];
			t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
			t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
			const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
			const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
			const float hx = s.focal_x, hy = s.focal_y;
			const float J00 = hx / t[2], J11 = hy / t[2];
			const float J02 = -(hx * t[0]) / (t[2] * t[2]), J12 = -(hy * t[1]) / (t[2] * t[2]);
			float A0[3], A1[3];   // rows of J * R_cw
#pragma unroll
			for (int k = 0; k < 3; k++) {
				A0[k] = J00 * vm[4 * k + 0] + J02 * vm[4 * k + 2];
				A1[k] = J11 * vm[4 * k + 1] + J12 * vm[4 * k + 2];
			}
			const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
			float VA0[3], VA1[3];
#pragma unroll
			for (int i = 0; i < 3; i++) {
				VA0[i] = V[i][0] * A0[0] + V[i][1] * A0[1] + V[i][2] * A0[2];
				VA1[i] = V[i][0] * A1[0] + V[i][1] * A1[1] + V[i][2] * A1[2];
			}
			const float ca = A0[0] * VA0[0] + A0[1] * VA0[1] + A0[2] * VA0[2] + 0.3f;
			const float cb = A0[0] * VA1[0] + A0[1] * VA1[1] + A0[2] * VA1[2];
			const float cc = A1[0] * VA1[0] + A1[1] * VA1[1] + A1[2] * VA1[2] + 0.3f;
			const float denom = ca * cc - cb * cb;
			float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
			const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
			// ---- the gradients arriving from the compositing backward ----
			if (kEarlyMath) {
				pdl_wait();
				a.a0 = __ldcg(&g.acc[idx].a0); a.a1 = __ldcg(&g.acc[idx].a1); a.a2 = __ldcg(&g.acc[idx].a2); a.a3 = __ldcg(&g.acc[idx].a3);
			}
			{
				GaussAcc z;
				z.a0 = make_float4(0.f, 0.f, 0.f, 0.f); z.a1 = z.a0; z.a2 = z.a0; z.a3 = z.a0;
				g.acc[idx] = z;   // consumed: ready for the next backward without a memset
			}
			dm2x = a.a0.x; dm2y = a.a0.y;
			const float dcx = a.a0.z, dcy = a.a1.x, dcz = a.a1.y;
			dopac = a.a2.x;
			const float ddepth = a.a2.y;
			dcol[0] = a.a2.z; dcol[1] = a.a3.x; dcol[2] = a.a3.y;
			if (denom2inv != 0.f) {
				dL_da = denom2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcz);
				dL_dc = denom2inv * (-ca * ca * dcz + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
				dL_db = denom2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcz);
				dcov[0] = A0[0] * A0[0] * dL_da + A0[0] * A1[0] * dL_db + A1[0] * A1[0] * dL_dc;
				dcov[3] = A0[1] * A0[1] * dL_da + A0[1] * A1[1] * dL_db + A1[1] * A1[1] * dL_dc;
				dcov[5] = A0[2] * A0[2] * dL_da + A0[2] * A1[2] * dL_db + A1[2] * A1[2] * dL_dc;
				dcov[1] = 2 * A0[0] * A0[1] * dL_da + (A0[0] * A1[1] + A0[1] * A1[0]) * dL_db + 2 * A1[0] * A1[1] * dL_dc;
				dcov[2] = 2 * A0[0] * A0[2] * dL_da + (A0[0] * A1[2] + A0[2] * A1[0]) * dL_db + 2 * A1[0] * A1[2] * dL_dc;
				dcov[4] = 2 * A0[2] * A0[1] * dL_da + (A0[1] * A1[2] + A0[2] * A1[1]) * dL_db + 2 * A1[1] * A1[2] * dL_dc;
			}
			float dT0[3], dT1[3];
#pragma unroll
			for (int k = 0; k < 3; k++) {
				dT0[k] = 2 * VA0[k] * dL_da + VA1[k] * dL_db;
				dT1[k] = 2 * VA1[k] * dL_dc + VA0[k] * dL_db;
			}
			float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
			for (int j = 0; j < 3; j++) {
				dJ00 += vm[4 * j + 0] * dT0[j];
				dJ02 += vm[4 * j + 2] * dT0[j];
				dJ11 += vm[4 * j + 1] * dT1[j];
				dJ12 += vm[4 * j + 2] * dT1[j];
			}
			const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
			const float gx = x_grad_mul * -hx * tz2 * dJ02;
			const float gy = y_grad_mul * -hy * tz2 * dJ12;
			const float gz = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * t[0]) * tz3 * dJ02 + (2 * hy * t[1]) * tz3 * dJ12;
			// pose through t: [I | -t^x] with the clamped t (backward.cu:275-290)
			tau[0] += gx; tau[1] += gy; tau[2] += gz;
			tau[3] += t[1] * gz - t[2] * gy;
			tau[4] += t[2] * gx - t[0] * gz;
			tau[5] += t[0] * gy - t[1] * gx;
			// mean through t (assignment in the reference, :299)
#pragma unroll
			for (int c = 0; c < 3; c++) dmean[c] = vm[4 * c + 0] * gx + vm[4 * c + 1] * gy + vm[4 * c + 2] * gz;
			// pose through W: dtheta = sum_k R[:,k] x dL/dR[:,k]  (backward.cu:301-345)
#pragma unroll
			for (int k = 0; k < 3; k++) {
				const float dW0 = J00 * dT0[k], dW1 = J11 * dT1[k], dW2 = J02 * dT0[k] + J12 * dT1[k];
				const float c0 = vm[4 * k + 0], c1 = vm[4 * k + 1], c2 = vm[4 * k + 2];
				tau[3] += c1 * dW2 - c2 * dW1;
				tau[4] += c2 * dW0 - c0 * dW2;
				tau[5] += c0 * dW1 - c1 * dW0;
			}
			// ---- NDC mean (backward.cu:522-597) ----
			const float mhx = pj[0] * mx + pj[4] * my + pj[8] * mz + pj[12];
			const float mhy = pj[1] * mx + pj[5] * my + pj[9] * mz + pj[13];
			const float mhw = pj[3] * mx + pj[7] * my + pj[11] * mz + pj[15];
			const float m_w = 1.0f / (mhw + 0.0000001f);
			const float mul1 = mhx * m_w * m_w, mul2 = mhy * m_w * m_w;
#pragma unroll
			for (int k = 0; k < 3; k++)
				dmean[k] += (pj[4 * k] * m_w - pj[4 * k + 3] * mul1) * dm2x + (pj[4 * k + 1] * m_w - pj[4 * k + 3] * mul2) * dm2y;
			{
				const float al = m_w, be = -mhx * m_w * m_w, ga = -mhy * m_w * m_w;
				const float pa = s_raw[0], pb = s_raw[5], pe = s_raw[11];
				const float pC[3] = {pCx, pCy, t[2]};
				const float v1[3] = {al * pa, 0.f, be * pe}, v2[3] = {0.f, al * pb, ga * pe};
				const float c1[3] = {pC[1] * v1[2] - pC[2] * v1[1], pC[2] * v1[0] - pC[0] * v1[2], pC[0] * v1[1] - pC[1] * v1[0]};
				const float c2[3] = {pC[1] * v2[2] - pC[2] * v2[1], pC[2] * v2[0] - pC[0] * v2[2], pC[0] * v2[1] - pC[1] * v2[0]};
#pragma unroll
				for (int i = 0; i < 3; i++) {
					tau[i] += dm2x * v1[i] + dm2y * v2[i];
					tau[3 + i] += dm2x * c1[i] + dm2y * c2[i];
				}
				// ---- depth (backward.cu:603-613) ----
				dmean[0] += ddepth * vm[2]; dmean[1] += ddepth * vm[6]; dmean[2] += ddepth * vm[10];
				tau[2] += ddepth;
				tau[3] += ddepth * pC[1];
				tau[4] += ddepth * -pC[0];
			}
			// ---- SH (backward.cu:618-619) ----
			if (sh_rows) {
				// degree 0: colour = C0 * sh + 0.5 does not depend on the view direction (backward.cu:21-145 with deg = 0)
				if (dL_dcolors) {      // the colour gradient itself is wanted as well (rows[3] carries dL_dsh): direct stores
					float* d = dL_dcolors + 3 * (size_t)idx;
					if (s.accumulate_grads == 2) { atomicAdd(d, dcol[0]); atomicAdd(d + 1, dcol[1]); atomicAdd(d + 2, dcol[2]); }
					else if (s.accumulate_grads) { d[0] += dcol[0]; d[1] += dcol[1]; d[2] += dcol[2]; }
					else { d[0] = dcol[0]; d[1] = dcol[1]; d[2] = dcol[2]; }
				}
				const unsigned cl = g.clamped[idx];
				dcol[0] = (cl & 1) ? 0.f : bSH_C0 * dcol[0]; dcol[1] = (cl & 2) ? 0.f : bSH_C0 * dcol[1];
				dcol[2] = (cl & 4) ? 0.f : bSH_C0 * dcol[2];
			} else if (s.shs) {
				const unsigned cl = g.clamped[idx];
				const float dRGB[3] = {(cl & 1) ? 0.f : dcol[0], (cl & 2) ? 0.f : dcol[1], (cl & 4) ? 0.f : dcol[2]};
				const float3 campos = {s.campos[0], s.campos[1], s.campos[2]};
				const float3 dm = sh_backward(s.D, s.M, s.shs + (size_t)idx * s.M * 3, make_float3(mx, my, mz), campos, dRGB,
				                              dL_dsh + (size_t)idx * s.M * 3, s.accumulate_grads);
				dmean[0] += dm.x; dmean[1] += dm.y; dmean[2] += dm.z;
				tau[0] += -dm.x; tau[1] += -dm.y; tau[2] += -dm.z;
			}
			// ---- cov3D -> scale / rotation (backward.cu:426-489) ----
			if (s.scales) {
				float Mm[3][3];   // M = S * Rq^T
#pragma unroll
				for (int i = 0; i < 3; i++)
#pragma unroll
					for (int j = 0; j < 3; j++) Mm[i][j] = sc[i] * Rq[j][i];
				const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
				                        {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
				                        {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
				float Gm[3][3];
#pragma unroll
				for (int i = 0; i < 3; i++) {
					float dM[3];
#pragma unroll
					for (int j = 0; j < 3; j++) dM[j] = 2.0f * (Mm[i][0] * dS[0][j] + Mm[i][1] * dS[1][j] + Mm[i][2] * dS[2][j]);
					dscale[i] = Rq[0][i] * dM[0] + Rq[1][i] * dM[1] + Rq[2][i] * dM[2];
#pragma unroll
					for (int j = 0; j < 3; j++) Gm[i][j] = sc[i] * dM[j];
				}
				const float r = q.x, x = q.y, y = q.z, zq = q.w;
				drot[0] = 2 * zq * (Gm[0][1] - Gm[1][0]) + 2 * y * (Gm[2][0] - Gm[0][2]) + 2 * x * (Gm[1][2] - Gm[2][1]);
				drot[1] = 2 * y * (Gm[1][0] + Gm[0][1]) + 2 * zq * (Gm[2][0] + Gm[0][2]) + 2 * r * (Gm[1][2] - Gm[2][1]) - 4 * x * (Gm[2][2] + Gm[1][1]);
				drot[2] = 2 * x * (Gm[1][0] + Gm[0][1]) + 2 * r * (Gm[2][0] - Gm[0][2]) + 2 * zq * (Gm[1][2] + Gm[2][1]) - 4 * y * (Gm[2][2] + Gm[0][0]);
				drot[3] = 2 * r * (Gm[0][1] - Gm[1][0]) + 2 * x * (Gm[2][0] + Gm[0][2]) + 2 * y * (Gm[1][2] + Gm[2][1]) - 4 * zq * (Gm[1][1] + Gm[0][0]);
			}
		} else if (s.shs && !s.accumulate_grads) {
			if (!sh_rows) {
				for (int i = 0; i < s.M * 3; i++) dL_dsh[(size_t)idx * s.M * 3 + i] = 0.f;
			} else if (dL_dcolors) {
				dL_dcolors[3 * (size_t)idx] = 0.f; dL_dcolors[3 * (size_t)idx + 1] = 0.f; dL_dcolors[3 * (size_t)idx + 2] = 0.f;
			}
		}
		if (visible) {      // densification statistics (gaussian_model.py:767-771, slam_backend.py:115-121)
			if (s.densify_grad_accum) s.densify_grad_accum[idx] += sqrtf(dm2x * dm2x + dm2y * dm2y);
			if (s.densify_denom) s.densify_denom[idx] += 1.f;
			if (s.max_radii2D) s.max_radii2D[idx] = fmaxf(s.max_radii2D[idx], (float)radii[idx]);
		}
		if (!s.accumulate_grads) {
			dL_dopacity[idx] = dopac;
			if (s.scales) reinterpret_cast<float4*>(dL_drot)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
			if (dL_dcov3D_out) {
#pragma unroll
				for (int i = 0; i < 6; i++) dL_dcov3D_out[(size_t)idx * 6 + i] = dcov[i];
			}
		} else if (visible && s.accumulate_grads == 2) {
			// window accumulation by views running concurrently on several streams: REDs
			atomicAdd(&dL_dopacity[idx], dopac);
			if (s.scales) red_add_v4(reinterpret_cast<float4*>(dL_drot) + idx, make_float4(drot[0], drot[1], drot[2], drot[3]));
			if (dL_dcov3D_out) {
#pragma unroll
				for (int i = 0; i < 6; i++) atomicAdd(&dL_dcov3D_out[(size_t)idx * 6 + i], dcov[i]);
			}
		} else if (visible) {
			// window accumulation: this thread owns row idx, plain read-modify-write (views run in stream order)
			dL_dopacity[idx] += dopac;
			if (s.scales) {
				float4 rr = reinterpret_cast<float4*>(dL_drot)[idx];
				rr.x += drot[0]; rr.y += drot[1]; rr.z += drot[2]; rr.w += drot[3];
				reinterpret_cast<float4*>(dL_drot)[idx] = rr;
			}
			if (dL_dcov3D_out) {
#pragma unroll
				for (int i = 0; i < 6; i++) dL_dcov3D_out[(size_t)idx * 6 + i] += dcov[i];
			}
		}
	}

	// ---- the [P,3] outputs: rows into shared memory, out as 16-byte stores (rows of culled Gaussians are zero: they are
	// written when every row is, and adding them in a window accumulation changes nothing) ----
	__syncwarp();      // every lane has taken its inputs out of rows0, rows1
	{
		const int t3 = 3 * lane;
		rows0[t3] = dmean[0]; rows0[t3 + 1] = dmean[1]; rows0[t3 + 2] = dmean[2];
		rows1[t3] = dscale[0]; rows1[t3 + 1] = dscale[1]; rows1[t3 + 2] = dscale[2];
		rows2[t3] = dm2x; rows2[t3 + 1] = dm2y; rows2[t3 + 2] = 0.f;
		rows3[t3] = dcol[0]; rows3[t3 + 1] = dcol[1]; rows3[t3 + 2] = dcol[2];
	}
	__syncwarp();
	{
		const int accum = s.accumulate_grads;
		store_rows3<32>(dL_dmeans2D, row0, s.P, rows2, vec_mask & 4, 0);
		store_rows3<32>(dL_dmeans3D, row0, s.P, rows0, vec_mask & 8, accum);
		if (s.scales) store_rows3<32>(dL_dscales, row0, s.P, rows1, vec_mask & 16, accum);
		if (sh_rows) store_rows3<32>(dL_dsh, row0, s.P, rows3, vec_mask & 32, accum);
		else if (dL_dcolors) store_rows3<32>(dL_dcolors, row0, s.P, rows3, vec_mask & 64, accum);
	}

	GSR_PROBE(2, 1);
	if (kEarlyMath) pdl_wait();      // lanes without a visible Gaussian have not waited yet: no CTA ends before its predecessor
	// ---- block-level 6-vector reduction of the pose gradient ----
#pragma unroll
	for (int i = 0; i < 6; i++) tau[i] = warp_sum(tau[i]);
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int i = 0; i < 6; i++) s_tau[threadIdx.x >> 5][i] = tau[i];
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		float v = 0.f;
#pragma unroll
		for (int w = 0; w < 8; w++) v += s_tau[w][threadIdx.x];
		g.tau_partial[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
	}
	// Release by ONE thread behind the CTA barrier (the pattern of cooperative-groups' grid sync): the barrier orders the six
	// partial stores before thread 0's fence, and the fence is cumulative.  A fence by every thread would also wait for
	// each thread's ~20 output stores to be acknowledged -- 30 % of this kernel's stall samples at P = 500 k.
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		s_last = (atomicAdd(&g.hdr->bwd_blocks_done, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	GSR_PROBE(2, 2);
	if (s_last) {
		__threadfence();
		// deterministic final sum: 6 warps, one component each, fixed strided order + shuffle tree
		const int comp = threadIdx.x >> 5, lane = threadIdx.x & 31;
		if (comp < 6) {
			float v = 0.f;
			for (unsigned b0 = lane; b0 < gridDim.x; b0 += 32 * 8) {     // eight independent loads in flight, fixed order
				float p[8];
#pragma unroll
				for (int u = 0; u < 8; u++) {
					const unsigned b = b0 + 32 * u;
					p[u] = b < gridDim.x ? __ldcg(&g.tau_partial[(size_t)b * 8 + comp]) : 0.f;
				}
#pragma unroll
				for (int u = 0; u < 8; u++) v += p[u];
			}
			v = warp_sum(v);
			if (lane == 0) dL_dtau[comp] = v;
		}
		if (threadIdx.x == 0) g.hdr->bwd_blocks_done = 0;
	}
	GSR_PROBE(2, 3);
}

}  // namespace

void launch_preprocess_backward(const Scene& s, const GeomView& g, const int* radii, float* dL_dmeans3D,
                                float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors, float* dL_dopacity,
                                float* dL_dscales, float* dL_drotations, float* dL_dcov3D, float* dL_dtau,
                                bool behind_render_backward, cudaStream_t stream)
{
	if (s.P == 0) {
		cudaMemsetAsync(dL_dtau, 0, 6 * sizeof(float), stream);
		return;
	}
	// behind_render_backward: programmatic dependent of the compositing backward launched just before on this stream -- the
	// CTAs move in during that kernel's tail, fetch their inputs and wait (griddepcontrol.wait) for its accumulators
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((s.P + 255) / 256); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = behind_render_backward ? 1 : 0;
	auto a16 = [](const void* p) { return ((size_t)p & 15) == 0; };
	const int vec_mask = (a16(s.means3D) ? 1 : 0) | (a16(s.scales) ? 2 : 0) | (a16(dL_dmeans2D) ? 4 : 0) | (a16(dL_dmeans3D) ? 8 : 0) |
	                     (a16(dL_dscales) ? 16 : 0) | (a16(dL_dsh) ? 32 : 0) | (a16(dL_dcolors) ? 64 : 0);
	// CTAs per SM the kernel is compiled for: one wave of latency-bound CTAs runs best without spills at 3 x 8 warps (80
	// registers); several waves of them gain more from a fourth CTA per SM (64 registers, 48 bytes spilled) -- 3 M Gaussians
	// 0.54 -> 0.47 ms, 500 k 0.086 -> 0.079 ms, 100 k 0.0166 -> 0.0176 ms.  GSR_PB_MINB overrides (A/B switch).
	static const int minb_env = getenv("GSR_PB_MINB") ? atoi(getenv("GSR_PB_MINB")) : 0;
	const int minb = minb_env ? minb_env : (s.P > 3 * 148 * 256 ? 4 : 3);
#define GSR_PB_LAUNCH(B)                                                                                                          \
	cudaLaunchKernelEx(&cfg, preprocess_backward_kernel<B>, s, g, radii, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacity, \
	                   dL_dscales, dL_drotations, dL_dcov3D, dL_dtau, vec_mask)
	if (minb == 2) GSR_PB_LAUNCH(2);
	else if (minb == 4) GSR_PB_LAUNCH(4);
	else if (minb == 5) GSR_PB_LAUNCH(5);
	else GSR_PB_LAUNCH(3);
#undef GSR_PB_LAUNCH
}

GSR_PROBE_READER(probe_read_preprocess_backward)

}  // namespace gsr
