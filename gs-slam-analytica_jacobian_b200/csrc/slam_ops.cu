// The callers on either side of the rasterizer in the SLAM tracking / mapping iteration,
// fused into two kernels so that a whole iteration (render -> loss -> backward -> optimiser -> pose update) can live
// in one CUDA graph without a single host round trip:
//
//   slam_loss_kernel     photometric + depth L1 loss of the reference (utils/slam_utils.py:56-128: tracking and
//                        mapping, monocular and RGB-D, exposure a/b, rgb boundary mask, gradient mask, opacity
//                        weighting) AND its gradients w.r.t. the rendered colour / depth images and the exposure
//                        parameters in one pass over the pixels: reads 9-10 floats, writes 4 per pixel, where the
//                        reference runs ~15 elementwise torch kernels forward + their autograd backward.
//   tracking_step_kernel torch.optim.Adam on [cam_rot_delta, cam_trans_delta, exposure_a, exposure_b]
//                        (utils/slam_frontend.py:129-162) + update_pose (utils/pose_utils.py:76-93: SE3_exp(tau) @ T_w2c,
//                        convergence test) + the camera tensors of the next render (utils/camera_utils.py:96-109,
//                        graphics_utils.py:33-46): one thread, ~20 tiny torch launches in the reference.
#include "gsr_params.h"

namespace gsr {

namespace {

// ---- loss ---------------------------------------------------------------------------------------------------------
// sums[0] = loss, sums[1] = dL/dexposure_a, sums[2] = dL/dexposure_b; per-CTA partials -> last CTA sums in order
__global__ void __launch_bounds__(256)
slam_loss_kernel(SlamLossArgs a, float* __restrict__ partials, unsigned* __restrict__ ticket)
{
	__shared__ float s_red[8][3];
	__shared__ bool s_last;
	const int HW = a.W * a.H;
	const float ea = a.exposure ? __expf(a.exposure[0]) : 1.f;       // image_ab = exp(a) * image + b   (slam_utils.py:57,92)
	const float eb = a.exposure ? a.exposure[1] : 0.f;
	const float inv_rgb = 1.f / (3.f * (float)HW), inv_d = 1.f / (float)HW;
	const float w_rgb = a.use_depth ? a.alpha : 1.f, w_d = 1.f - a.alpha;
	float loss = 0.f, ga = 0.f, gb = 0.f;
	for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
		const float gt0 = a.gt_color[p], gt1 = a.gt_color[HW + p], gt2 = a.gt_color[2 * HW + p];
		// rgb_pixel_mask = (gt.sum(0) > threshold) [* grad_mask when tracking]   (slam_utils.py:68-69,103)
		float m = (gt0 + gt1 + gt2 > a.rgb_boundary_threshold) ? 1.f : 0.f;
		if (a.grad_mask) m *= a.grad_mask[p] ? 1.f : 0.f;
		const float op = a.opacity[p];
		const float wgt = a.opacity_weighted ? op : 1.f;               // tracking: l1 = opacity * |...|   (:70)
		const float c[3] = {a.color[p], a.color[HW + p], a.color[2 * HW + p]};
		const float gt[3] = {gt0, gt1, gt2};
#pragma unroll
		for (int ch = 0; ch < 3; ch++) {
			const float iab = ea * c[ch] + eb;
			const float d = iab * m - gt[ch] * m;
			loss += w_rgb * inv_rgb * wgt * fabsf(d);
			const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
			const float g_iab = w_rgb * inv_rgb * wgt * m * sgn;        // dL/dimage_ab
			a.dL_dcolor[ch * HW + p] = g_iab * ea;
			ga += g_iab * ea * c[ch];                                   // d image_ab / d a = exp(a) * image
			gb += g_iab;
		}
		float gd = 0.f;
		if (a.use_depth) {
			const float gtd = a.gt_depth[p];
			float dm = (gtd > 0.01f) ? 1.f : 0.f;                       // depth_pixel_mask   (:83,108)
			if (a.opacity_weighted) dm *= (op > 0.95f) ? 1.f : 0.f;     // tracking only: opacity_mask   (:84)
			const float dd = a.depth[p] * dm - gtd * dm;
			loss += w_d * inv_d * fabsf(dd);
			gd = w_d * inv_d * dm * ((dd > 0.f) ? 1.f : ((dd < 0.f) ? -1.f : 0.f));
		}
		a.dL_ddepth[p] = gd;
	}
	loss = warp_sum(loss); ga = warp_sum(ga); gb = warp_sum(gb);
	if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = loss; s_red[threadIdx.x >> 5][1] = ga; s_red[threadIdx.x >> 5][2] = gb; }
	__syncthreads();
	if (threadIdx.x < 3) {
		float v = 0.f;
#pragma unroll
		for (int w = 0; w < 8; w++) v += s_red[w][threadIdx.x];
		partials[blockIdx.x * 4 + threadIdx.x] = v;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();      // one cumulative release behind the CTA barrier
		s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	if (threadIdx.x < 96) {      // deterministic: warp k sums component k over the CTAs in a fixed order
		const int comp = threadIdx.x >> 5, lane = threadIdx.x & 31;
		float v = 0.f;
		for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(&partials[b * 4 + comp]);
		v = warp_sum(v);
		if (lane == 0) a.sums[comp] = v;
	}
	if (threadIdx.x == 0) *ticket = 0;
}

// ---- optimiser + pose update ----------------------------------------------------------------------------------------
__device__ void mat3_mul(const float* A, const float* B, float* C)      // row-major 3x3
{
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__global__ void tracking_step_kernel(TrackingStepArgs a)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	// The reference leaves its loop on the iteration that converged (`if converged: break`, slam_frontend.py:180,192-193): replays
	// of a captured iteration behind it must leave pose, exposure, Adam state and the iteration counter exactly as they are
	// (the camera block already holds the converged pose).
	if (a.status[2] != 0) return;
	// ---- Adam (torch defaults: betas 0.9 / 0.999, eps 1e-8), parameters start from 0 for the pose deltas ----
	// parameter order: rot delta (3), trans delta (3), exposure_a, exposure_b   (slam_frontend.py:132-160)
	float grad[8];
	for (int i = 0; i < 3; i++) { grad[i] = a.dL_dtau[3 + i]; grad[3 + i] = a.dL_dtau[i]; }   // theta = tau[3:], rho = tau[:3]
	grad[6] = a.dL_dexposure ? a.dL_dexposure[1] : 0.f;      // sums[1] = dL/da (caller passes the loss kernel's `sums`)
	grad[7] = a.dL_dexposure ? a.dL_dexposure[2] : 0.f;
	float delta[8];
	const float step = a.adam_state[16] + 1.f;
	a.adam_state[16] = step;
	const float bc1 = 1.f - powf(0.9f, step), bc2 = 1.f - powf(0.999f, step);
	for (int i = 0; i < 8; i++) {
		const float lr = i < 3 ? a.lr_rot : (i < 6 ? a.lr_trans : a.lr_exposure);
		float m = a.adam_state[i], v = a.adam_state[8 + i];
		m = 0.9f * m + 0.1f * grad[i];                       // exp_avg.lerp_(grad, 1 - beta1)
		v = 0.999f * v + 0.001f * grad[i] * grad[i];
		a.adam_state[i] = m; a.adam_state[8 + i] = v;
		const float denom = sqrtf(v) / sqrtf(bc2) + 1e-8f;
		delta[i] = -(lr / bc1) * (m / denom);
	}
	if (a.exposure) { a.exposure[0] += delta[6]; a.exposure[1] += delta[7]; }
	// ---- update_pose: new_w2c = SE3_exp([trans delta, rot delta]) @ [R | T]   (pose_utils.py:61-93) ----
	const float rho[3] = {delta[3], delta[4], delta[5]}, th[3] = {delta[0], delta[1], delta[2]};
	const float Wm[9] = {0.f, -th[2], th[1], th[2], 0.f, -th[0], -th[1], th[0], 0.f};
	float W2[9];
	mat3_mul(Wm, Wm, W2);
	const float angle = sqrtf(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
	float cA, cB, vB, vC;      // R = I + cA W + cB W2 ; V = I + vB W + vC W2
	if (angle < 1e-5f) { cA = 1.f; cB = 0.5f; vB = 0.5f; vC = 1.f / 6.f; }
	else {
		const float s = sinf(angle), c = cosf(angle);
		cA = s / angle; cB = (1.f - c) / (angle * angle);
		vB = cB; vC = (angle - s) / (angle * angle * angle);
	}
	float Rd[9], Vm[9];
	for (int i = 0; i < 9; i++) {
		const float I = (i % 4 == 0) ? 1.f : 0.f;
		Rd[i] = I + cA * Wm[i] + cB * W2[i];
		Vm[i] = I + vB * Wm[i] + vC * W2[i];
	}
	float td[3];
	for (int i = 0; i < 3; i++) td[i] = Vm[3 * i] * rho[0] + Vm[3 * i + 1] * rho[1] + Vm[3 * i + 2] * rho[2];
	float R[9], T[3], Rn[9], Tn[3];
	for (int i = 0; i < 9; i++) R[i] = a.RT[i];
	for (int i = 0; i < 3; i++) T[i] = a.RT[9 + i];
	mat3_mul(Rd, R, Rn);
	for (int i = 0; i < 3; i++) Tn[i] = Rd[3 * i] * T[0] + Rd[3 * i + 1] * T[1] + Rd[3 * i + 2] * T[2] + td[i];
	for (int i = 0; i < 9; i++) a.RT[i] = Rn[i];
	for (int i = 0; i < 3; i++) a.RT[9 + i] = Tn[i];
	float n2 = 0.f;
	for (int i = 0; i < 6; i++) n2 += delta[i] * delta[i];
	const int conv = sqrtf(n2) < a.converged_threshold ? 1 : 0;
	a.status[0] = conv;
	a.status[1] = a.status[1] + 1;                            // iterations done
	if (conv && a.status[2] == 0) a.status[2] = a.status[1];   // first converged iteration (sticky)
	// ---- camera tensors of the next render (camera_utils.py:96-109): packed block view | proj | proj_raw | campos ----
	float* view = a.camera_block;                             // world_view_transform = [R|T; 0 0 0 1]^T, row-major
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) view[4 * i + j] = Rn[3 * j + i];
		view[4 * i + 3] = 0.f;
		view[12 + i] = Tn[i];
	}
	view[15] = 1.f;
	float* proj = a.camera_block + 16;                        // full_proj_transform = world_view_transform @ projection_matrix
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) {
			float acc = 0.f;
			for (int k = 0; k < 4; k++) acc += view[4 * i + k] * a.proj_raw[4 * k + j];
			proj[4 * i + j] = acc;
		}
	for (int i = 0; i < 16; i++) a.camera_block[32 + i] = a.proj_raw[i];
	// camera_center = inverse(world_view_transform)[3, :3] = -R^{-1} T (true inverse of R, like torch's)
	const float det = Rn[0] * (Rn[4] * Rn[8] - Rn[5] * Rn[7]) - Rn[1] * (Rn[3] * Rn[8] - Rn[5] * Rn[6]) + Rn[2] * (Rn[3] * Rn[7] - Rn[4] * Rn[6]);
	const float id = 1.f / det;
	const float Ri[9] = {(Rn[4] * Rn[8] - Rn[5] * Rn[7]) * id, (Rn[2] * Rn[7] - Rn[1] * Rn[8]) * id, (Rn[1] * Rn[5] - Rn[2] * Rn[4]) * id,
	                     (Rn[5] * Rn[6] - Rn[3] * Rn[8]) * id, (Rn[0] * Rn[8] - Rn[2] * Rn[6]) * id, (Rn[2] * Rn[3] - Rn[0] * Rn[5]) * id,
	                     (Rn[3] * Rn[7] - Rn[4] * Rn[6]) * id, (Rn[1] * Rn[6] - Rn[0] * Rn[7]) * id, (Rn[0] * Rn[4] - Rn[1] * Rn[3]) * id};
	for (int i = 0; i < 3; i++) a.camera_block[48 + i] = -(Ri[3 * i] * Tn[0] + Ri[3 * i + 1] * Tn[1] + Ri[3 * i + 2] * Tn[2]);
	a.camera_block[51] = 0.f;
}

}  // namespace

size_t slam_loss_scratch_bytes(int W, int H)
{
	(void)W; (void)H;
	return (size_t)(4 * 1024 + 4) * sizeof(float);      // <= 1024 CTAs x 4 partials + ticket
}

void launch_slam_loss(const SlamLossArgs& a, void* scratch, cudaStream_t stream)
{
	const int HW = a.W * a.H;
	int grid = (HW + 255) / 256;
	if (grid > 1024) grid = 1024;
	if (grid < 1) grid = 1;
	float* partials = (float*)scratch;
	unsigned* ticket = (unsigned*)(partials + 4 * 1024);
	slam_loss_kernel<<<grid, 256, 0, stream>>>(a, partials, ticket);
}

void launch_tracking_step(const TrackingStepArgs& a, cudaStream_t stream) { tracking_step_kernel<<<1, 32, 0, stream>>>(a); }

}  // namespace gsr
