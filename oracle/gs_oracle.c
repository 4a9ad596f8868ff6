/*
 * TEST INFRASTRUCTURE (oracle) -- NOT product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * CPU restatement of the reference differentiable Gaussian rasterizer
 * (notu97/GS-SLAM-Analytica_Jacobian, submodules/diff-gaussian-rasterization/cuda_rasterizer).
 * Paths cited below are relative to that directory ("CR/").  The file is compiled twice,
 * with GSO_REAL=float (suffix _f32: same working precision as the CUDA kernels) and
 * GSO_REAL=double (suffix _f64: the accuracy reference), into oracle/libgs_oracle.so.
 *
 * Parity pinning: (i) the single-Gaussian known answers of the reference's
 * 3DGS_Analytical_Jacobian.ipynb (cells 1,7,8: d mu_I/d tau, d Sigma_I/d tau) via
 * gso_cov2d_pose_jacobian / gso_mean2d_pose_jacobian (tests/test_oracle_kat.py),
 * (ii) torch-autograd through SE3_exp(tau)*T_cw (utils/pose_utils.py:61-93), and
 * (iii) on the GPU box the unmodified reference kernels themselves (oracle/_ref/libgsref.so,
 * tests/test_parity_gpu.py, tests/test_config_scale_gpu.py at the BASELINE shapes) plus committed fixtures generated
 * from them (tests/golden/make_ref_golden.py -> tests/test_golden_reference.py, on the CPU), and (iv) the reference's
 * analytic-Jacobian script chain executed on C0 (tests/golden/make_script_chain_golden.py -> tests/test_script_chain.py;
 * gso_set_math_mode selects the script's dense formulation).
 *
 * Stages (each takes/returns plain arrays so a test can substitute the CUDA intermediates):
 *   gso_preprocess   CR/forward.cu:157-401  (+ auxiliary.h:41-56,139-164)
 *   gso_bin          CR/rasterizer_impl.cu:70-138, 339-368 (duplicateWithKeys, stable sort, ranges)
 *   gso_render       CR/forward.cu:406-535
 *   gso_render_bwd   CR/backward.cu:648-872
 *   gso_preprocess_bwd  CR/backward.cu:150-345 (computeCov2DCUDA), :494-624 (preprocessCUDA),
 *                       :21-145 (SH), :426-489 (cov3D)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef GSO_REAL
#define GSO_REAL double
#define GSO_SUFFIX _f64
#endif
#define GSO_CAT2(a, b) a##b
#define GSO_CAT(a, b) GSO_CAT2(a, b)
#define FN(name) GSO_CAT(name, GSO_SUFFIX)
typedef GSO_REAL real;

#define BLK 16

static const real SH_C0 = (real)0.28209479177387814;
static const real SH_C1 = (real)0.4886025119029199;
static const real SH_C2[5] = {(real)1.0925484305920792, (real)-1.0925484305920792, (real)0.31539156525252005,
                              (real)-1.0925484305920792, (real)0.5462742152960396};
static const real SH_C3[7] = {(real)-0.5900435899266435, (real)2.890611442640554, (real)-0.4570457994644658,
                              (real)0.3731763325901154, (real)-0.4570457994644658, (real)1.445305721320277,
                              (real)-0.5900435899266435};

/* float->int conversion with the saturating / NaN->0 behaviour of CUDA's cvt.rzi.s32.f32 */
static int f2i_sat(real v)
{
	if (v != v) return 0;
	if (v >= (real)2147483647.0) return 2147483647;
	if (v <= (real)-2147483648.0) return (-2147483647 - 1);
	return (int)v;
}
static real rmin(real a, real b) { return a < b ? a : b; }
static real rmax(real a, real b) { return a > b ? a : b; }

/* CR/auxiliary.h:46-56 getRect */
static void get_rect(real px, real py, int max_radius, int gx, int gy, int* rmin_, int* rmax_)
{
	int v;
	v = f2i_sat((px - (real)max_radius) / (real)BLK); if (v < 0) v = 0; if (v > gx) v = gx; rmin_[0] = v;
	v = f2i_sat((py - (real)max_radius) / (real)BLK); if (v < 0) v = 0; if (v > gy) v = gy; rmin_[1] = v;
	v = f2i_sat((px + (real)max_radius + (real)(BLK - 1)) / (real)BLK); if (v < 0) v = 0; if (v > gx) v = gx; rmax_[0] = v;
	v = f2i_sat((py + (real)max_radius + (real)(BLK - 1)) / (real)BLK); if (v < 0) v = 0; if (v > gy) v = gy; rmax_[1] = v;
}

/* CR/forward.cu:120-154 computeCov3D (quaternion NOT normalised, :129) */
static void cov3d_from_scale_rot(const float* scale, real mod, const float* rot, real* cov3D)
{
	real s[3] = {mod * (real)scale[0], mod * (real)scale[1], mod * (real)scale[2]};
	real r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	real Rq[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)},
	                 {2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)},
	                 {2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)}};
	/* M = S * Rq^T ; Sigma = M^T M = Rq S^2 Rq^T */
	real Sg[3][3];
	for (int a = 0; a < 3; a++)
		for (int b = 0; b < 3; b++) {
			real acc = 0;
			for (int i = 0; i < 3; i++) acc += (s[i] * Rq[a][i]) * (s[i] * Rq[b][i]);
			Sg[a][b] = acc;
		}
	cov3D[0] = Sg[0][0]; cov3D[1] = Sg[0][1]; cov3D[2] = Sg[0][2];
	cov3D[3] = Sg[1][1]; cov3D[4] = Sg[1][2]; cov3D[5] = Sg[2][2];
}

/* view matrix is column-major: element (r,c) = v[c*4+r] (CR/auxiliary.h:58-66) */
#define VR(v, r, c) ((real)(v)[(c) * 4 + (r)])

typedef struct {
	real t[3];        /* clamped camera-space mean */
	real xmul, ymul;  /* 0 where the 1.3*tanfov clamp is active (CR/backward.cu:182-183) */
	real A[2][3];     /* first two rows of J*R_cw  (glm T[0][k], T[1][k]) */
	real J00, J11, J02, J12;
	real a, b, c;     /* cov2D incl. +0.3 low-pass */
} cov2d_fw;

/* CR/forward.cu:76-115 computeCov2D */
static void cov2d_forward(const real* mean, real fx, real fy, real tanx, real tany, const real* cov3D,
                          const float* view, cov2d_fw* o)
{
	real t[3];
	for (int r = 0; r < 3; r++)
		t[r] = VR(view, r, 0) * mean[0] + VR(view, r, 1) * mean[1] + VR(view, r, 2) * mean[2] + VR(view, r, 3);
	real limx = (real)1.3f * tanx, limy = (real)1.3f * tany;
	real txtz = t[0] / t[2], tytz = t[1] / t[2];
	o->xmul = (txtz < -limx || txtz > limx) ? 0 : 1;
	o->ymul = (tytz < -limy || tytz > limy) ? 0 : 1;
	t[0] = rmin(limx, rmax(-limx, txtz)) * t[2];
	t[1] = rmin(limy, rmax(-limy, tytz)) * t[2];
	o->t[0] = t[0]; o->t[1] = t[1]; o->t[2] = t[2];
	o->J00 = fx / t[2]; o->J11 = fy / t[2];
	o->J02 = -(fx * t[0]) / (t[2] * t[2]);
	o->J12 = -(fy * t[1]) / (t[2] * t[2]);
	for (int k = 0; k < 3; k++) {
		o->A[0][k] = o->J00 * VR(view, 0, k) + o->J02 * VR(view, 2, k);
		o->A[1][k] = o->J11 * VR(view, 1, k) + o->J12 * VR(view, 2, k);
	}
	real V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
	real VA0[3], VA1[3];
	for (int i = 0; i < 3; i++) {
		VA0[i] = V[i][0] * o->A[0][0] + V[i][1] * o->A[0][1] + V[i][2] * o->A[0][2];
		VA1[i] = V[i][0] * o->A[1][0] + V[i][1] * o->A[1][1] + V[i][2] * o->A[1][2];
	}
	o->a = o->A[0][0] * VA0[0] + o->A[0][1] * VA0[1] + o->A[0][2] * VA0[2] + (real)0.3f;
	o->b = o->A[0][0] * VA1[0] + o->A[0][1] * VA1[1] + o->A[0][2] * VA1[2];
	o->c = o->A[1][0] * VA1[0] + o->A[1][1] * VA1[1] + o->A[1][2] * VA1[2] + (real)0.3f;
}

/* CR/forward.cu:22-73 computeColorFromSH */
static void sh_forward(int deg, int M, const float* mean, const float* campos, const float* sh, real* rgb,
                       uint8_t* clamped)
{
	real dir[3] = {(real)mean[0] - (real)campos[0], (real)mean[1] - (real)campos[1], (real)mean[2] - (real)campos[2]};
	real len = (real)sqrt((double)(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]));
	real x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
	(void)M;
	for (int ch = 0; ch < 3; ch++) {
#define SH(i) ((real)sh[(i) * 3 + ch])
		real res = SH_C0 * SH(0);
		if (deg > 0) {
			res = res - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
			if (deg > 1) {
				real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				res = res + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * (2 * zz - xx - yy) * SH(6) +
				      SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
				if (deg > 2) {
					res = res + SH_C3[0] * y * (3 * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
					      SH_C3[2] * y * (4 * zz - xx - yy) * SH(11) +
					      SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * SH(12) +
					      SH_C3[4] * x * (4 * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
					      SH_C3[6] * x * (xx - 3 * yy) * SH(15);
				}
			}
		}
#undef SH
		res += (real)0.5;
		clamped[ch] = (res < 0);
		rgb[ch] = res < 0 ? 0 : res;
	}
}

/* ------------------------------------------------------------------------------------------
 * Stage 1: per-Gaussian preprocess.  CR/forward.cu:157-401.
 * Outputs for culled Gaussians (radii==0) are left as passed in (reference: stale, B6).
 * Returns sum(tiles_touched) (= num_rendered, CR/rasterizer_impl.cu:330-331).
 * ------------------------------------------------------------------------------------------ */
long long FN(gso_preprocess)(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                             const float* rotations, const float* opacities, const float* shs,
                             const float* cov3D_precomp, const float* colors_precomp, const float* view,
                             const float* proj, const float* campos, int W, int H, float tan_fovx, float tan_fovy,
                             int* radii, real* means2D, real* depths, real* cov3Ds, real* rgb, real* conic_opacity,
                             uint8_t* clamped, uint32_t* tiles_touched)
{
	const real fy = (real)H / ((real)2 * (real)tan_fovy); /* CR/rasterizer_impl.cu:272-273 */
	const real fx = (real)W / ((real)2 * (real)tan_fovx);
	const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
	long long total = 0;
	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0;
		tiles_touched[idx] = 0;
		real p[3] = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
		real hom[4], pv[3];
		for (int r = 0; r < 4; r++)
			hom[r] = (real)proj[r] * p[0] + (real)proj[4 + r] * p[1] + (real)proj[8 + r] * p[2] + (real)proj[12 + r];
		for (int r = 0; r < 3; r++)
			pv[r] = VR(view, r, 0) * p[0] + VR(view, r, 1) * p[1] + VR(view, r, 2) * p[2] + VR(view, r, 3);
		if (pv[2] <= (real)0.2f) continue; /* CR/auxiliary.h:154 */
		real pw = (real)1 / (hom[3] + (real)0.0000001f);
		real pproj[2] = {hom[0] * pw, hom[1] * pw};
		real c3[6];
		if (cov3D_precomp) {
			for (int i = 0; i < 6; i++) c3[i] = cov3D_precomp[6 * idx + i];
		} else {
			cov3d_from_scale_rot(scales + 3 * idx, (real)scale_modifier, rotations + 4 * idx, c3);
			for (int i = 0; i < 6; i++) cov3Ds[6 * idx + i] = c3[i];
		}
		cov2d_fw cf;
		cov2d_forward(p, fx, fy, (real)tan_fovx, (real)tan_fovy, c3, view, &cf);
		real det = cf.a * cf.c - cf.b * cf.b;
		if (det == 0) continue;
		real det_inv = (real)1 / det;
		real conic[3] = {cf.c * det_inv, -cf.b * det_inv, cf.a * det_inv};
		real mid = (real)0.5 * (cf.a + cf.c);
		real sq = (real)sqrt((double)rmax((real)0.1f, mid * mid - det));
		real l1 = mid + sq, l2 = mid - sq;
		real my_radius = (real)ceil((double)((real)3 * (real)sqrt((double)rmax(l1, l2))));
		/* ndc2Pix evaluated in double then rounded once (CR/auxiliary.h:41-44) */
		real pix[2] = {(real)((((double)pproj[0] + 1.0) * W - 1.0) * 0.5), (real)((((double)pproj[1] + 1.0) * H - 1.0) * 0.5)};
		int irad = f2i_sat(my_radius);
		int rmn[2], rmx[2];
		get_rect(pix[0], pix[1], irad, gx, gy, rmn, rmx);
		if ((rmx[0] - rmn[0]) * (rmx[1] - rmn[1]) == 0) continue;
		if (!colors_precomp) sh_forward(D, M, means3D + 3 * idx, campos, shs + (size_t)idx * M * 3, rgb + 3 * idx, clamped + 3 * idx);
		depths[idx] = pv[2];
		radii[idx] = irad;
		means2D[2 * idx] = pix[0]; means2D[2 * idx + 1] = pix[1];
		conic_opacity[4 * idx] = conic[0]; conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2]; conic_opacity[4 * idx + 3] = opacities[idx];
		tiles_touched[idx] = (uint32_t)((rmx[1] - rmn[1]) * (rmx[0] - rmn[0]));
		total += tiles_touched[idx];
	}
	return total;
}

/* CR/rasterizer_impl.cu:54-66 checkFrustum / markVisible */
void FN(gso_mark_visible)(int P, const float* means3D, const float* view, uint8_t* present)
{
	for (int idx = 0; idx < P; idx++) {
		real z = VR(view, 2, 0) * (real)means3D[3 * idx] + VR(view, 2, 1) * (real)means3D[3 * idx + 1] +
		         VR(view, 2, 2) * (real)means3D[3 * idx + 2] + VR(view, 2, 3);
		present[idx] = !(z <= (real)0.2f);
	}
}

/* ------------------------------------------------------------------------------------------
 * Stage 2: binning.  duplicateWithKeys (CR/rasterizer_impl.cu:70-111), stable sort of
 * (tile<<32 | float_bits(depth)) keys (:353-358), identifyTileRanges (:116-138, ranges zeroed :360).
 * point_list[R], ranges[2*tiles]. keys_out (optional) receives the sorted 64-bit keys.
 * ------------------------------------------------------------------------------------------ */
int FN(gso_bin)(int P, int W, int H, const int* radii, const real* means2D, const real* depths, long long R,
                uint32_t* point_list, uint32_t* ranges, uint64_t* keys_out)
{
	const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
	memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
	if (R <= 0) return 0;
	uint64_t* k0 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
	uint64_t* k1 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
	uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
	uint32_t* v1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
	if (!k0 || !k1 || !v0 || !v1) { free(k0); free(k1); free(v0); free(v1); return -1; }
	long long off = 0;
	for (int idx = 0; idx < P; idx++) {
		if (radii[idx] <= 0) continue;
		int rmn[2], rmx[2];
		get_rect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, rmn, rmx);
		float df = (float)depths[idx];
		uint32_t dbits;
		memcpy(&dbits, &df, 4);
		for (int y = rmn[1]; y < rmx[1]; y++)
			for (int x = rmn[0]; x < rmx[0]; x++) {
				if (off >= R) { free(k0); free(k1); free(v0); free(v1); return -2; }
				k0[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
				v0[off] = (uint32_t)idx;
				off++;
			}
	}
	if (off != R) { free(k0); free(k1); free(v0); free(v1); return -3; }
	/* stable LSD radix sort, 16-bit digits over the low 48 bits (tile ids < 2^16) */
	size_t* hist = (size_t*)malloc(sizeof(size_t) * 65536);
	for (int pass = 0; pass < 3; pass++) {
		int sh = 16 * pass;
		memset(hist, 0, sizeof(size_t) * 65536);
		for (long long i = 0; i < R; i++) hist[(k0[i] >> sh) & 0xFFFF]++;
		size_t acc = 0;
		for (int d = 0; d < 65536; d++) { size_t c = hist[d]; hist[d] = acc; acc += c; }
		for (long long i = 0; i < R; i++) {
			size_t dst = hist[(k0[i] >> sh) & 0xFFFF]++;
			k1[dst] = k0[i]; v1[dst] = v0[i];
		}
		uint64_t* tk = k0; k0 = k1; k1 = tk;
		uint32_t* tv = v0; v0 = v1; v1 = tv;
	}
	free(hist);
	for (long long i = 0; i < R; i++) {
		point_list[i] = v0[i];
		if (keys_out) keys_out[i] = k0[i];
		uint32_t cur = (uint32_t)(k0[i] >> 32);
		if (i == 0) ranges[2 * cur] = 0;
		else {
			uint32_t prev = (uint32_t)(k0[i - 1] >> 32);
			if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
		}
		if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
	}
	free(k0); free(k1); free(v0); free(v1);
	return 0;
}

/* "Math mode" (tests only): compositing without the rasterizer's cut-offs -- alpha = min(1, o G), no power > 0 skip, no
 * alpha < 1/255 skip, no T < 1e-4 stop -- i.e. the dense formulation of the reference's analytic-Jacobian scripts
 * (Loss_Derivative_script_compare.py:1173-1351), so that their chain can be compared term by term. */
static int g_math_mode = 0; /* one per working precision (this file is compiled twice) */
void FN(gso_set_math_mode)(int on) { g_math_mode = on; }

/* ------------------------------------------------------------------------------------------
 * Stage 3: forward compositing.  CR/forward.cu:406-535.  features = colors_precomp or rgb.
 * ------------------------------------------------------------------------------------------ */
void FN(gso_render)(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const real* means2D,
                    const real* features, const real* conic_opacity, const real* depths, const float* bg,
                    real* out_color, real* out_depth, real* out_opacity, real* final_T, uint32_t* n_contrib,
                    int* n_touched)
{
	const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < gx * gy; tile++) {
		int tx = tile % gx, ty = tile / gx;
		uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		for (int py = ty * BLK; py < ty * BLK + BLK && py < H; py++)
			for (int px = tx * BLK; px < tx * BLK + BLK && px < W; px++) {
				real T = 1, C[3] = {0, 0, 0}, Dp = 0;
				uint32_t contributor = 0, last_contributor = 0;
				for (uint32_t i = r0; i < r1; i++) {
					contributor++;
					uint32_t id = point_list[i];
					real dx = means2D[2 * id] - (real)px, dy = means2D[2 * id + 1] - (real)py;
					const real* co = conic_opacity + 4 * id;
					real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
					if (power > 0 && !g_math_mode) continue;
					real alpha = rmin(g_math_mode ? (real)1 : (real)0.99f, co[3] * (real)exp((double)power));
					if (alpha < (real)(1.0f / 255.0f) && !g_math_mode) continue;
					real test_T = T * (1 - alpha);
					if (test_T < (real)0.0001f && !g_math_mode) break; /* done = true */
					for (int ch = 0; ch < 3; ch++) C[ch] += features[3 * id + ch] * alpha * T;
					Dp += depths[id] * alpha * T;
					if (test_T > (real)0.5f) {
#pragma omp atomic
						n_touched[id] += 1;
					}
					T = test_T;
					last_contributor = contributor;
				}
				size_t pix = (size_t)W * py + px;
				final_T[pix] = T;
				n_contrib[pix] = last_contributor;
				for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pix] = C[ch] + T * (real)bg[ch];
				out_depth[pix] = Dp;
				out_opacity[pix] = 1 - T;
			}
	}
}

/* ------------------------------------------------------------------------------------------
 * Stage 4: backward compositing.  CR/backward.cu:648-872.
 * Accumulates (+=) into dL_dmean2D[P*3] (x,y used; NDC units), dL_dconic[P*4] (0,1,3 used),
 * dL_dopacity[P], dL_dcolor[P*3], dL_ddepth[P]; caller zero-fills (rasterize_points.cu:175-185).
 * ------------------------------------------------------------------------------------------ */
void FN(gso_render_bwd)(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const real* means2D,
                        const real* colors, const real* conic_opacity, const real* depths, const float* bg,
                        const real* final_T, const uint32_t* n_contrib, const real* dL_dpix, const real* dL_dpix_depth,
                        real* dL_dmean2D, real* dL_dconic, real* dL_dopacity, real* dL_dcolor, real* dL_ddepth)
{
	const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
	const real ddelx_dx = (real)0.5 * W, ddely_dy = (real)0.5 * H;
#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < gx * gy; tile++) {
		int tx = tile % gx, ty = tile / gx;
		uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		for (int py = ty * BLK; py < ty * BLK + BLK && py < H; py++)
			for (int px = tx * BLK; px < tx * BLK + BLK && px < W; px++) {
				size_t pix = (size_t)W * py + px;
				const real T_final = final_T[pix];
				real T = T_final;
				const uint32_t last_contributor = n_contrib[pix];
				real accum_rec[3] = {0, 0, 0}, accum_rec_depth = 0;
				real dpx[3] = {dL_dpix[pix], dL_dpix[(size_t)H * W + pix], dL_dpix[(size_t)2 * H * W + pix]};
				real dpd = dL_dpix_depth[pix];
				real last_alpha = 0, last_color[3] = {0, 0, 0}, last_depth = 0;
				real bg_dot = (real)bg[0] * dpx[0] + (real)bg[1] * dpx[1] + (real)bg[2] * dpx[2];
				uint32_t n = r1 - r0;
				uint32_t start = last_contributor < n ? last_contributor : n;
				for (uint32_t j = start; j-- > 0;) { /* j = contributor index, only j < last_contributor */
					uint32_t id = point_list[r0 + j];
					real dx = means2D[2 * id] - (real)px, dy = means2D[2 * id + 1] - (real)py;
					const real* co = conic_opacity + 4 * id;
					real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
					if (power > 0 && !g_math_mode) continue;
					real G = (real)exp((double)power);
					real alpha = rmin(g_math_mode ? (real)1 : (real)0.99f, co[3] * G);
					if (alpha < (real)(1.0f / 255.0f) && !g_math_mode) continue;
					T = T / (1 - alpha);
					real dchannel_dcolor = alpha * T;
					real dL_dalpha = 0;
					real gc[3];
					for (int ch = 0; ch < 3; ch++) {
						real c = colors[3 * id + ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1 - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						dL_dalpha += (c - accum_rec[ch]) * dpx[ch];
						gc[ch] = dchannel_dcolor * dpx[ch];
					}
					real depth = depths[id];
					accum_rec_depth = last_alpha * last_depth + (1 - last_alpha) * accum_rec_depth;
					last_depth = depth;
					dL_dalpha += (depth - accum_rec_depth) * dpd;
					real gd = dchannel_dcolor * dpd;
					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += (-T_final / (1 - alpha)) * bg_dot;
					real dL_dG = co[3] * dL_dalpha;
					real gdx = G * dx, gdy = G * dy;
					real dG_ddelx = -gdx * co[0] - gdy * co[1];
					real dG_ddely = -gdy * co[2] - gdx * co[1];
					real v[10] = {dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy,
					              (real)-0.5 * gdx * dx * dL_dG, (real)-0.5 * gdx * dy * dL_dG, (real)-0.5 * gdy * dy * dL_dG,
					              G * dL_dalpha, gc[0], gc[1], gc[2], gd};
					real* dst[10] = {&dL_dmean2D[3 * id], &dL_dmean2D[3 * id + 1], &dL_dconic[4 * id], &dL_dconic[4 * id + 1],
					                 &dL_dconic[4 * id + 3], &dL_dopacity[id], &dL_dcolor[3 * id], &dL_dcolor[3 * id + 1],
					                 &dL_dcolor[3 * id + 2], &dL_ddepth[id]};
					for (int q = 0; q < 10; q++) {
#pragma omp atomic
						*dst[q] += v[q];
					}
				}
			}
	}
}

/* CR/backward.cu:208-290,301-345: gradient of (a,b,c)=cov2D w.r.t. cov3D, mean and pose. */
static void cov2d_backward(const real* mean, real fx, real fy, real tanx, real tany, const real* cov3D,
                           const float* view, real dL_da, real dL_db, real dL_dc, real* dL_dcov /*6, assigned*/,
                           real* dL_dmean /*3, assigned*/, real* dtau /*6, += */)
{
	cov2d_fw f;
	cov2d_forward(mean, fx, fy, tanx, tany, cov3D, view, &f);
	const real(*A)[3] = f.A;
	if (dL_dcov) {
		dL_dcov[0] = A[0][0] * A[0][0] * dL_da + A[0][0] * A[1][0] * dL_db + A[1][0] * A[1][0] * dL_dc;
		dL_dcov[3] = A[0][1] * A[0][1] * dL_da + A[0][1] * A[1][1] * dL_db + A[1][1] * A[1][1] * dL_dc;
		dL_dcov[5] = A[0][2] * A[0][2] * dL_da + A[0][2] * A[1][2] * dL_db + A[1][2] * A[1][2] * dL_dc;
		dL_dcov[1] = 2 * A[0][0] * A[0][1] * dL_da + (A[0][0] * A[1][1] + A[0][1] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][1] * dL_dc;
		dL_dcov[2] = 2 * A[0][0] * A[0][2] * dL_da + (A[0][0] * A[1][2] + A[0][2] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][2] * dL_dc;
		dL_dcov[4] = 2 * A[0][2] * A[0][1] * dL_da + (A[0][1] * A[1][2] + A[0][2] * A[1][1]) * dL_db + 2 * A[1][1] * A[1][2] * dL_dc;
	}
	real V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
	real dT0[3], dT1[3];
	for (int k = 0; k < 3; k++) {
		real a0v = A[0][0] * V[k][0] + A[0][1] * V[k][1] + A[0][2] * V[k][2];
		real a1v = A[1][0] * V[k][0] + A[1][1] * V[k][1] + A[1][2] * V[k][2];
		dT0[k] = 2 * a0v * dL_da + a1v * dL_db;
		dT1[k] = 2 * a1v * dL_dc + a0v * dL_db;
	}
	real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
	for (int j = 0; j < 3; j++) {
		dJ00 += VR(view, 0, j) * dT0[j];
		dJ02 += VR(view, 2, j) * dT0[j];
		dJ11 += VR(view, 1, j) * dT1[j];
		dJ12 += VR(view, 2, j) * dT1[j];
	}
	real tz = 1 / f.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
	real g[3];
	g[0] = f.xmul * -fx * tz2 * dJ02;
	g[1] = f.ymul * -fy * tz2 * dJ12;
	g[2] = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * f.t[0]) * tz3 * dJ02 + (2 * fy * f.t[1]) * tz3 * dJ12;
	/* pose via t: [I | -t^x] with the CLAMPED t (CR/backward.cu:275-290) */
	dtau[0] += g[0]; dtau[1] += g[1]; dtau[2] += g[2];
	dtau[3] += f.t[1] * g[2] - f.t[2] * g[1];
	dtau[4] += f.t[2] * g[0] - f.t[0] * g[2];
	dtau[5] += f.t[0] * g[1] - f.t[1] * g[0];
	if (dL_dmean)
		for (int c = 0; c < 3; c++) dL_dmean[c] = VR(view, 0, c) * g[0] + VR(view, 1, c) * g[1] + VR(view, 2, c) * g[2];
	/* pose via W (columns of R_cw), CR/backward.cu:301-345: dtheta = sum_k R[:,k] x dL/dR[:,k] */
	for (int k = 0; k < 3; k++) {
		real dW[3] = {f.J00 * dT0[k], f.J11 * dT1[k], f.J02 * dT0[k] + f.J12 * dT1[k]};
		real c[3] = {VR(view, 0, k), VR(view, 1, k), VR(view, 2, k)};
		dtau[3] += c[1] * dW[2] - c[2] * dW[1];
		dtau[4] += c[2] * dW[0] - c[0] * dW[2];
		dtau[5] += c[0] * dW[1] - c[1] * dW[0];
	}
}

/* Known-answer hooks (3DGS_Analytical_Jacobian.ipynb): rows of d(a,b,c)/dtau and d(ndc mean)/dtau. */
void FN(gso_cov2d_pose_jacobian)(const double* mean, double fx, double fy, double tanx, double tany, const double* cov3D,
                                 const float* view, double* jac /*3x6*/)
{
	real m[3] = {(real)mean[0], (real)mean[1], (real)mean[2]}, c3[6];
	for (int i = 0; i < 6; i++) c3[i] = (real)cov3D[i];
	for (int comp = 0; comp < 3; comp++) {
		real dt[6] = {0, 0, 0, 0, 0, 0};
		cov2d_backward(m, (real)fx, (real)fy, (real)tanx, (real)tany, c3, view, comp == 0, comp == 1, comp == 2, 0, 0, dt);
		for (int i = 0; i < 6; i++) jac[comp * 6 + i] = dt[i];
	}
}

/* CR/backward.cu:543-597: pose gradient through the NDC mean; g2 = dL/dmean2D (NDC units) */
static void mean2d_pose_backward(const real* m, const float* view, const float* proj, const float* proj_raw,
                                 real g2x, real g2y, real* dtau)
{
	real hom[4];
	for (int r = 0; r < 4; r++)
		hom[r] = (real)proj[r] * m[0] + (real)proj[4 + r] * m[1] + (real)proj[8 + r] * m[2] + (real)proj[12 + r];
	real m_w = (real)1 / (hom[3] + (real)0.0000001f);
	real alpha = m_w, beta = -hom[0] * m_w * m_w, gamma = -hom[1] * m_w * m_w;
	real a = proj_raw[0], b = proj_raw[5], e = proj_raw[11];
	real pC[3];
	for (int r = 0; r < 3; r++)
		pC[r] = VR(view, r, 0) * m[0] + VR(view, r, 1) * m[1] + VR(view, r, 2) * m[2] + VR(view, r, 3);
	real v1[3] = {alpha * a, 0, beta * e}, v2[3] = {0, alpha * b, gamma * e};
	real c1[3] = {pC[1] * v1[2] - pC[2] * v1[1], pC[2] * v1[0] - pC[0] * v1[2], pC[0] * v1[1] - pC[1] * v1[0]};
	real c2[3] = {pC[1] * v2[2] - pC[2] * v2[1], pC[2] * v2[0] - pC[0] * v2[2], pC[0] * v2[1] - pC[1] * v2[0]};
	for (int i = 0; i < 3; i++) {
		dtau[i] += g2x * v1[i] + g2y * v2[i];
		dtau[3 + i] += g2x * c1[i] + g2y * c2[i];
	}
}

void FN(gso_mean2d_pose_jacobian)(const double* mean, const float* view, const float* proj, const float* proj_raw,
                                  double* jac /*2x6*/)
{
	real m[3] = {(real)mean[0], (real)mean[1], (real)mean[2]};
	for (int comp = 0; comp < 2; comp++) {
		real dt[6] = {0, 0, 0, 0, 0, 0};
		mean2d_pose_backward(m, view, proj, proj_raw, comp == 0, comp == 1, dt);
		for (int i = 0; i < 6; i++) jac[comp * 6 + i] = dt[i];
	}
}

/* CR/backward.cu:21-145 computeColorFromSH backward; adds to dL_dmean, dtau[0:3]; assigns dL_dsh */
static void sh_backward(int deg, int M, const float* mean, const float* campos, const float* sh, const uint8_t* clamped,
                        const real* dL_dcolor, real* dL_dmean, real* dL_dsh, real* dtau)
{
	real dir_o[3] = {(real)mean[0] - (real)campos[0], (real)mean[1] - (real)campos[1], (real)mean[2] - (real)campos[2]};
	real len = (real)sqrt((double)(dir_o[0] * dir_o[0] + dir_o[1] * dir_o[1] + dir_o[2] * dir_o[2]));
	real x = dir_o[0] / len, y = dir_o[1] / len, z = dir_o[2] / len;
	real dRGB[3] = {dL_dcolor[0] * (clamped[0] ? 0 : 1), dL_dcolor[1] * (clamped[1] ? 0 : 1), dL_dcolor[2] * (clamped[2] ? 0 : 1)};
	real ddir[3] = {0, 0, 0};
	(void)M;
	for (int ch = 0; ch < 3; ch++) {
#define SH(i) ((real)sh[(i) * 3 + ch])
#define DSH(i) dL_dsh[(i) * 3 + ch]
		real g = dRGB[ch];
		real dx = 0, dy = 0, dz = 0;
		DSH(0) = SH_C0 * g;
		if (deg > 0) {
			DSH(1) = -SH_C1 * y * g; DSH(2) = SH_C1 * z * g; DSH(3) = -SH_C1 * x * g;
			dx = -SH_C1 * SH(3); dy = -SH_C1 * SH(1); dz = SH_C1 * SH(2);
			if (deg > 1) {
				real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				DSH(4) = SH_C2[0] * xy * g; DSH(5) = SH_C2[1] * yz * g; DSH(6) = SH_C2[2] * (2 * zz - xx - yy) * g;
				DSH(7) = SH_C2[3] * xz * g; DSH(8) = SH_C2[4] * (xx - yy) * g;
				dx += SH_C2[0] * y * SH(4) + SH_C2[2] * 2 * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2 * x * SH(8);
				dy += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2 * -y * SH(6) + SH_C2[4] * 2 * -y * SH(8);
				dz += SH_C2[1] * y * SH(5) + SH_C2[2] * 2 * 2 * z * SH(6) + SH_C2[3] * x * SH(7);
				if (deg > 2) {
					DSH(9) = SH_C3[0] * y * (3 * xx - yy) * g; DSH(10) = SH_C3[1] * xy * z * g;
					DSH(11) = SH_C3[2] * y * (4 * zz - xx - yy) * g; DSH(12) = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * g;
					DSH(13) = SH_C3[4] * x * (4 * zz - xx - yy) * g; DSH(14) = SH_C3[5] * z * (xx - yy) * g;
					DSH(15) = SH_C3[6] * x * (xx - 3 * yy) * g;
					dx += SH_C3[0] * SH(9) * 3 * 2 * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2 * xy +
					      SH_C3[3] * SH(12) * -3 * 2 * xz + SH_C3[4] * SH(13) * (-3 * xx + 4 * zz - yy) +
					      SH_C3[5] * SH(14) * 2 * xz + SH_C3[6] * SH(15) * 3 * (xx - yy);
					dy += SH_C3[0] * SH(9) * 3 * (xx - yy) + SH_C3[1] * SH(10) * xz + SH_C3[2] * SH(11) * (-3 * yy + 4 * zz - xx) +
					      SH_C3[3] * SH(12) * -3 * 2 * yz + SH_C3[4] * SH(13) * -2 * xy + SH_C3[5] * SH(14) * -2 * yz +
					      SH_C3[6] * SH(15) * -3 * 2 * xy;
					dz += SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4 * 2 * yz + SH_C3[3] * SH(12) * 3 * (2 * zz - xx - yy) +
					      SH_C3[4] * SH(13) * 4 * 2 * xz + SH_C3[5] * SH(14) * (xx - yy);
				}
			}
		}
#undef SH
#undef DSH
		ddir[0] += dx * g; ddir[1] += dy * g; ddir[2] += dz * g;
	}
	/* dnormvdv (CR/auxiliary.h:107-117) */
	real sum2 = dir_o[0] * dir_o[0] + dir_o[1] * dir_o[1] + dir_o[2] * dir_o[2];
	real inv32 = (real)1 / (real)sqrt((double)(sum2 * sum2 * sum2));
	real dm[3];
	dm[0] = ((+sum2 - dir_o[0] * dir_o[0]) * ddir[0] - dir_o[1] * dir_o[0] * ddir[1] - dir_o[2] * dir_o[0] * ddir[2]) * inv32;
	dm[1] = (-dir_o[0] * dir_o[1] * ddir[0] + (sum2 - dir_o[1] * dir_o[1]) * ddir[1] - dir_o[2] * dir_o[1] * ddir[2]) * inv32;
	dm[2] = (-dir_o[0] * dir_o[2] * ddir[0] - dir_o[1] * dir_o[2] * ddir[1] + (sum2 - dir_o[2] * dir_o[2]) * ddir[2]) * inv32;
	for (int i = 0; i < 3; i++) { dL_dmean[i] += dm[i]; dtau[i] += -dm[i]; }
}

/* CR/backward.cu:426-489 computeCov3D backward (no quaternion-normalisation Jacobian, :488) */
static void cov3d_backward(const float* scale, real mod, const float* rot, const real* dL_dcov3D, real* dL_dscale, real* dL_drot)
{
	real r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	real Rq[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)},
	                 {2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)},
	                 {2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)}};
	real s[3] = {mod * (real)scale[0], mod * (real)scale[1], mod * (real)scale[2]};
	real M[3][3]; /* M = S * Rq^T */
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) M[i][j] = s[i] * Rq[j][i];
	real dS[3][3] = {{dL_dcov3D[0], (real)0.5 * dL_dcov3D[1], (real)0.5 * dL_dcov3D[2]},
	                 {(real)0.5 * dL_dcov3D[1], dL_dcov3D[3], (real)0.5 * dL_dcov3D[4]},
	                 {(real)0.5 * dL_dcov3D[2], (real)0.5 * dL_dcov3D[4], dL_dcov3D[5]}};
	real dM[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) dM[i][j] = 2 * (M[i][0] * dS[0][j] + M[i][1] * dS[1][j] + M[i][2] * dS[2][j]);
	for (int i = 0; i < 3; i++) dL_dscale[i] = Rq[0][i] * dM[i][0] + Rq[1][i] * dM[i][1] + Rq[2][i] * dM[i][2];
	real G[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) G[i][j] = s[i] * dM[i][j];
	dL_drot[0] = 2 * z * (G[0][1] - G[1][0]) + 2 * y * (G[2][0] - G[0][2]) + 2 * x * (G[1][2] - G[2][1]);
	dL_drot[1] = 2 * y * (G[1][0] + G[0][1]) + 2 * z * (G[2][0] + G[0][2]) + 2 * r * (G[1][2] - G[2][1]) - 4 * x * (G[2][2] + G[1][1]);
	dL_drot[2] = 2 * x * (G[1][0] + G[0][1]) + 2 * r * (G[2][0] - G[0][2]) + 2 * z * (G[1][2] + G[2][1]) - 4 * y * (G[2][2] + G[0][0]);
	dL_drot[3] = 2 * r * (G[0][1] - G[1][0]) + 2 * x * (G[2][0] + G[0][2]) + 2 * y * (G[1][2] + G[2][1]) - 4 * z * (G[1][1] + G[0][0]);
}

/* ------------------------------------------------------------------------------------------
 * Stage 5: per-Gaussian backward.  computeCov2DCUDA then preprocessCUDA (launch order matters,
 * CR/backward.cu:905-944; computeCov2DCUDA ASSIGNS dL_dmeans, :299).  Output arrays must be
 * zero-filled by the caller; cov3D = geometry-state cov3D or cov3D_precomp.
 * dL_dtau is [P,6] like the reference (summed in Python there, py/__init__.py:162-164).
 * ------------------------------------------------------------------------------------------ */
void FN(gso_preprocess_bwd)(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                            const uint8_t* clamped, const float* scales, const float* rotations, float scale_modifier,
                            const real* cov3D, const float* view, const float* proj, const float* proj_raw,
                            const float* campos, int W, int H, float tan_fovx, float tan_fovy, const real* dL_dmean2D,
                            const real* dL_dconic, const real* dL_dcolor, const real* dL_ddepth, real* dL_dmeans3D,
                            real* dL_dcov3D, real* dL_dsh, real* dL_dscale, real* dL_drot, real* dL_dtau)
{
	const real fy = (real)H / ((real)2 * (real)tan_fovy);
	const real fx = (real)W / ((real)2 * (real)tan_fovx);
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		real m[3] = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
		real* dtau = dL_dtau + 6 * idx;
		/* --- computeCov2DCUDA --- */
		cov2d_fw f;
		cov2d_forward(m, fx, fy, (real)tan_fovx, (real)tan_fovy, cov3D + 6 * idx, view, &f);
		real a = f.a, b = f.b, c = f.c;
		real dcx = dL_dconic[4 * idx], dcy = dL_dconic[4 * idx + 1], dcz = dL_dconic[4 * idx + 3];
		real denom = a * c - b * b;
		real dL_da = 0, dL_db = 0, dL_dc = 0;
		real denom2inv = (real)1 / ((denom * denom) + (real)0.0000001f);
		if (denom2inv != 0) {
			dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
			dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
			dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
		}
		real dmean[3];
		cov2d_backward(m, fx, fy, (real)tan_fovx, (real)tan_fovy, cov3D + 6 * idx, view, dL_da, dL_db, dL_dc,
		               dL_dcov3D + 6 * idx, dmean, dtau);
		if (denom2inv == 0)
			for (int i = 0; i < 6; i++) dL_dcov3D[6 * idx + i] = 0;
		for (int i = 0; i < 3; i++) dL_dmeans3D[3 * idx + i] = dmean[i];
		/* --- preprocessCUDA (backward) --- */
		real hom[4];
		for (int r = 0; r < 4; r++)
			hom[r] = (real)proj[r] * m[0] + (real)proj[4 + r] * m[1] + (real)proj[8 + r] * m[2] + (real)proj[12 + r];
		real m_w = (real)1 / (hom[3] + (real)0.0000001f);
		real mul1 = hom[0] * m_w * m_w, mul2 = hom[1] * m_w * m_w;
		real g2x = dL_dmean2D[3 * idx], g2y = dL_dmean2D[3 * idx + 1];
		for (int k = 0; k < 3; k++)
			dL_dmeans3D[3 * idx + k] += ((real)proj[4 * k] * m_w - (real)proj[4 * k + 3] * mul1) * g2x +
			                            ((real)proj[4 * k + 1] * m_w - (real)proj[4 * k + 3] * mul2) * g2y;
		mean2d_pose_backward(m, view, proj, proj_raw, g2x, g2y, dtau);
		/* depth path, CR/backward.cu:603-613 */
		real dz = dL_ddepth[idx];
		dL_dmeans3D[3 * idx] += dz * VR(view, 2, 0);
		dL_dmeans3D[3 * idx + 1] += dz * VR(view, 2, 1);
		dL_dmeans3D[3 * idx + 2] += dz * VR(view, 2, 2);
		real pC[3];
		for (int r = 0; r < 3; r++)
			pC[r] = VR(view, r, 0) * m[0] + VR(view, r, 1) * m[1] + VR(view, r, 2) * m[2] + VR(view, r, 3);
		dtau[2] += dz;
		dtau[3] += dz * pC[1];
		dtau[4] += dz * -pC[0];
		if (shs)
			sh_backward(D, M, means3D + 3 * idx, campos, shs + (size_t)idx * M * 3, clamped + 3 * idx, dL_dcolor + 3 * idx,
			            dL_dmeans3D + 3 * idx, dL_dsh + (size_t)idx * M * 3, dtau);
		if (scales)
			cov3d_backward(scales + 3 * idx, (real)scale_modifier, rotations + 4 * idx, dL_dcov3D + 6 * idx, dL_dscale + 3 * idx,
			               dL_drot + 4 * idx);
	}
}

int FN(gso_real_bytes)(void) { return (int)sizeof(real); }
