/*
 * oracle/gs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32, optional OpenMP) of the differentiable Gaussian
 * rasterizer that ManiGaussian calls (the LangSplat fork of diff_gaussian_rasterization,
 * "DGR" = third_party/gaussian-splatting/submodules/diff-gaussian-rasterization).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's library; nothing under manigaussian_b200/ does.
 *
 * Pinning: the reference ships no tests or golden vectors for this path,
 * so this restatement is pinned against outputs of the reference itself: the unmodified
 * DGR sources compiled into oracle/_ref (oracle/build_ref.py) are run on a B200 by
 * tests/golden/make_golden.py and the resulting small fixtures are committed under
 * tests/golden/.  tests/test_oracle_cpu.py::test_oracle_against_reference_golden checks this file against them.
 *
 * Every function cites the reference file:line it follows.  Channel count for the
 * "language feature" planes is a run-time value F here (compile-time
 * NUM_CHANNELS_language_feature in the reference, config.h:16).
 *
 * Matrix convention: glm::mat3 is column-major, m[c][r]; products are evaluated as
 * (A*B)[j][i] = A[0][i]*B[j][0] + A[1][i]*B[j][1] + A[2][i]*B[j][2], left to right.
 * Built with -ffp-contract=off: plain IEEE fp32, no FMA contraction, so last-bit
 * differences against nvcc's contracted code are expected (tests use tolerances for
 * floats and stage-wise exact checks for integers).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16 /* config.h:17 */
#define BLOCK_Y 16 /* config.h:18 */

/* auxiliary.h:22-39 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
	-1.0925484305920792f, 0.5462742152960396f };
static const float SH_C3[] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
	0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f };

typedef struct { float x, y, z; } f3;
typedef struct { float m[3][3]; } mat3; /* m[col][row] like glm */

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

static inline mat3 m3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2)
{
	mat3 r;
	r.m[0][0] = x0; r.m[0][1] = y0; r.m[0][2] = z0;
	r.m[1][0] = x1; r.m[1][1] = y1; r.m[1][2] = z1;
	r.m[2][0] = x2; r.m[2][1] = y2; r.m[2][2] = z2;
	return r;
}
static inline mat3 m3mul(mat3 a, mat3 b)
{
	mat3 r;
	for (int j = 0; j < 3; j++)
		for (int i = 0; i < 3; i++)
			r.m[j][i] = a.m[0][i] * b.m[j][0] + a.m[1][i] * b.m[j][1] + a.m[2][i] * b.m[j][2];
	return r;
}
static inline mat3 m3t(mat3 a)
{
	mat3 r;
	for (int j = 0; j < 3; j++)
		for (int i = 0; i < 3; i++)
			r.m[j][i] = a.m[i][j];
	return r;
}
static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* auxiliary.h:58-66 */
static inline f3 transformPoint4x3(f3 p, const float* m)
{
	f3 t = {
		m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
		m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
		m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14] };
	return t;
}
/* auxiliary.h:68-77 */
static inline void transformPoint4x4(f3 p, const float* m, float out[4])
{
	out[0] = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
	out[1] = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
	out[2] = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
	out[3] = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
}
/* auxiliary.h:90-97 */
static inline f3 transformVec4x3Transpose(f3 p, const float* m)
{
	f3 t = {
		m[0] * p.x + m[1] * p.y + m[2] * p.z,
		m[4] * p.x + m[5] * p.y + m[6] * p.z,
		m[8] * p.x + m[9] * p.y + m[10] * p.z };
	return t;
}
/* auxiliary.h:41-44 -- evaluated in double (the literals are doubles), rounded to float */
static inline float ndc2Pix(float v, int S) { return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

/* auxiliary.h:46-56 */
static inline void getRect(float px, float py, int max_radius, uint32_t rmin[2], uint32_t rmax[2], uint32_t gx, uint32_t gy)
{
	int a;
	a = (int)((px - max_radius) / BLOCK_X); if (a < 0) a = 0; rmin[0] = (uint32_t)a < gx ? (uint32_t)a : gx;
	a = (int)((py - max_radius) / BLOCK_Y); if (a < 0) a = 0; rmin[1] = (uint32_t)a < gy ? (uint32_t)a : gy;
	a = (int)((px + max_radius + BLOCK_X - 1) / BLOCK_X); if (a < 0) a = 0; rmax[0] = (uint32_t)a < gx ? (uint32_t)a : gx;
	a = (int)((py + max_radius + BLOCK_Y - 1) / BLOCK_Y); if (a < 0) a = 0; rmax[1] = (uint32_t)a < gy ? (uint32_t)a : gy;
}

/* rasterizer_impl.cu:35-50 */
uint32_t gso_get_higher_msb(uint32_t n)
{
	uint32_t msb = sizeof(n) * 4;
	uint32_t step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

/* forward.cu:119-153 (quaternion used as given, not normalised, :128) */
static void computeCov3D(const float* scale, float mod, const float* rot, float* cov3D)
{
	float sx = mod * scale[0], sy = mod * scale[1], sz = mod * scale[2];
	mat3 S = m3(sx, 0, 0, 0, sy, 0, 0, 0, sz);
	float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	mat3 R = m3(
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
	mat3 M = m3mul(S, R);
	mat3 Sigma = m3mul(m3t(M), M);
	cov3D[0] = Sigma.m[0][0]; cov3D[1] = Sigma.m[0][1]; cov3D[2] = Sigma.m[0][2];
	cov3D[3] = Sigma.m[1][1]; cov3D[4] = Sigma.m[1][2]; cov3D[5] = Sigma.m[2][2];
}

/* shared by forward.cu:75-114 and backward.cu:163-196: t (clamped), J, W, T=W*J, Vrk */
static void cov2d_terms(f3 mean, float fx, float fy, float tanx, float tany, const float* cov3D, const float* vm,
	f3* t_out, float* txtz_out, float* tytz_out, mat3* T_out, mat3* Vrk_out, mat3* W_out)
{
	f3 t = transformPoint4x3(mean, vm);
	const float limx = 1.3f * tanx, limy = 1.3f * tany;
	const float txtz = t.x / t.z, tytz = t.y / t.z;
	t.x = fminf_(limx, fmaxf_(-limx, txtz)) * t.z;
	t.y = fminf_(limy, fmaxf_(-limy, tytz)) * t.z;
	mat3 J = m3(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z),
		0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z),
		0, 0, 0);
	mat3 W = m3(vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]);
	*T_out = m3mul(W, J);
	*Vrk_out = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
	*W_out = W; *t_out = t; *txtz_out = txtz; *tytz_out = tytz;
}

/* forward.cu:21-72 */
static void computeColorFromSH(int idx, int deg, int max_coeffs, const float* means, const float* campos,
	const float* shs, uint8_t* clamped, float out[3])
{
	float dir[3] = { means[3 * idx] - campos[0], means[3 * idx + 1] - campos[1], means[3 * idx + 2] - campos[2] };
	float len = sqrtf(dot3(dir, dir));
	dir[0] /= len; dir[1] /= len; dir[2] /= len;
	const float* sh = shs + (size_t)idx * max_coeffs * 3;
	float x = dir[0], y = dir[1], z = dir[2];
	for (int c = 0; c < 3; c++) {
#define SH(k) sh[3 * (k) + c]
		float result = SH_C0 * SH(0);
		if (deg > 0) {
			result = result - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				result = result + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * (2.0f * zz - xx - yy) * SH(6) +
					SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
				if (deg > 2) {
					result = result + SH_C3[0] * y * (3.0f * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
						SH_C3[2] * y * (4.0f * zz - xx - yy) * SH(11) +
						SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) +
						SH_C3[4] * x * (4.0f * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
						SH_C3[6] * x * (xx - 3.0f * yy) * SH(15);
				}
			}
		}
#undef SH
		result += 0.5f;
		clamped[3 * idx + c] = (result < 0);
		out[c] = result < 0.0f ? 0.0f : result;
	}
}

/*
 * K1: forward.cu:156-257 (preprocessCUDA) + auxiliary.h:139-164 (in_frustum).
 * Null pointers mean "not provided" exactly like the reference (forward.cu:206,242).
 * Outputs must be caller-allocated: depths[P], radii[P], means2D[2P], cov3Ds[6P], conic_opacity[4P],
 * rgb[3P], clamped[3P], tiles_touched[P].  Untouched entries are left as the caller initialised them
 * except radii/tiles_touched, which are zeroed first (forward.cu:191-192).
 */
void gso_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
	const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
	const float* colors_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	int W, int H, float tan_fovx, float tan_fovy,
	int* radii, float* means2D, float* depths, float* cov3Ds, float* rgb, float* conic_opacity,
	uint8_t* clamped, uint32_t* tiles_touched)
{
	const float focal_y = H / (2.0f * tan_fovy); /* rasterizer_impl.cu:225-226 */
	const float focal_x = W / (2.0f * tan_fovx);
	const uint32_t gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0;
		tiles_touched[idx] = 0;
		f3 p_orig = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
		float p_hom[4];
		transformPoint4x4(p_orig, projmatrix, p_hom);
		float p_w = 1.0f / (p_hom[3] + 0.0000001f);
		float p_proj[3] = { p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w };
		f3 p_view = transformPoint4x3(p_orig, viewmatrix);
		if (p_view.z <= 0.2f) continue; /* auxiliary.h:154 */

		const float* cov3D;
		if (cov3D_precomp) cov3D = cov3D_precomp + 6 * (size_t)idx;
		else { computeCov3D(scales + 3 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx, cov3Ds + 6 * (size_t)idx); cov3D = cov3Ds + 6 * (size_t)idx; }

		/* computeCov2D, forward.cu:75-114 */
		f3 t; float txtz, tytz; mat3 T, Vrk, Wm;
		cov2d_terms(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, &t, &txtz, &tytz, &T, &Vrk, &Wm);
		mat3 cov = m3mul(m3mul(m3t(T), m3t(Vrk)), T);
		float cx = cov.m[0][0] + 0.3f, cy = cov.m[0][1], cz = cov.m[1][1] + 0.3f;

		float det = (cx * cz - cy * cy);
		if (det == 0.0f) continue;
		float det_inv = 1.f / det;
		float conic[3] = { cz * det_inv, -cy * det_inv, cx * det_inv };
		float mid = 0.5f * (cx + cz);
		float lambda1 = mid + sqrtf(fmaxf_(0.1f, mid * mid - det));
		float lambda2 = mid - sqrtf(fmaxf_(0.1f, mid * mid - det));
		float my_radius = ceilf(3.f * sqrtf(fmaxf_(lambda1, lambda2)));
		float px = ndc2Pix(p_proj[0], W), py = ndc2Pix(p_proj[1], H);
		uint32_t rmin[2], rmax[2];
		getRect(px, py, (int)my_radius, rmin, rmax, gx, gy);
		if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;

		if (!colors_precomp) {
			float c[3];
			computeColorFromSH(idx, D, M, means3D, cam_pos, shs, clamped, c);
			rgb[3 * idx] = c[0]; rgb[3 * idx + 1] = c[1]; rgb[3 * idx + 2] = c[2];
		}
		depths[idx] = p_view.z;
		radii[idx] = (int)my_radius;
		means2D[2 * idx] = px; means2D[2 * idx + 1] = py;
		conic_opacity[4 * idx] = conic[0]; conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2]; conic_opacity[4 * idx + 3] = opacities[idx];
		tiles_touched[idx] = (rmax[1] - rmin[1]) * (rmax[0] - rmin[0]);
	}
}

/* K10: rasterizer_impl.cu:54-66 */
void gso_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present)
{
	(void)projmatrix;
	for (int idx = 0; idx < P; idx++) {
		f3 p = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
		present[idx] = transformPoint4x3(p, viewmatrix).z > 0.2f;
	}
}

/* K2: rasterizer_impl.cu:280 -- inclusive sum; returns R = num_rendered (:284) */
uint32_t gso_scan(int P, const uint32_t* tiles_touched, uint32_t* point_offsets)
{
	uint32_t s = 0;
	for (int i = 0; i < P; i++) { s += tiles_touched[i]; point_offsets[i] = s; }
	return s;
}

/* K3: rasterizer_impl.cu:70-111 (duplicateWithKeys) */
void gso_duplicate_with_keys(int P, const float* means2D, const float* depths, const uint32_t* offsets,
	const int* radii, int W, int H, uint64_t* keys, uint32_t* values)
{
	const uint32_t gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	for (int idx = 0; idx < P; idx++) {
		if (radii[idx] <= 0) continue;
		uint32_t off = idx == 0 ? 0 : offsets[idx - 1];
		uint32_t rmin[2], rmax[2];
		getRect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], rmin, rmax, gx, gy);
		uint32_t dbits;
		memcpy(&dbits, &depths[idx], 4);
		for (uint32_t y = rmin[1]; y < rmax[1]; y++)
			for (uint32_t x = rmin[0]; x < rmax[0]; x++) {
				uint64_t key = y * gx + x;
				key <<= 32;
				key |= dbits;
				keys[off] = key;
				values[off] = (uint32_t)idx;
				off++;
			}
	}
}

/* K4: rasterizer_impl.cu:303-311 -- stable LSD radix sort over bits [0, 32+bit) (cub::DeviceRadixSort::SortPairs).
 * tmp_keys/tmp_vals: scratch of R entries.  Result ends up in keys_out/vals_out. */
void gso_sort_pairs(uint32_t R, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
	int end_bit)
{
	if (R == 0) return;
	uint64_t* ka = (uint64_t*)malloc(sizeof(uint64_t) * R), * kb = (uint64_t*)malloc(sizeof(uint64_t) * R);
	uint32_t* va = (uint32_t*)malloc(sizeof(uint32_t) * R), * vb = (uint32_t*)malloc(sizeof(uint32_t) * R);
	memcpy(ka, keys_in, sizeof(uint64_t) * R);
	memcpy(va, vals_in, sizeof(uint32_t) * R);
	size_t* cnt = (size_t*)malloc(sizeof(size_t) * 65537);
	for (int shift = 0; shift < end_bit; shift += 16) {
		int nb = end_bit - shift < 16 ? end_bit - shift : 16;
		uint64_t mask = ((uint64_t)1 << nb) - 1;
		memset(cnt, 0, sizeof(size_t) * 65537);
		for (uint32_t i = 0; i < R; i++) cnt[((ka[i] >> shift) & mask) + 1]++;
		for (int i = 0; i < 65536; i++) cnt[i + 1] += cnt[i];
		for (uint32_t i = 0; i < R; i++) {
			size_t d = cnt[(ka[i] >> shift) & mask]++;
			kb[d] = ka[i]; vb[d] = va[i];
		}
		uint64_t* tk = ka; ka = kb; kb = tk;
		uint32_t* tv = va; va = vb; vb = tv;
	}
	memcpy(keys_out, ka, sizeof(uint64_t) * R);
	memcpy(vals_out, va, sizeof(uint32_t) * R);
	free(ka); free(kb); free(va); free(vb); free(cnt);
}

/* K5: rasterizer_impl.cu:313 (memset) + :116-138 (identifyTileRanges).  ranges: uint2[T] */
void gso_identify_tile_ranges(uint32_t L, const uint64_t* keys, int num_tiles, uint32_t* ranges)
{
	memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)num_tiles);
	for (uint32_t idx = 0; idx < L; idx++) {
		uint32_t currtile = (uint32_t)(keys[idx] >> 32);
		if (idx == 0) ranges[2 * currtile] = 0;
		else {
			uint32_t prevtile = (uint32_t)(keys[idx - 1] >> 32);
			if (currtile != prevtile) { ranges[2 * prevtile + 1] = idx; ranges[2 * currtile] = idx; }
		}
		if (idx == L - 1) ranges[2 * currtile + 1] = L;
	}
}

/*
 * K6: forward.cu:262-398 (renderCUDA forward), one pixel at a time.  The block-cooperative staging and
 * the block-wide early exit (:319) do not change any pixel's result: a pixel stops when it is `done`.
 * colors: [P,3]; feature: [P,F] or NULL (include_feature == false).
 */
void gso_render_forward(int W, int H, int F, const uint32_t* ranges, const uint32_t* point_list,
	const float* means2D, const float* colors, const float* feature, const float* conic_opacity,
	const float* bg_color, float* final_T, uint32_t* n_contrib, float* out_color, float* out_feature)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < gx * gy; tile++) {
		const int ty = tile / gx, tx = tile % gx;
		const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		float* Facc = (float*)malloc(sizeof(float) * (F > 0 ? F : 1));
		for (int ly = 0; ly < BLOCK_Y; ly++)
			for (int lx = 0; lx < BLOCK_X; lx++) {
				const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
				if (pxi >= W || pyi >= H) continue;
				const size_t pix_id = (size_t)W * pyi + pxi;
				const float pfx = (float)pxi, pfy = (float)pyi;
				float T = 1.0f;
				uint32_t contributor = 0, last_contributor = 0;
				float C[3] = { 0, 0, 0 };
				for (int ch = 0; ch < F; ch++) Facc[ch] = 0.f;
				for (uint32_t k = r0; k < r1; k++) {
					contributor++;
					const uint32_t g = point_list[k];
					const float dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
					const float* co = conic_opacity + 4 * (size_t)g;
					const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
					if (power > 0.0f) continue;
					float alpha = fminf_(0.99f, co[3] * expf(power));
					if (alpha < 1.0f / 255.0f) continue;
					float test_T = T * (1 - alpha);
					if (test_T < 0.0001f) break; /* done = true (:358-362) */
					for (int ch = 0; ch < 3; ch++) C[ch] += colors[3 * (size_t)g + ch] * alpha * T;
					for (int ch = 0; ch < F; ch++) Facc[ch] += feature[(size_t)g * F + ch] * alpha * T;
					T = test_T;
					last_contributor = contributor;
				}
				final_T[pix_id] = T;
				n_contrib[pix_id] = last_contributor;
				for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[ch] + T * bg_color[ch];
				for (int ch = 0; ch < F; ch++) out_feature[ch * HW + pix_id] = Facc[ch]; /* no background (:393) */
			}
		free(Facc);
	}
}

/* The reference sums per-pixel contributions with unordered fp32 atomics (backward.cu:541-590), so its own gradients
 * carry summation-order noise (~1e-4 relative on cancelling sums).  The oracle accumulates the same addends in double
 * and rounds once, which makes it the better yardstick for both the reference and the CUDA implementation. */
static inline void atomic_addd(double* p, double v)
{
#pragma omp atomic
	*p += v;
}

/*
 * K7: backward.cu:399-593 (renderCUDA backward), one pixel at a time, back to front.
 * Outputs (caller-zeroed, rasterize_points.cu:167-184): dL_dmean2D [P,3], dL_dconic [P,4] (slots x,y,w used),
 * dL_dopacity [P], dL_dcolors [P,3], dL_dfeature [P,F].
 */
void gso_render_backward(int P, int W, int H, int F, const uint32_t* ranges, const uint32_t* point_list,
	const float* bg_color, const float* means2D, const float* conic_opacity, const float* colors, const float* feature,
	const float* final_Ts, const uint32_t* n_contrib, const float* dL_dpixels, const float* dL_dpixels_F,
	float* dL_dmean2D, float* dL_dconic2D, float* dL_dopacity, float* dL_dcolors, float* dL_dfeature)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	const size_t HW = (size_t)H * W;
	const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H); /* :462-463 */
	const int nf = F > 0 ? F : 1;
	double* A_mean2D = (double*)calloc((size_t)P * 3, sizeof(double));
	double* A_conic = (double*)calloc((size_t)P * 4, sizeof(double));
	double* A_opac = (double*)calloc((size_t)P, sizeof(double));
	double* A_col = (double*)calloc((size_t)P * 3, sizeof(double));
	double* A_feat = (double*)calloc((size_t)P * nf, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < gx * gy; tile++) {
		const int ty = tile / gx, tx = tile % gx;
		const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		float* accum_rec_F = (float*)malloc(sizeof(float) * nf * 3);
		float* last_F = accum_rec_F + nf;
		float* dpixF = accum_rec_F + 2 * nf;
		for (int ly = 0; ly < BLOCK_Y; ly++)
			for (int lx = 0; lx < BLOCK_X; lx++) {
				const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
				if (pxi >= W || pyi >= H) continue;
				const size_t pix_id = (size_t)W * pyi + pxi;
				const float pfx = (float)pxi, pfy = (float)pyi;
				const float T_final = final_Ts[pix_id];
				float T = T_final;
				uint32_t contributor = r1 - r0;
				const uint32_t last_contributor = n_contrib[pix_id];
				float accum_rec[3] = { 0, 0, 0 }, last_color[3] = { 0, 0, 0 }, dL_dpixel[3];
				for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[i * HW + pix_id];
				for (int i = 0; i < F; i++) { accum_rec_F[i] = 0; last_F[i] = 0; dpixF[i] = dL_dpixels_F[i * HW + pix_id]; }
				float last_alpha = 0;
				for (uint32_t kk = r1; kk > r0; kk--) {
					const uint32_t g = point_list[kk - 1];
					contributor--;
					if (contributor >= last_contributor) continue;
					const float dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
					const float* co = conic_opacity + 4 * (size_t)g;
					const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
					if (power > 0.0f) continue;
					const float G = expf(power);
					const float alpha = fminf_(0.99f, co[3] * G);
					if (alpha < 1.0f / 255.0f) continue;
					T = T / (1.f - alpha);
					const float dchannel_dcolor = alpha * T;
					float dL_dalpha = 0.0f;
					for (int ch = 0; ch < 3; ch++) {
						const float c = colors[3 * (size_t)g + ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
						atomic_addd(&A_col[3 * (size_t)g + ch], dchannel_dcolor * dL_dpixel[ch]);
					}
					for (int ch = 0; ch < F; ch++) {
						const float f = feature[(size_t)g * F + ch];
						accum_rec_F[ch] = last_alpha * last_F[ch] + (1.f - last_alpha) * accum_rec_F[ch];
						last_F[ch] = f;
						dL_dalpha += (f - accum_rec_F[ch]) * dpixF[ch];
						atomic_addd(&A_feat[(size_t)g * F + ch], dchannel_dcolor * dpixF[ch]);
					}
					dL_dalpha *= T;
					last_alpha = alpha;
					float bg_dot_dpixel = 0;
					for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_dpixel[i];
					dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
					const float dL_dG = co[3] * dL_dalpha; /* no mask for the 0.99 clamp (:574) */
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * co[0] - gdy * co[1];
					const float dG_ddely = -gdy * co[2] - gdx * co[1];
					atomic_addd(&A_mean2D[3 * (size_t)g + 0], dL_dG * dG_ddelx * ddelx_dx);
					atomic_addd(&A_mean2D[3 * (size_t)g + 1], dL_dG * dG_ddely * ddely_dy);
					atomic_addd(&A_conic[4 * (size_t)g + 0], -0.5f * gdx * dx * dL_dG);
					atomic_addd(&A_conic[4 * (size_t)g + 1], -0.5f * gdx * dy * dL_dG);
					atomic_addd(&A_conic[4 * (size_t)g + 3], -0.5f * gdy * dy * dL_dG);
					atomic_addd(&A_opac[g], G * dL_dalpha);
				}
			}
		free(accum_rec_F);
	}
	for (size_t i = 0; i < (size_t)P * 3; i++) { dL_dmean2D[i] = (float)A_mean2D[i]; dL_dcolors[i] = (float)A_col[i]; }
	for (size_t i = 0; i < (size_t)P * 4; i++) dL_dconic2D[i] = (float)A_conic[i];
	for (size_t i = 0; i < (size_t)P; i++) dL_dopacity[i] = (float)A_opac[i];
	if (F > 0) for (size_t i = 0; i < (size_t)P * F; i++) dL_dfeature[i] = (float)A_feat[i];
	free(A_mean2D); free(A_conic); free(A_opac); free(A_col); free(A_feat);
}

/* auxiliary.h:107-117 */
static inline f3 dnormvdv(f3 v, f3 dv)
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	f3 r;
	r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return r;
}

/* backward.cu:20-139 (SH backward incl. the view-direction term into dL_dmeans, accumulate) */
static void computeColorFromSH_bwd(int idx, int deg, int max_coeffs, const float* means, const float* campos, const float* shs,
	const uint8_t* clamped, const float* dL_dcolor, float* dL_dmeans, float* dL_dshs)
{
	float dir_orig[3] = { means[3 * idx] - campos[0], means[3 * idx + 1] - campos[1], means[3 * idx + 2] - campos[2] };
	float len = sqrtf(dot3(dir_orig, dir_orig));
	float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
	const float* sh = shs + (size_t)idx * max_coeffs * 3;
	float* dL_dsh = dL_dshs + (size_t)idx * max_coeffs * 3;
	float dL_ddir[3] = { 0, 0, 0 };
	float dRGBdx[3], dRGBdy[3], dRGBdz[3], dL_dRGB[3];
	for (int c = 0; c < 3; c++) {
		dL_dRGB[c] = dL_dcolor[3 * idx + c] * (clamped[3 * idx + c] ? 0.f : 1.f);
		dRGBdx[c] = dRGBdy[c] = dRGBdz[c] = 0.f;
	}
#define SH(k) sh[3 * (k) + c]
#define DSH(k, v) for (int c = 0; c < 3; c++) dL_dsh[3 * (k) + c] = (v) * dL_dRGB[c]
	DSH(0, SH_C0);
	if (deg > 0) {
		float dRGBdsh1 = -SH_C1 * y, dRGBdsh2 = SH_C1 * z, dRGBdsh3 = -SH_C1 * x;
		DSH(1, dRGBdsh1); DSH(2, dRGBdsh2); DSH(3, dRGBdsh3);
		for (int c = 0; c < 3; c++) { dRGBdx[c] = -SH_C1 * SH(3); dRGBdy[c] = -SH_C1 * SH(1); dRGBdz[c] = SH_C1 * SH(2); }
		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
			DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
			DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
			for (int c = 0; c < 3; c++) {
				dRGBdx[c] += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2.f * x * SH(8);
				dRGBdy[c] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) + SH_C2[4] * 2.f * -y * SH(8);
				dRGBdz[c] += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
			}
			if (deg > 2) {
				DSH(9, SH_C3[0] * y * (3.f * xx - yy)); DSH(10, SH_C3[1] * xy * z);
				DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy)); DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
				DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, SH_C3[5] * z * (xx - yy));
				DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
				for (int c = 0; c < 3; c++) {
					dRGBdx[c] += (SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2.f * xy +
						SH_C3[3] * SH(12) * -3.f * 2.f * xz + SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
						SH_C3[5] * SH(14) * 2.f * xz + SH_C3[6] * SH(15) * 3.f * (xx - yy));
					dRGBdy[c] += (SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
						SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
						SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz + SH_C3[6] * SH(15) * -3.f * 2.f * xy);
					dRGBdz[c] += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
						SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
						SH_C3[5] * SH(14) * (xx - yy));
				}
			}
		}
	}
#undef SH
#undef DSH
	dL_ddir[0] = dot3(dRGBdx, dL_dRGB); dL_ddir[1] = dot3(dRGBdy, dL_dRGB); dL_ddir[2] = dot3(dRGBdz, dL_dRGB);
	f3 v = { dir_orig[0], dir_orig[1], dir_orig[2] }, dv = { dL_ddir[0], dL_ddir[1], dL_ddir[2] };
	f3 dm = dnormvdv(v, dv);
	dL_dmeans[3 * idx] += dm.x; dL_dmeans[3 * idx + 1] += dm.y; dL_dmeans[3 * idx + 2] += dm.z;
}

/* backward.cu:278-341 (cov3D -> scale, quaternion; written un-normalised, :340) */
static void computeCov3D_bwd(int idx, const float* scale, float mod, const float* rot, const float* dL_dcov3Ds,
	float* dL_dscales, float* dL_drots)
{
	float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	mat3 R = m3(
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
	float s[3] = { mod * scale[0], mod * scale[1], mod * scale[2] };
	mat3 S = m3(s[0], 0, 0, 0, s[1], 0, 0, 0, s[2]);
	mat3 M = m3mul(S, R);
	const float* d = dL_dcov3Ds + 6 * (size_t)idx;
	mat3 dL_dSigma = m3(d[0], 0.5f * d[1], 0.5f * d[2], 0.5f * d[1], d[3], 0.5f * d[4], 0.5f * d[2], 0.5f * d[4], d[5]);
	mat3 M2;
	for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) M2.m[j][i] = 2.0f * M.m[j][i];
	mat3 dL_dM = m3mul(M2, dL_dSigma);
	mat3 Rt = m3t(R), dL_dMt = m3t(dL_dM);
	float* ds = dL_dscales + 3 * (size_t)idx;
	ds[0] = dot3(Rt.m[0], dL_dMt.m[0]);
	ds[1] = dot3(Rt.m[1], dL_dMt.m[1]);
	ds[2] = dot3(Rt.m[2], dL_dMt.m[2]);
	for (int i = 0; i < 3; i++) { dL_dMt.m[0][i] *= s[0]; dL_dMt.m[1][i] *= s[1]; dL_dMt.m[2][i] *= s[2]; }
#define Mt(a, b) dL_dMt.m[a][b]
	float* dq = dL_drots + 4 * (size_t)idx;
	dq[0] = 2 * z * (Mt(0, 1) - Mt(1, 0)) + 2 * y * (Mt(2, 0) - Mt(0, 2)) + 2 * x * (Mt(1, 2) - Mt(2, 1));
	dq[1] = 2 * y * (Mt(1, 0) + Mt(0, 1)) + 2 * z * (Mt(2, 0) + Mt(0, 2)) + 2 * r * (Mt(1, 2) - Mt(2, 1)) - 4 * x * (Mt(2, 2) + Mt(1, 1));
	dq[2] = 2 * x * (Mt(1, 0) + Mt(0, 1)) + 2 * r * (Mt(2, 0) - Mt(0, 2)) + 2 * z * (Mt(1, 2) + Mt(2, 1)) - 4 * y * (Mt(2, 2) + Mt(0, 0));
	dq[3] = 2 * r * (Mt(0, 1) - Mt(1, 0)) + 2 * x * (Mt(2, 0) + Mt(0, 2)) + 2 * y * (Mt(1, 2) + Mt(2, 1)) - 4 * z * (Mt(1, 1) + Mt(0, 0));
#undef Mt
}

/*
 * K8 + K9: backward.cu:144-274 (computeCov2DCUDA), :346-396 (preprocessCUDA bwd), launched back to back (:595-657).
 * dL_dmean3D is assigned by K8 (:273) then accumulated by K9 (:387) and the SH term (:138).
 * Outputs caller-zeroed; Gaussians with radii <= 0 are skipped in both kernels (:156,:367).
 * dL_dcolor [P,3] is the blend-stage colour gradient (input; consumed by the SH backward).
 */
void gso_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
	const uint8_t* clamped, const float* scales, const float* rotations, float scale_modifier, const float* cov3Ds,
	const float* viewmatrix, const float* projmatrix, int W, int H, float tan_fovx, float tan_fovy, const float* campos,
	const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot)
{
	const float h_y = H / (2.0f * tan_fovy), h_x = W / (2.0f * tan_fovx); /* rasterizer_impl.cu:404-405 */
#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		/* ---- K8 ---- */
		const float* cov3D = cov3Ds + 6 * (size_t)idx;
		f3 mean = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
		float dLc[3] = { dL_dconic[4 * idx], dL_dconic[4 * idx + 1], dL_dconic[4 * idx + 3] };
		f3 t; float txtz, tytz; mat3 T, Vrk, Wm;
		cov2d_terms(mean, h_x, h_y, tan_fovx, tan_fovy, cov3D, viewmatrix, &t, &txtz, &tytz, &T, &Vrk, &Wm);
		const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
		const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
		const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
		mat3 cov2D = m3mul(m3mul(m3t(T), m3t(Vrk)), T);
		float a = cov2D.m[0][0] + 0.3f, b = cov2D.m[0][1], c = cov2D.m[1][1] + 0.3f;
		/* backward.cu:201: `denom = a*c - b*b`, and `(denom - a*c)` two lines below, cancel catastrophically when
		 * b*b << a*c (large, nearly axis-aligned splats), so the result depends on how nvcc contracts them.  The
		 * sm_100 SASS of this expression is FMUL t=a*c; FFMA denom=fma(-b,b,t); FADD (denom - t) -- restated here
		 * with fmaf so the oracle reproduces the reference's rounding instead of gcc's uncontracted one. */
		const float ac = a * c;
		float denom = fmaf(-b, b, ac);
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
		float* dcov = dL_dcov3D + 6 * (size_t)idx;
#define T_(i, j) T.m[i][j]
#define V_(i, j) Vrk.m[i][j]
#define W_(i, j) Wm.m[i][j]
		if (denom2inv != 0) {
			dL_da = denom2inv * (-c * c * dLc[0] + 2 * b * c * dLc[1] + (denom - ac) * dLc[2]);
			dL_dc = denom2inv * (-a * a * dLc[2] + 2 * a * b * dLc[1] + (denom - ac) * dLc[0]);
			dL_db = denom2inv * 2 * (b * c * dLc[0] - (denom + 2 * b * b) * dLc[1] + a * b * dLc[2]);
			dcov[0] = (T_(0, 0) * T_(0, 0) * dL_da + T_(0, 0) * T_(1, 0) * dL_db + T_(1, 0) * T_(1, 0) * dL_dc);
			dcov[3] = (T_(0, 1) * T_(0, 1) * dL_da + T_(0, 1) * T_(1, 1) * dL_db + T_(1, 1) * T_(1, 1) * dL_dc);
			dcov[5] = (T_(0, 2) * T_(0, 2) * dL_da + T_(0, 2) * T_(1, 2) * dL_db + T_(1, 2) * T_(1, 2) * dL_dc);
			dcov[1] = 2 * T_(0, 0) * T_(0, 1) * dL_da + (T_(0, 0) * T_(1, 1) + T_(0, 1) * T_(1, 0)) * dL_db + 2 * T_(1, 0) * T_(1, 1) * dL_dc;
			dcov[2] = 2 * T_(0, 0) * T_(0, 2) * dL_da + (T_(0, 0) * T_(1, 2) + T_(0, 2) * T_(1, 0)) * dL_db + 2 * T_(1, 0) * T_(1, 2) * dL_dc;
			dcov[4] = 2 * T_(0, 2) * T_(0, 1) * dL_da + (T_(0, 1) * T_(1, 2) + T_(0, 2) * T_(1, 1)) * dL_db + 2 * T_(1, 1) * T_(1, 2) * dL_dc;
		} else {
			for (int i = 0; i < 6; i++) dcov[i] = 0;
		}
		float dL_dT00 = 2 * (T_(0, 0) * V_(0, 0) + T_(0, 1) * V_(0, 1) + T_(0, 2) * V_(0, 2)) * dL_da + (T_(1, 0) * V_(0, 0) + T_(1, 1) * V_(0, 1) + T_(1, 2) * V_(0, 2)) * dL_db;
		float dL_dT01 = 2 * (T_(0, 0) * V_(1, 0) + T_(0, 1) * V_(1, 1) + T_(0, 2) * V_(1, 2)) * dL_da + (T_(1, 0) * V_(1, 0) + T_(1, 1) * V_(1, 1) + T_(1, 2) * V_(1, 2)) * dL_db;
		float dL_dT02 = 2 * (T_(0, 0) * V_(2, 0) + T_(0, 1) * V_(2, 1) + T_(0, 2) * V_(2, 2)) * dL_da + (T_(1, 0) * V_(2, 0) + T_(1, 1) * V_(2, 1) + T_(1, 2) * V_(2, 2)) * dL_db;
		float dL_dT10 = 2 * (T_(1, 0) * V_(0, 0) + T_(1, 1) * V_(0, 1) + T_(1, 2) * V_(0, 2)) * dL_dc + (T_(0, 0) * V_(0, 0) + T_(0, 1) * V_(0, 1) + T_(0, 2) * V_(0, 2)) * dL_db;
		float dL_dT11 = 2 * (T_(1, 0) * V_(1, 0) + T_(1, 1) * V_(1, 1) + T_(1, 2) * V_(1, 2)) * dL_dc + (T_(0, 0) * V_(1, 0) + T_(0, 1) * V_(1, 1) + T_(0, 2) * V_(1, 2)) * dL_db;
		float dL_dT12 = 2 * (T_(1, 0) * V_(2, 0) + T_(1, 1) * V_(2, 1) + T_(1, 2) * V_(2, 2)) * dL_dc + (T_(0, 0) * V_(2, 0) + T_(0, 1) * V_(2, 1) + T_(0, 2) * V_(2, 2)) * dL_db;
		float dL_dJ00 = W_(0, 0) * dL_dT00 + W_(0, 1) * dL_dT01 + W_(0, 2) * dL_dT02;
		float dL_dJ02 = W_(2, 0) * dL_dT00 + W_(2, 1) * dL_dT01 + W_(2, 2) * dL_dT02;
		float dL_dJ11 = W_(1, 0) * dL_dT10 + W_(1, 1) * dL_dT11 + W_(1, 2) * dL_dT12;
		float dL_dJ12 = W_(2, 0) * dL_dT10 + W_(2, 1) * dL_dT11 + W_(2, 2) * dL_dT12;
#undef T_
#undef V_
#undef W_
		float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
		float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
		float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
		float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
		f3 dt = { dL_dtx, dL_dty, dL_dtz };
		f3 dmean = transformVec4x3Transpose(dt, viewmatrix);
		dL_dmean3D[3 * idx] = dmean.x; dL_dmean3D[3 * idx + 1] = dmean.y; dL_dmean3D[3 * idx + 2] = dmean.z; /* assign (:273) */

		/* ---- K9 ---- */
		const float* proj = projmatrix;
		float m_hom[4];
		transformPoint4x4(mean, proj, m_hom);
		float m_w = 1.0f / (m_hom[3] + 0.0000001f);
		float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
		float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
		const float g2x = dL_dmean2D[3 * idx], g2y = dL_dmean2D[3 * idx + 1];
		dL_dmean3D[3 * idx + 0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
		dL_dmean3D[3 * idx + 1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
		dL_dmean3D[3 * idx + 2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
		if (shs)
			computeColorFromSH_bwd(idx, D, M, means3D, campos, shs, clamped, dL_dcolor, dL_dmean3D, dL_dsh);
		if (scales)
			computeCov3D_bwd(idx, scales + 3 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx, dL_dcov3D, dL_dscale, dL_drot);
	}
}

int gso_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
void gso_set_threads(int n)
{
#ifdef _OPENMP
	omp_set_num_threads(n);
#else
	(void)n;
#endif
}
