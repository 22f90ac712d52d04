// Per-Gaussian stages: 3D->2D projection (forward), its chain-rule backward, and the frustum test.
//
// Replaces, behind the C-ABI in include/mgs_rasterizer.h:
//   forward  : FORWARD::preprocess / preprocessCUDA    (DGR/cuda_rasterizer/forward.cu:156-257)
//   backward : BACKWARD::preprocess = computeCov2DCUDA + preprocessCUDA (backward.cu:144-274, :346-396)
//              -- fused here into ONE kernel that also folds the reference's ten zero-fills
//              (rasterize_points.cu:167-184): every output row is written exactly once.
//   visible  : checkFrustum (rasterizer_impl.cu:54-66)
#include "mgs_common.cuh"
#include "mgs_kernels.h"

namespace mgs {

// SH constants (values as published with 3DGS; reference auxiliary.h:22-39)
__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
	-1.0925484305920792f, 0.5462742152960396f };
__device__ const float kSH_C3[] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
	0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f };

// Sigma = (S R)^T (S R) from scale and (un-normalised) quaternion; forward.cu:119-153
__device__ __forceinline__ void cov3d_from_scale_rot(const float* scale, float mod, const float* rot, float* cov3D)
{
	M3 S = m3(mod * scale[0], 0.f, 0.f, 0.f, mod * scale[1], 0.f, 0.f, 0.f, mod * scale[2]);
	float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	M3 R = m3(
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
	M3 Mx = mul(S, R);
	M3 Sigma = mul(transpose(Mx), Mx);
	cov3D[0] = Sigma.m[0][0]; cov3D[1] = Sigma.m[0][1]; cov3D[2] = Sigma.m[0][2];
	cov3D[3] = Sigma.m[1][1]; cov3D[4] = Sigma.m[1][2]; cov3D[5] = Sigma.m[2][2];
}

struct Cov2DTerms {
	V3 t;
	float txtz, tytz;
	M3 T, Vrk, W;
};
// Shared by the forward (forward.cu:75-107) and the backward recomputation (backward.cu:163-196)
__device__ __forceinline__ Cov2DTerms cov2d_terms(const V3& mean, float fx, float fy, float tanx, float tany,
	const float* cov3D, const float* vm)
{
	Cov2DTerms o;
	V3 t = xform4x3(mean, vm);
	const float limx = 1.3f * tanx;
	const float limy = 1.3f * tany;
	const float txtz = t.x / t.z;
	const float tytz = t.y / t.z;
	t.x = min(limx, max(-limx, txtz)) * t.z;
	t.y = min(limy, max(-limy, tytz)) * t.z;
	M3 J = m3(
		fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z),
		0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z),
		0.f, 0.f, 0.f);
	o.W = m3(vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]);
	o.T = mul(o.W, J);
	o.Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
	o.t = t; o.txtz = txtz; o.tytz = tytz;
	return o;
}

// SH -> RGB with the +0.5 offset and clamp-at-zero mask; forward.cu:21-72
__device__ __forceinline__ void sh_to_rgb(int idx, int deg, int max_coeffs, const float* means, const float* campos,
	const float* shs, uint8_t* clamped_bits, float out[3])
{
	float dx = means[3 * idx] - campos[0], dy = means[3 * idx + 1] - campos[1], dz = means[3 * idx + 2] - campos[2];
	float len = sqrtf(dx * dx + dy * dy + dz * dz);
	float x = dx / len, y = dy / len, z = dz / len;
	const float* sh = shs + (size_t)idx * max_coeffs * 3;
	uint8_t bits = 0;
#pragma unroll
	for (int c = 0; c < 3; c++) {
#define SHK(k) sh[3 * (k) + c]
		float result = kSH_C0 * SHK(0);
		if (deg > 0) {
			result = result - kSH_C1 * y * SHK(1) + kSH_C1 * z * SHK(2) - kSH_C1 * x * SHK(3);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z;
				float xy = x * y, yz = y * z, xz = x * z;
				result = result +
					kSH_C2[0] * xy * SHK(4) +
					kSH_C2[1] * yz * SHK(5) +
					kSH_C2[2] * (2.0f * zz - xx - yy) * SHK(6) +
					kSH_C2[3] * xz * SHK(7) +
					kSH_C2[4] * (xx - yy) * SHK(8);
				if (deg > 2) {
					result = result +
						kSH_C3[0] * y * (3.0f * xx - yy) * SHK(9) +
						kSH_C3[1] * xy * z * SHK(10) +
						kSH_C3[2] * y * (4.0f * zz - xx - yy) * SHK(11) +
						kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHK(12) +
						kSH_C3[4] * x * (4.0f * zz - xx - yy) * SHK(13) +
						kSH_C3[5] * z * (xx - yy) * SHK(14) +
						kSH_C3[6] * x * (xx - 3.0f * yy) * SHK(15);
				}
			}
		}
#undef SHK
		result += 0.5f;
		if (result < 0) bits |= (uint8_t)(1u << c);
		out[c] = max(result, 0.0f);
	}
	clamped_bits[idx] = bits;
}

__global__ void __launch_bounds__(256) project_fwd_kernel(ProjectFwdArgs a)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= a.P) return;

	// defaults for a culled Gaussian (forward.cu:191-192); everything the later stages read is defined
	int out_radius = 0;
	uint32_t out_tiles = 0;
	a.radii[idx] = 0;
	a.tiles_touched[idx] = 0;
	a.iota[idx] = (uint32_t)idx;
	a.depths[idx] = __uint_as_float(0x7f800000u);  // +inf: culled Gaussians sort behind every visible one

	V3 p_orig = { a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2] };
	float4 p_hom = xform4x4(p_orig, a.projmatrix);
	float p_w = 1.0f / (p_hom.w + 0.0000001f);
	float p_proj_x = p_hom.x * p_w, p_proj_y = p_hom.y * p_w;
	V3 p_view = xform4x3(p_orig, a.viewmatrix);
	if (p_view.z <= NEAR_Z) return;  // near cull only (auxiliary.h:154)

	const float* cov3D;
	if (a.cov3D_precomp != nullptr) {
		cov3D = a.cov3D_precomp + (size_t)idx * 6;
	} else {
		cov3d_from_scale_rot(a.scales + 3 * (size_t)idx, a.scale_modifier, a.rotations + 4 * (size_t)idx, a.cov3D + (size_t)idx * 6);
		cov3D = a.cov3D + (size_t)idx * 6;
	}

	Cov2DTerms ct = cov2d_terms(p_orig, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, a.viewmatrix);
	M3 cov = mul(mul(transpose(ct.T), transpose(ct.Vrk)), ct.T);
	cov.m[0][0] += 0.3f;  // low-pass (forward.cu:111-112)
	cov.m[1][1] += 0.3f;
	const float cx = cov.m[0][0], cy = cov.m[0][1], cz = cov.m[1][1];

	float det = (cx * cz - cy * cy);
	if (det == 0.0f) return;
	float det_inv = 1.f / det;
	float conic_x = cz * det_inv, conic_y = -cy * det_inv, conic_z = cx * det_inv;

	float mid = 0.5f * (cx + cz);
	float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
	float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
	float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
	float px = ndc2pix(p_proj_x, a.W), py = ndc2pix(p_proj_y, a.H);
	uint2 rmin, rmax;
	tile_rect(px, py, (int)my_radius, rmin, rmax, a.grid_x, a.grid_y);
	if ((rmax.x - rmin.x) * (rmax.y - rmin.y) == 0) return;

	float c[3];
	if (a.colors_precomp == nullptr) {
		sh_to_rgb(idx, a.D, a.M, a.means3D, a.cam_pos, a.shs, a.clamped, c);
	} else {
		c[0] = a.colors_precomp[3 * idx]; c[1] = a.colors_precomp[3 * idx + 1]; c[2] = a.colors_precomp[3 * idx + 2];
	}
	a.rgbd[idx] = make_float4(c[0], c[1], c[2], p_view.z);  // blend channel quad 0 (depth rides as a channel)

	const float opacity = a.opacities[idx];
	out_radius = (int)my_radius;
	out_tiles = (rmax.y - rmin.y) * (rmax.x - rmin.x);
	a.depths[idx] = p_view.z;
	a.radii[idx] = out_radius;
	a.means2D[idx] = make_float2(px, py);
	a.conic_opacity[idx] = make_float4(conic_x, conic_y, conic_z, opacity);
	a.tiles_touched[idx] = out_tiles;

	// Conservative footprint of {alpha >= 1/255}: 0.5 d^T Q d <= tau, tau = ln(255 * opacity).
	// Used only to skip whole 8x4 pixel blocks in the blend kernels; never changes a pixel's result.
	float hx = 1e30f, hy = 1e30f;
	const float o255 = 255.0f * opacity;
	if (!(o255 >= 0.999f)) {
		hx = hy = -1.0f;  // alpha = min(0.99, o*G) <= o < 1/255 everywhere: never contributes
	} else {
		const float tau = logf(o255) * 1.0005f + 1e-3f;
		const float ac = conic_x * conic_z;
		const float dq = ac - conic_y * conic_y;
		if (conic_x > 0.f && conic_z > 0.f && dq > 1e-3f * ac) {
			hx = sqrtf(2.0f * tau * conic_z / dq) * 1.001f + 0.01f;
			hy = sqrtf(2.0f * tau * conic_x / dq) * 1.001f + 0.01f;
		}
	}
	a.extent[idx] = make_float2(hx, hy);
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* means3D, const float* viewmatrix, uint8_t* present)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	V3 p = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
	present[idx] = xform4x3(p, viewmatrix).z > NEAR_Z;
}

// d(normalize(v))/dv applied to dv; auxiliary.h:107-117
__device__ __forceinline__ V3 dnormvdv(V3 v, V3 dv)
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	V3 r;
	r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return r;
}

// One kernel for the whole per-Gaussian chain rule.  Reads the blend-stage gradient record gb[idx]
// (12 floats) and writes every per-Gaussian output exactly once (zeros for culled Gaussians, like the
// reference's zero-initialised tensors, backward.cu:156,367).
__global__ void __launch_bounds__(256) project_bwd_kernel(ProjectBwdArgs a)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= a.P) return;
	const bool live = a.radii[idx] > 0;

	float4 g0 = make_float4(0, 0, 0, 0), g1 = g0, g2 = g0;
	if (live) {
		const float4* gb = reinterpret_cast<const float4*>(a.gb + (size_t)idx * GB_STRIDE);
		g0 = gb[0]; g1 = gb[1]; g2 = gb[2];
	}
	const float dmx = g0.x, dmy = g0.y, dca = g0.z, dcb = g0.w, dcc = g1.x, dop = g1.y;
	float dcol[3] = { g1.z, g1.w, g2.x };
	const float ddepth = g2.y;

	// plain copies of blend-stage gradients into the reference's output tensors
	// accumulate mode (multi-view batches): outputs are SUMMED into the caller's buffers with L2 reductions, so several
	// views -- on different streams -- can write straight into one packed gradient buffer (the all-reduce message)
	const bool acc = a.accumulate != 0;
	if (acc && !live) return;
	auto put = [acc](float* p, float v) { if (acc) red_add(p, v); else *p = v; };
	put(a.dL_dmean2D + 3 * idx + 0, dmx); put(a.dL_dmean2D + 3 * idx + 1, dmy);
	if (!acc) a.dL_dmean2D[3 * idx + 2] = 0.f;
	if (a.dL_dconic) { put(a.dL_dconic + 4 * idx, dca); put(a.dL_dconic + 4 * idx + 1, dcb); if (!acc) a.dL_dconic[4 * idx + 2] = 0.f; put(a.dL_dconic + 4 * idx + 3, dcc); }
	put(a.dL_dopacity + idx, dop);
	if (a.dL_dcolor) { put(a.dL_dcolor + 3 * idx, dcol[0]); put(a.dL_dcolor + 3 * idx + 1, dcol[1]); put(a.dL_dcolor + 3 * idx + 2, dcol[2]); }

	float dmean[3] = { 0.f, 0.f, 0.f };
	float dcov[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
	float dscale[3] = { 0.f, 0.f, 0.f };
	float drot[4] = { 0.f, 0.f, 0.f, 0.f };
	const int M = a.M;
	float* dsh = a.dL_dsh ? a.dL_dsh + (size_t)idx * M * 3 : nullptr;

	if (!live) {
		if (dsh) for (int k = 0; k < 3 * M; k++) dsh[k] = 0.f;
	} else {
		const float* cov3D = (a.cov3D_precomp ? a.cov3D_precomp : a.cov3D) + 6 * (size_t)idx;
		V3 mean = { a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2] };

		// ---- conic -> cov2D -> (cov3D, mean); backward.cu:144-274 ----
		Cov2DTerms ct = cov2d_terms(mean, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, a.viewmatrix);
		const float limx = 1.3f * a.tan_fovx;
		const float limy = 1.3f * a.tan_fovy;
		const float x_grad_mul = ct.txtz < -limx || ct.txtz > limx ? 0 : 1;
		const float y_grad_mul = ct.tytz < -limy || ct.tytz > limy ? 0 : 1;
		const M3& T = ct.T; const M3& Vrk = ct.Vrk; const M3& W = ct.W;
		M3 cov2D = mul(mul(transpose(T), transpose(Vrk)), T);
		float ca = cov2D.m[0][0] += 0.3f;
		float cb = cov2D.m[0][1];
		float cc = cov2D.m[1][1] += 0.3f;
		float denom = ca * cc - cb * cb;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
		if (denom2inv != 0) {
			dL_da = denom2inv * (-cc * cc * dca + 2 * cb * cc * dcb + (denom - ca * cc) * dcc);
			dL_dc = denom2inv * (-ca * ca * dcc + 2 * ca * cb * dcb + (denom - ca * cc) * dca);
			dL_db = denom2inv * 2 * (cb * cc * dca - (denom + 2 * cb * cb) * dcb + ca * cb * dcc);
			dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
			dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
			dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
			dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][1] * dL_dc;
			dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][2] * dL_dc;
			dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db + 2 * T.m[1][1] * T.m[1][2] * dL_dc;
		}
		float dL_dT00 = 2 * (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_da +
			(T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_db;
		float dL_dT01 = 2 * (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_da +
			(T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_db;
		float dL_dT02 = 2 * (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_da +
			(T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_db;
		float dL_dT10 = 2 * (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_dc +
			(T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_db;
		float dL_dT11 = 2 * (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_dc +
			(T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_db;
		float dL_dT12 = 2 * (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_dc +
			(T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_db;
		float dL_dJ00 = W.m[0][0] * dL_dT00 + W.m[0][1] * dL_dT01 + W.m[0][2] * dL_dT02;
		float dL_dJ02 = W.m[2][0] * dL_dT00 + W.m[2][1] * dL_dT01 + W.m[2][2] * dL_dT02;
		float dL_dJ11 = W.m[1][0] * dL_dT10 + W.m[1][1] * dL_dT11 + W.m[1][2] * dL_dT12;
		float dL_dJ12 = W.m[2][0] * dL_dT10 + W.m[2][1] * dL_dT11 + W.m[2][2] * dL_dT12;
		const V3 t = ct.t;
		float tz = 1.f / t.z;
		float tz2 = tz * tz;
		float tz3 = tz2 * tz;
		// clamp convention of the reference: tx,ty gradients masked, dtz uses the clamped t (backward.cu:262-264)
		float dL_dtx = x_grad_mul * -a.focal_x * tz2 * dL_dJ02;
		float dL_dty = y_grad_mul * -a.focal_y * tz2 * dL_dJ12;
		float dL_dtz = -a.focal_x * tz2 * dL_dJ00 - a.focal_y * tz2 * dL_dJ11 + (2 * a.focal_x * t.x) * tz3 * dL_dJ02 + (2 * a.focal_y * t.y) * tz3 * dL_dJ12;
		const float* vm = a.viewmatrix;
		dmean[0] = vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz;
		dmean[1] = vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz;
		dmean[2] = vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz;

		// ---- mean2D -> mean3D through the full projection; backward.cu:371-387 ----
		const float* proj = a.projmatrix;
		float4 m_hom = xform4x4(mean, proj);
		float m_w = 1.0f / (m_hom.w + 0.0000001f);
		float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
		float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
		float dm2[3];
		dm2[0] = (proj[0] * m_w - proj[3] * mul1) * dmx + (proj[1] * m_w - proj[3] * mul2) * dmy;
		dm2[1] = (proj[4] * m_w - proj[7] * mul1) * dmx + (proj[5] * m_w - proj[7] * mul2) * dmy;
		dm2[2] = (proj[8] * m_w - proj[11] * mul1) * dmx + (proj[9] * m_w - proj[11] * mul2) * dmy;
		dmean[0] += dm2[0]; dmean[1] += dm2[1]; dmean[2] += dm2[2];

		// ---- depth channel (extension; the reference renders no depth): depth = (V p).z ----
		dmean[0] += vm[2] * ddepth; dmean[1] += vm[6] * ddepth; dmean[2] += vm[10] * ddepth;

		// ---- colour -> SH (+ view-direction term into the mean); backward.cu:20-139 ----
		if (a.shs) {
			const int deg = a.D;
			float dox = mean.x - a.cam_pos[0], doy = mean.y - a.cam_pos[1], doz = mean.z - a.cam_pos[2];
			float len = sqrtf(dox * dox + doy * doy + doz * doz);
			float x = dox / len, y = doy / len, z = doz / len;
			const float* sh = a.shs + (size_t)idx * M * 3;
			const uint8_t cl = a.clamped[idx];
			float dRGB[3];
#pragma unroll
			for (int c = 0; c < 3; c++) dRGB[c] = dcol[c] * (((cl >> c) & 1) ? 0.f : 1.f);
			float dRGBdx[3] = { 0, 0, 0 }, dRGBdy[3] = { 0, 0, 0 }, dRGBdz[3] = { 0, 0, 0 };
#define SHK(k) sh[3 * (k) + c]
#define DSH(k, v) { float v_ = (v); _Pragma("unroll") for (int c = 0; c < 3; c++) put(&dsh[3 * (k) + c], v_ * dRGB[c]); }
			DSH(0, kSH_C0);
			if (deg > 0) {
				float dRGBdsh1 = -kSH_C1 * y;
				float dRGBdsh2 = kSH_C1 * z;
				float dRGBdsh3 = -kSH_C1 * x;
				DSH(1, dRGBdsh1); DSH(2, dRGBdsh2); DSH(3, dRGBdsh3);
#pragma unroll
				for (int c = 0; c < 3; c++) { dRGBdx[c] = -kSH_C1 * SHK(3); dRGBdy[c] = -kSH_C1 * SHK(1); dRGBdz[c] = kSH_C1 * SHK(2); }
				if (deg > 1) {
					float xx = x * x, yy = y * y, zz = z * z;
					float xy = x * y, yz = y * z, xz = x * z;
					DSH(4, kSH_C2[0] * xy); DSH(5, kSH_C2[1] * yz); DSH(6, kSH_C2[2] * (2.f * zz - xx - yy));
					DSH(7, kSH_C2[3] * xz); DSH(8, kSH_C2[4] * (xx - yy));
#pragma unroll
					for (int c = 0; c < 3; c++) {
						dRGBdx[c] += kSH_C2[0] * y * SHK(4) + kSH_C2[2] * 2.f * -x * SHK(6) + kSH_C2[3] * z * SHK(7) + kSH_C2[4] * 2.f * x * SHK(8);
						dRGBdy[c] += kSH_C2[0] * x * SHK(4) + kSH_C2[1] * z * SHK(5) + kSH_C2[2] * 2.f * -y * SHK(6) + kSH_C2[4] * 2.f * -y * SHK(8);
						dRGBdz[c] += kSH_C2[1] * y * SHK(5) + kSH_C2[2] * 2.f * 2.f * z * SHK(6) + kSH_C2[3] * x * SHK(7);
					}
					if (deg > 2) {
						DSH(9, kSH_C3[0] * y * (3.f * xx - yy)); DSH(10, kSH_C3[1] * xy * z);
						DSH(11, kSH_C3[2] * y * (4.f * zz - xx - yy)); DSH(12, kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
						DSH(13, kSH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, kSH_C3[5] * z * (xx - yy));
						DSH(15, kSH_C3[6] * x * (xx - 3.f * yy));
#pragma unroll
						for (int c = 0; c < 3; c++) {
							dRGBdx[c] += (
								kSH_C3[0] * SHK(9) * 3.f * 2.f * xy +
								kSH_C3[1] * SHK(10) * yz +
								kSH_C3[2] * SHK(11) * -2.f * xy +
								kSH_C3[3] * SHK(12) * -3.f * 2.f * xz +
								kSH_C3[4] * SHK(13) * (-3.f * xx + 4.f * zz - yy) +
								kSH_C3[5] * SHK(14) * 2.f * xz +
								kSH_C3[6] * SHK(15) * 3.f * (xx - yy));
							dRGBdy[c] += (
								kSH_C3[0] * SHK(9) * 3.f * (xx - yy) +
								kSH_C3[1] * SHK(10) * xz +
								kSH_C3[2] * SHK(11) * (-3.f * yy + 4.f * zz - xx) +
								kSH_C3[3] * SHK(12) * -3.f * 2.f * yz +
								kSH_C3[4] * SHK(13) * -2.f * xy +
								kSH_C3[5] * SHK(14) * -2.f * yz +
								kSH_C3[6] * SHK(15) * -3.f * 2.f * xy);
							dRGBdz[c] += (
								kSH_C3[1] * SHK(10) * xy +
								kSH_C3[2] * SHK(11) * 4.f * 2.f * yz +
								kSH_C3[3] * SHK(12) * 3.f * (2.f * zz - xx - yy) +
								kSH_C3[4] * SHK(13) * 4.f * 2.f * xz +
								kSH_C3[5] * SHK(14) * (xx - yy));
						}
					}
				}
			}
			// coefficients above the active degree receive no gradient (the reference leaves its zeros)
			{
				const int used = (deg + 1) * (deg + 1);
				if (!acc) for (int k = used; k < M; k++) { dsh[3 * k] = 0.f; dsh[3 * k + 1] = 0.f; dsh[3 * k + 2] = 0.f; }
			}
#undef SHK
#undef DSH
			V3 dL_ddir = { dot(dRGBdx, dRGB), dot(dRGBdy, dRGB), dot(dRGBdz, dRGB) };
			V3 dmm = dnormvdv(V3{ dox, doy, doz }, dL_ddir);
			dmean[0] += dmm.x; dmean[1] += dmm.y; dmean[2] += dmm.z;
		}

		// ---- cov3D -> scale, quaternion; backward.cu:278-341 ----
		if (a.scales) {
			const float* rot = a.rotations + 4 * (size_t)idx;
			const float* sc = a.scales + 3 * (size_t)idx;
			float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
			M3 R = m3(
				1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
				2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
				2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
			float s[3] = { a.scale_modifier * sc[0], a.scale_modifier * sc[1], a.scale_modifier * sc[2] };
			M3 S = m3(s[0], 0.f, 0.f, 0.f, s[1], 0.f, 0.f, 0.f, s[2]);
			M3 Mx = mul(S, R);
			M3 dL_dSigma = m3(
				dcov[0], 0.5f * dcov[1], 0.5f * dcov[2],
				0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
				0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
			M3 M2;
#pragma unroll
			for (int j = 0; j < 3; j++)
#pragma unroll
				for (int i = 0; i < 3; i++) M2.m[j][i] = 2.0f * Mx.m[j][i];
			M3 dL_dM = mul(M2, dL_dSigma);
			M3 Rt = transpose(R);
			M3 dMt = transpose(dL_dM);
			dscale[0] = dot(Rt.m[0], dMt.m[0]);
			dscale[1] = dot(Rt.m[1], dMt.m[1]);
			dscale[2] = dot(Rt.m[2], dMt.m[2]);
#pragma unroll
			for (int i = 0; i < 3; i++) { dMt.m[0][i] *= s[0]; dMt.m[1][i] *= s[1]; dMt.m[2][i] *= s[2]; }
			drot[0] = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) + 2 * x * (dMt.m[1][2] - dMt.m[2][1]);
			drot[1] = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) + 2 * r * (dMt.m[1][2] - dMt.m[2][1]) - 4 * x * (dMt.m[2][2] + dMt.m[1][1]);
			drot[2] = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * r * (dMt.m[2][0] - dMt.m[0][2]) + 2 * z * (dMt.m[1][2] + dMt.m[2][1]) - 4 * y * (dMt.m[2][2] + dMt.m[0][0]);
			drot[3] = 2 * r * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) + 2 * y * (dMt.m[1][2] + dMt.m[2][1]) - 4 * z * (dMt.m[1][1] + dMt.m[0][0]);
		}
	}

	put(a.dL_dmean3D + 3 * idx, dmean[0]); put(a.dL_dmean3D + 3 * idx + 1, dmean[1]); put(a.dL_dmean3D + 3 * idx + 2, dmean[2]);
	if (a.dL_dcov3D) {
#pragma unroll
		for (int i = 0; i < 6; i++) put(a.dL_dcov3D + 6 * (size_t)idx + i, dcov[i]);
	}
	if (a.dL_dscale) { put(a.dL_dscale + 3 * idx, dscale[0]); put(a.dL_dscale + 3 * idx + 1, dscale[1]); put(a.dL_dscale + 3 * idx + 2, dscale[2]); }
	if (a.dL_drot) {
		float* dr = a.dL_drot + 4 * (size_t)idx;
		if ((reinterpret_cast<uintptr_t>(dr) & 15) != 0) { put(dr, drot[0]); put(dr + 1, drot[1]); put(dr + 2, drot[2]); put(dr + 3, drot[3]); }
		else if (acc) red_add_v4(dr, drot[0], drot[1], drot[2], drot[3]);
		else *reinterpret_cast<float4*>(dr) = make_float4(drot[0], drot[1], drot[2], drot[3]);
	}
	if (a.dL_ddepth) a.dL_ddepth[idx] = ddepth;
}

void launch_project_fwd(const ProjectFwdArgs& a, cudaStream_t s)
{
	if (a.P > 0) project_fwd_kernel<<<ceil_div(a.P, 256), 256, 0, s>>>(a);
}
void launch_project_bwd(const ProjectBwdArgs& a, cudaStream_t s)
{
	if (a.P > 0) project_bwd_kernel<<<ceil_div(a.P, 256), 256, 0, s>>>(a);
}
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s)
{
	if (P > 0) mark_visible_kernel<<<ceil_div(P, 256), 256, 0, s>>>(P, means3D, viewmatrix, present);
}

}  // namespace mgs
