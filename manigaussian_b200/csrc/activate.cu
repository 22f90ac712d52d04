// Per-Gaussian parameter activations + deformation offsets, forward and backward, one fused kernel each.
//
// Replaces the chain of elementwise PyTorch ops ManiGaussian runs on the rasterizer's inputs once per step:
//   current frame : xyz + xyz_maps, exp + clamp_max(0.05) on scales, F.normalize on rotations, sigmoid on opacity
//                   (MG/agents/manigaussian_bc/models_embed.py:245-252)
//   next frame    : xyz.detach() + d_xyz, F.normalize(rot.detach() + d_rot), the rest carried over (models_embed.py:297-304)
//   every render  : feature / (||feature|| + 1e-12)      (MG/agents/manigaussian_bc/gaussian_renderer/__init__.py:66-68)
// These are view-independent, so they run once per step (not once per view) and write the activated arrays the
// projection kernel reads; the backward consumes the per-Gaussian gradients summed over all views and chains them
// to the raw parameters and to the offsets in one pass.  HBM-bound streaming kernels: each input/output float moves once.
//
// Arithmetic mirrors the PyTorch operators (IEEE division, expf, sqrtf); reductions over a row run in a different
// order than ATen's, so results agree to an ulp or two, not bit for bit (tests/test_render_gpu.py::test_fused_activations_match_torch_golden_and_oracle states 2e-6).
#include "mgs_common.cuh"
#include "mgs_kernels.h"

namespace mgs {

static constexpr int ACT_THREADS = 128;
static constexpr float ROT_EPS = 1e-12f;      // torch.nn.functional.normalize default eps (clamp_min on the norm)
static constexpr float FEAT_EPS = 1e-12f;     // MIN_DENOMINATOR, gaussian_renderer/__init__.py:67 (added to the norm)

__device__ __forceinline__ float group_sum(float v, int G)
{
	for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// rows of F floats, G = power-of-two lanes per row (32/G rows per warp per trip), coalesced along the row
template <bool BWD>
__device__ __forceinline__ void feature_rows(const ActivateArgs& a)
{
	const int G = a.feature_group, F = a.F;
	const int lane = threadIdx.x & 31;
	const int sub = lane / G, l = lane % G;
	const int rows_per_warp = 32 / G;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long base = warp * rows_per_warp; base < a.P; base += nwarps * rows_per_warp) {
		const long long row = base + sub;
		const bool valid = row < a.P;
		const float* x = a.feature + row * F;
		float ss = 0.f, gx = 0.f;
		if (valid)
			for (int c = l; c < F; c += G) {
				const float v = x[c];
				ss += v * v;
				if (BWD) gx += a.g_feature[row * F + c] * v;
			}
		ss = group_sum(ss, G);
		if (BWD) gx = group_sum(gx, G);
		if (!valid) continue;
		if (!a.feature_normalize) {
			for (int c = l; c < F; c += G) {
				if (BWD) a.dL_feature[row * F + c] = a.g_feature[row * F + c];
				else a.o_feature[row * F + c] = x[c];
			}
			continue;
		}
		const float n = sqrtf(ss), d = n + FEAT_EPS;
		if (!BWD) {
			for (int c = l; c < F; c += G) a.o_feature[row * F + c] = x[c] / d;
		} else {
			// y = x / (n + eps):  dx = g / d - x * (g.x) / (d^2 n); the norm's subgradient at n = 0 is 0 (ATen's norm backward)
			const float k = n > 0.f ? gx / (d * d * n) : 0.f;
			for (int c = l; c < F; c += G) a.dL_feature[row * F + c] = a.g_feature[row * F + c] / d - x[c] * k;
		}
	}
}

// 128-bit path (F a multiple of 4, rows 16-byte aligned, F <= 128): one float4 per lane per row, G4 = power-of-two lanes
// per row, and ACT_UNROLL rows per lane in flight -- the scalar path above keeps one 4-byte load per lane in flight,
// which Little's law caps near a quarter of the HBM peak (measured 27 % forward / 34 % backward at P = 500k, F = 32).
static constexpr int ACT_UNROLL = 4;

template <bool BWD>
__device__ __forceinline__ void feature_rows_vec4(const ActivateArgs& a)
{
	const int G = a.feature_group4, F4 = a.F >> 2;
	const int lane = threadIdx.x & 31;
	const int sub = lane / G, l = lane % G;
	const int rows_per_warp = 32 / G;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	const long long stride = nwarps * rows_per_warp;
	const bool col = l < F4;
	const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
	for (long long base = warp * rows_per_warp; base < a.P; base += stride * ACT_UNROLL) {
		float4 x[ACT_UNROLL], g[ACT_UNROLL];
		long long row[ACT_UNROLL];
		bool ok[ACT_UNROLL];
#pragma unroll
		for (int u = 0; u < ACT_UNROLL; u++) {
			row[u] = base + u * stride + sub;
			ok[u] = col && row[u] < a.P;
			x[u] = ok[u] ? ldg_nc_v4(reinterpret_cast<const float4*>(a.feature + row[u] * a.F) + l) : zero;
			if (BWD) g[u] = ok[u] ? ldg_nc_v4(reinterpret_cast<const float4*>(a.g_feature + row[u] * a.F) + l) : zero;
		}
#pragma unroll
		for (int u = 0; u < ACT_UNROLL; u++) {
			float ss = x[u].x * x[u].x + x[u].y * x[u].y + x[u].z * x[u].z + x[u].w * x[u].w;
			float gx = 0.f;
			if (BWD) gx = g[u].x * x[u].x + g[u].y * x[u].y + g[u].z * x[u].z + g[u].w * x[u].w;
			ss = group_sum(ss, G);
			if (BWD) gx = group_sum(gx, G);
			if (!ok[u]) continue;
			float4 o;
			if (!a.feature_normalize) {
				o = BWD ? g[u] : x[u];
			} else {
				const float n = sqrtf(ss), d = n + FEAT_EPS;
				if (!BWD) {
					o = make_float4(x[u].x / d, x[u].y / d, x[u].z / d, x[u].w / d);
				} else {
					const float k = n > 0.f ? gx / (d * d * n) : 0.f;
					o = make_float4(g[u].x / d - x[u].x * k, g[u].y / d - x[u].y * k, g[u].z / d - x[u].z * k, g[u].w / d - x[u].w * k);
				}
			}
			reinterpret_cast<float4*>((BWD ? a.dL_feature : a.o_feature) + row[u] * a.F)[l] = o;
		}
	}
}

__global__ void __launch_bounds__(ACT_THREADS) activate_fwd_kernel(ActivateArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.P && blockIdx.x < a.small_blocks) {
		if (a.o_means) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				float v = a.means[3 * i + k];
				if (a.d_means) v += a.d_means[3 * i + k];
				a.o_means[3 * i + k] = v;
			}
		}
		if (a.o_scales) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				float v = a.scales[3 * i + k];
				if (a.d_scales) v += a.d_scales[3 * i + k];
				if (a.scale_mode == 1) v = fminf(expf(v), a.scale_max);
				a.o_scales[3 * i + k] = v;
			}
		}
		if (a.o_rot) {
			float q[4];
			float ss = 0.f;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				q[k] = a.rot[4 * i + k];
				if (a.d_rot) q[k] += a.d_rot[4 * i + k];
				ss += q[k] * q[k];
			}
			const float d = a.rot_normalize ? fmaxf(sqrtf(ss), ROT_EPS) : 1.f;
#pragma unroll
			for (int k = 0; k < 4; k++) a.o_rot[4 * i + k] = a.rot_normalize ? q[k] / d : q[k];
		}
		if (a.o_opac) {
			const float v = a.opac[i];
			a.o_opac[i] = a.opacity_mode == 1 ? 1.0f / (1.0f + expf(-v)) : v;
		}
	}
	if (a.o_feature) { if (a.feature_vec4) feature_rows_vec4<false>(a); else feature_rows<false>(a); }
}

__global__ void __launch_bounds__(ACT_THREADS) activate_bwd_kernel(ActivateArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.P && blockIdx.x < a.small_blocks) {
		if (a.g_means) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				const float g = a.g_means[3 * i + k];
				if (a.dL_means) a.dL_means[3 * i + k] = g;
				if (a.dL_dmeans) a.dL_dmeans[3 * i + k] = g;
			}
		}
		if (a.g_scales) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				float g = a.g_scales[3 * i + k];
				if (a.scale_mode == 1) {
					float v = a.scales[3 * i + k];
					if (a.d_scales) v += a.d_scales[3 * i + k];
					const float e = expf(v);
					g = e <= a.scale_max ? g * e : 0.f;  // clamp_max passes the gradient where input <= max, exp' = exp
				}
				if (a.dL_scales) a.dL_scales[3 * i + k] = g;
				if (a.dL_dscales) a.dL_dscales[3 * i + k] = g;
			}
		}
		if (a.g_rot) {
			float q[4], g[4];
			float ss = 0.f, gq = 0.f;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				q[k] = a.rot[4 * i + k];
				if (a.d_rot) q[k] += a.d_rot[4 * i + k];
				g[k] = a.g_rot[4 * i + k];
				ss += q[k] * q[k];
				gq += g[k] * q[k];
			}
			if (a.rot_normalize) {
				// y = q / max(n, eps): dq = g / d - q * (g.q) / (d^2 n) where the clamp passes (n >= eps), else g / d
				const float n = sqrtf(ss), d = fmaxf(n, ROT_EPS);
				const float k2 = (n >= ROT_EPS && n > 0.f) ? gq / (d * d * n) : 0.f;
#pragma unroll
				for (int k = 0; k < 4; k++) g[k] = g[k] / d - q[k] * k2;
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				if (a.dL_rot) a.dL_rot[4 * i + k] = g[k];
				if (a.dL_drot) a.dL_drot[4 * i + k] = g[k];
			}
		}
		if (a.g_opac && a.dL_opac) {
			float g = a.g_opac[i];
			if (a.opacity_mode == 1) {
				const float y = 1.0f / (1.0f + expf(-a.opac[i]));
				g = (g * (1.0f - y)) * y;  // ATen sigmoid_backward: grad * (1 - y) * y
			}
			a.dL_opac[i] = g;
		}
	}
	if (a.g_feature && a.dL_feature) { if (a.feature_vec4) feature_rows_vec4<true>(a); else feature_rows<true>(a); }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static void grid_for(ActivateArgs& a, bool features, bool bwd, int& grid)
{
	a.small_blocks = (a.P + ACT_THREADS - 1) / ACT_THREADS;
	int g = 1;
	while (g < a.F && g < 32) g <<= 1;
	a.feature_group = g;
	int g4 = 1;
	while (g4 < (a.F >> 2) && g4 < 32) g4 <<= 1;
	a.feature_group4 = g4;
	a.feature_vec4 = features && a.F > 0 && (a.F & 3) == 0 && a.F <= 128 && aligned16(a.feature) &&
		(bwd ? (aligned16(a.g_feature) && aligned16(a.dL_feature)) : aligned16(a.o_feature));
	grid = a.small_blocks;
	if (features) {
		// feature rows: a grid-stride loop over enough warps to fill the SMs a few times
		const long long rows_per_block = (long long)(ACT_THREADS / 32) * (32 / (a.feature_vec4 ? g4 : g)) * (a.feature_vec4 ? ACT_UNROLL : 1);
		const long long need = ((long long)a.P + rows_per_block - 1) / rows_per_block;
		const long long cap = 148LL * 16 * 4;
		const long long fb = need < cap ? need : cap;
		if (fb > grid) grid = (int)fb;
	}
}

void launch_activate_fwd(ActivateArgs a, cudaStream_t s)
{
	int grid;
	grid_for(a, a.o_feature != nullptr, false, grid);
	if (a.P > 0) activate_fwd_kernel<<<grid, ACT_THREADS, 0, s>>>(a);
}

void launch_activate_bwd(ActivateArgs a, cudaStream_t s)
{
	int grid;
	grid_for(a, a.g_feature != nullptr && a.dL_feature != nullptr, true, grid);
	if (a.P > 0) activate_bwd_kernel<<<grid, ACT_THREADS, 0, s>>>(a);
}

}  // namespace mgs
