// Forward blend: front-to-back alpha compositing of RGB + depth + F feature channels per tile.
//
// Replaces FORWARD::render / renderCUDA<3,F> (DGR/cuda_rasterizer/forward.cu:262-398) behind the C-ABI.
// Same per-pixel semantics (power > 0 skip, alpha = min(0.99, o*exp(power)), alpha < 1/255 skip, stop
// when T*(1-alpha) < 1e-4, colour gets + T*bg, features/depth do not; final_T and n_contrib saved).
//
// B200 design (not the reference's):
//  * one CTA per 16x16 tile (tile ids must match the reference), 8 warps; each warp owns an 8x4 pixel
//    block and culls the tile's work list against that block with the per-Gaussian alpha >= 1/255
//    footprint (exact-conservative), so a pixel only evaluates Gaussians that can reach its block;
//  * per-instance records arrive in tile order by ONE TMA bulk copy per batch, feature rows by one
//    bulk copy per row (blend_common.cuh); channel rows are read from shared memory as 128-bit
//    broadcasts instead of per-pair scalar global gathers (forward.cu:364-371);
//  * feature width is a run-time value dispatched to NQ = ceil((4+F)/4) in {1,2,3,5,9}.
#include "blend_common.cuh"

namespace mgs {

template <int NQ>
__global__ void __launch_bounds__(BLEND_THREADS) blend_fwd_kernel(BlendArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	InstRec* s_rec = reinterpret_cast<InstRec*>(smem_raw);
	float4* s_ch = reinterpret_cast<float4*>(smem_raw + BATCH * sizeof(InstRec));
	uint32_t* s_id = reinterpret_cast<uint32_t*>(smem_raw + BATCH * sizeof(InstRec) + (size_t)BATCH * NQ * sizeof(float4));
	__shared__ __align__(8) uint64_t bar;

	const int tile = blockIdx.x;
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int bx0 = tile_x * TILE_X + (warp & 1) * WARP_BX;
	const int by0 = tile_y * TILE_Y + (warp >> 1) * WARP_BY;
	const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
	const bool inside = pxi < a.W && pyi < a.H;
	const float pfx = (float)pxi, pfy = (float)pyi;
	const float fbx0 = (float)bx0, fbx1 = (float)(bx0 + WARP_BX - 1), fby0 = (float)by0, fby1 = (float)(by0 + WARP_BY - 1);

	if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
	uint32_t phase = 0;

	const uint2 range = a.ranges[tile];
	float T = 1.0f;
	uint32_t last_contributor = 0;
	bool done = !inside;
	float acc[4 * NQ];
#pragma unroll
	for (int i = 0; i < 4 * NQ; i++) acc[i] = 0.f;
	bool warp_done = __all_sync(0xffffffffu, done);
	const float4* s_rec4 = reinterpret_cast<const float4*>(s_rec);

	for (uint32_t lo = range.x; lo < range.y; lo += BATCH) {
		// retire the previous batch; stop once every warp of the tile is finished
		if (!__syncthreads_or(!warp_done)) break;
		const int n = min((int)BATCH, (int)(range.y - lo));
		stage_batch<NQ>(a, lo, n, s_rec, s_id, s_ch, &bar, phase);
		if (warp_done) continue;

		for (int c = 0; c < n; c += 32) {
			const int j = c + lane;
			bool hit = false;
			if (j < n) {
				const float4 r0 = s_rec4[2 * j], r1 = s_rec4[2 * j + 1];
				hit = (r1.z >= 0.f) && (r0.x + r1.z >= fbx0) && (r0.x - r1.z <= fbx1) && (r0.y + r1.w >= fby0) && (r0.y - r1.w <= fby1);
			}
			uint32_t mask = __ballot_sync(0xffffffffu, hit);
			while (mask) {
				const int jj = c + __ffs(mask) - 1;
				mask &= mask - 1;
				const float4 r0 = s_rec4[2 * jj], r1 = s_rec4[2 * jj + 1];  // {x, y, ca, cb}, {cc, op, hx, hy}
				const float dx = r0.x - pfx, dy = r0.y - pfy;
				const float power = -0.5f * (r0.z * dx * dx + r1.x * dy * dy) - r0.w * dx * dy;
				if (done || power > 0.0f) continue;
				const float alpha = min(ALPHA_MAX, r1.y * expf(power));
				if (alpha < ALPHA_MIN) continue;
				const float test_T = T * (1 - alpha);
				if (test_T < T_STOP) { done = true; continue; }
				const float w = alpha * T;
				const float4* row = s_ch + (size_t)jj * NQ;
#pragma unroll
				for (int q = 0; q < NQ; q++) {
					const float4 v = row[q];
					acc[4 * q + 0] += v.x * w; acc[4 * q + 1] += v.y * w;
					acc[4 * q + 2] += v.z * w; acc[4 * q + 3] += v.w * w;
				}
				T = test_T;
				last_contributor = (lo - range.x) + (uint32_t)jj + 1u;  // 1-based position in the tile list
			}
			if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
		}
	}

	if (inside) {
		const size_t HW = (size_t)a.H * a.W;
		const size_t pix = (size_t)a.W * pyi + pxi;
		a.final_T[pix] = T;
		a.n_contrib[pix] = last_contributor;
#pragma unroll
		for (int ch = 0; ch < 3; ch++) a.out_color[ch * HW + pix] = acc[ch] + T * a.bg[ch];
		if (a.out_depth) a.out_depth[pix] = acc[3];
		if (NQ > 1) {
#pragma unroll
			for (int k = 0; k < 4 * (NQ - 1); k++)
				if (k < a.F) a.out_feature[(size_t)k * HW + pix] = acc[4 + k];
		}
	}
}

static size_t fwd_smem_bytes(int nq) { return BATCH * sizeof(InstRec) + (size_t)BATCH * nq * sizeof(float4) + BATCH * sizeof(uint32_t); }

template <int NQ>
static void launch_fwd_t(const BlendArgs& a, cudaStream_t s)
{
	const size_t smem = fwd_smem_bytes(NQ);
	static bool configured = false;
	if (!configured) {
		cudaFuncSetAttribute(blend_fwd_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = true;
	}
	blend_fwd_kernel<NQ><<<a.grid_x * a.grid_y, BLEND_THREADS, smem, s>>>(a);
}

int blend_supported(int F) { return F >= 0 && F <= 32; }

int nq_for(int F)
{
	const int need = (4 + F + 3) / 4;
	if (need <= 1) return 1;
	if (need <= 2) return 2;
	if (need <= 3) return 3;
	if (need <= 5) return 5;
	return 9;
}

void launch_blend_fwd(const BlendArgs& a, cudaStream_t s)
{
	switch (a.nq) {
	case 1: launch_fwd_t<1>(a, s); break;
	case 2: launch_fwd_t<2>(a, s); break;
	case 3: launch_fwd_t<3>(a, s); break;
	case 5: launch_fwd_t<5>(a, s); break;
	default: launch_fwd_t<9>(a, s); break;
	}
}

}  // namespace mgs
