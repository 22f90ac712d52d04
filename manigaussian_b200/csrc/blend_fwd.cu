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
//  * software pipeline over batches of FB instances, all data movement asynchronous (TMA, SASS UBLKCP):
//      records of batch k+2 : ONE bulk copy of the contiguous tile-ordered 32-byte records (3-deep ring),
//      channel rows of k+1  : per instance one 16-byte bulk copy ({r,g,b,depth}) and one F*4-byte bulk copy
//                             (feature row), gathered by Gaussian id into a double-buffered row array,
//    each tracked by an mbarrier (expect_tx = bytes), so batch k is blended while k+1/k+2 are in flight and
//    the only per-batch synchronisation is one CTA barrier;
//  * channel rows are consumed from shared memory as 128-bit broadcasts instead of per-pair scalar
//    global gathers (forward.cu:364-371);
//  * feature width is a run-time value dispatched to NQ = ceil((4+F)/4) in {1,2,3,5,9}.
#include "blend_common.cuh"

namespace mgs {

constexpr int FB = 256;  // forward batch (instances per pipeline stage)

template <int NQ, bool VEC>
__global__ void __launch_bounds__(BLEND_THREADS) blend_fwd_kernel(BlendArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	InstRec* s_rec = reinterpret_cast<InstRec*>(smem_raw);                                   // RING x FB records
	float4* s_ch = reinterpret_cast<float4*>(smem_raw + (size_t)RING * FB * sizeof(InstRec));  // 2 x FB x NQ quads
	__shared__ __align__(8) uint64_t s_bar_rec[RING];
	__shared__ __align__(8) uint64_t s_bar_ch[2];

	const int tile = blockIdx.x;
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int bx0 = tile_x * TILE_X + (warp & 1) * WARP_BX;
	const int by0 = tile_y * TILE_Y + (warp >> 1) * WARP_BY;
	const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
	const bool inside = pxi < a.W && pyi < a.H;
	const float pfx = (float)pxi, pfy = (float)pyi;
	const float fbx0 = (float)bx0, fbx1 = (float)(bx0 + WARP_BX - 1), fby0 = (float)by0, fby1 = (float)(by0 + WARP_BY - 1);
	const int F = a.F;

	const uint2 range = a.ranges[tile];
	const int total = (int)(range.y - range.x);
	const int nb = (total + FB - 1) / FB;
	auto batch_n = [&](int k) { return min(FB, total - k * FB); };

	if (threadIdx.x == 0) {
#pragma unroll
		for (int i = 0; i < RING; i++) mbar_init(&s_bar_rec[i], 1);
		mbar_init(&s_bar_ch[0], 1);
		mbar_init(&s_bar_ch[1], 1);
		mbar_fence_init();
	}
	__syncthreads();

	auto issue_recs = [&](int k) {  // thread 0
		const int b = k % RING, n = batch_n(k);
		fence_proxy_async();
		mbar_arrive_expect_tx(&s_bar_rec[b], (uint32_t)n * (uint32_t)sizeof(InstRec));
		bulk_g2s(s_rec + b * FB, a.recs + range.x + (size_t)k * FB, (uint32_t)n * (uint32_t)sizeof(InstRec), &s_bar_rec[b]);
	};
	auto wait_recs = [&](int k) -> const float4* {
		const int b = k % RING;
		mbar_wait(&s_bar_rec[b], (uint32_t)((k / RING) & 1));
		return reinterpret_cast<const float4*>(s_rec + b * FB);
	};
	// gather the channel rows of batch k (whose records have landed) into row buffer k & 1
	auto issue_rows = [&](int k, const float4* rec4) {
		const int n = batch_n(k), t = threadIdx.x;
		float4* rows = s_ch + (size_t)(k & 1) * FB * NQ;
		if (VEC) {
			if (t == 0) mbar_arrive_expect_tx(&s_bar_ch[k & 1], (uint32_t)n * (16u + (NQ > 1 ? (uint32_t)F * 4u : 0u)));
			if (t < n) {
				const uint32_t id = rec_id(rec4[2 * t + 1]);
				fence_proxy_async();
				bulk_g2s(rows + (size_t)t * NQ, a.rgbd + id, 16u, &s_bar_ch[k & 1]);
				if (NQ > 1) bulk_g2s(rows + (size_t)t * NQ + 1, a.feature + (size_t)id * F, (uint32_t)F * 4u, &s_bar_ch[k & 1]);
			}
		} else if (t < n) {  // rows that are not 16-byte multiples (e.g. F = 3): plain loads, made visible by a CTA barrier
			const uint32_t id = rec_id(rec4[2 * t + 1]);
			rows[(size_t)t * NQ] = a.rgbd[id];
			float* rf = reinterpret_cast<float*>(rows + (size_t)t * NQ + 1);
			const float* f = a.feature + (size_t)id * F;
#pragma unroll
			for (int i = 0; i < 4 * (NQ - 1); i++) rf[i] = (i < F) ? f[i] : 0.f;
		}
	};
	auto wait_rows = [&](int k) {
		if (VEC) mbar_wait(&s_bar_ch[k & 1], (uint32_t)((k >> 1) & 1));
	};

	if (nb > 0) {
		if (threadIdx.x == 0) {
			issue_recs(0);
			if (nb > 1) issue_recs(1);
		}
		const float4* r0 = wait_recs(0);
		if (VEC) issue_rows(0, r0);
	}

	float T = 1.0f;
	uint32_t last_contributor = 0;
	bool done = !inside;
	float acc[4 * NQ];
#pragma unroll
	for (int i = 0; i < 4 * NQ; i++) acc[i] = 0.f;
	bool warp_done = __all_sync(0xffffffffu, done);

	for (int k = 0; k < nb; k++) {
		// retires batch k-1 (record buffer (k+2)%3 and row buffer (k+1)&1 become free) and votes on early termination
		const int alive = __syncthreads_or(!warp_done);
		if (!alive) {
			// drain what is still in flight before the CTA may exit
			if (k + 1 < nb) wait_recs(k + 1);
			if (!VEC && k > 0) wait_recs(k);
			wait_rows(k);
			break;
		}
		if (threadIdx.x == 0 && k + 2 < nb) issue_recs(k + 2);
		const float4* rec4 = reinterpret_cast<const float4*>(s_rec + (k % RING) * FB);
		if (VEC) {
			if (k + 1 < nb) issue_rows(k + 1, wait_recs(k + 1));
			wait_rows(k);
		} else {
			if (k > 0) rec4 = wait_recs(k);
			issue_rows(k, rec4);
			__syncthreads();
		}
		if (warp_done) continue;
		const int n = batch_n(k);
		const float4* rows = s_ch + (size_t)(k & 1) * FB * NQ;
		const uint32_t pos0 = (uint32_t)(k * FB) + 1u;  // 1-based position of the batch's first record in the tile list

		for (int c = 0; c < n; c += 32) {
			const int j = c + lane;
			bool hit = false;
			if (j < n) hit = rec_hits_block(rec4[2 * j], rec4[2 * j + 1], fbx0, fbx1, fby0, fby1);
			uint32_t mask = __ballot_sync(0xffffffffu, hit);
			while (mask) {
				const int jj = c + __ffs(mask) - 1;
				mask &= mask - 1;
				const float4 r0 = rec4[2 * jj], r1 = rec4[2 * jj + 1];  // {x, y, ca, cb}, {cc, op, ext, id}
				const float dx = r0.x - pfx, dy = r0.y - pfy;
				const float power = -0.5f * (r0.z * dx * dx + r1.x * dy * dy) - r0.w * dx * dy;
				if (done || power > 0.0f) continue;
				const float alpha = min(ALPHA_MAX, r1.y * expf(power));
				if (alpha < ALPHA_MIN) continue;
				const float test_T = T * (1 - alpha);
				if (test_T < T_STOP) { done = true; continue; }
				const float w = alpha * T;
				const float4* row = rows + (size_t)jj * NQ;
#pragma unroll
				for (int q = 0; q < NQ; q++) {
					const float4 v = row[q];
					acc[4 * q + 0] += v.x * w; acc[4 * q + 1] += v.y * w;
					acc[4 * q + 2] += v.z * w; acc[4 * q + 3] += v.w * w;
				}
				T = test_T;
				last_contributor = pos0 + (uint32_t)jj;
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
			for (int i = 0; i < 4 * (NQ - 1); i++)
				if (i < F) a.out_feature[(size_t)i * HW + pix] = acc[4 + i];
		}
	}
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

// 128-bit / bulk feature copies need 16-byte aligned rows of a 16-byte multiple
bool feature_rows_vectorizable(const float* feature, int F)
{
	return F > 0 && (F & 3) == 0 && (reinterpret_cast<uintptr_t>(feature) & 15) == 0;
}

template <int NQ, bool VEC>
static void launch_fwd_tv(const BlendArgs& a, cudaStream_t s)
{
	const size_t smem = (size_t)RING * FB * sizeof(InstRec) + (size_t)2 * FB * NQ * sizeof(float4);
	static bool configured = false;
	if (!configured) {
		cudaFuncSetAttribute(blend_fwd_kernel<NQ, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = true;
	}
	blend_fwd_kernel<NQ, VEC><<<a.grid_x * a.grid_y, BLEND_THREADS, smem, s>>>(a);
}

template <int NQ>
static void launch_fwd_t(const BlendArgs& a, cudaStream_t s)
{
	// a feature row shorter than its padded NQ-1 quads would leave stale shared memory in the tail quads: only the
	// exact fits take the bulk-copy path
	const bool vec = (NQ == 1) || (feature_rows_vectorizable(a.feature, a.F) && a.F == 4 * (NQ - 1));
	if (vec) launch_fwd_tv<NQ, true>(a, s);
	else launch_fwd_tv<NQ, false>(a, s);
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
