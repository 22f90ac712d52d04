// Staging shared by the forward and backward blend kernels: one batch of a tile's sorted work list is
// brought into shared memory as
//   s_rec[j]  : 32-byte InstRec, ONE contiguous TMA bulk copy for the whole batch (UBLKCP)
//   s_id[j]   : Gaussian id
//   s_ch[j][q]: channel row {r, g, b, depth | f0..f3 | ...}, NQ float4s; the feature part is one TMA
//               bulk copy per row (rows are 16-byte multiples when F % 4 == 0), rgb/depth by plain loads.
// Completion of all bulk copies of a batch is tracked by a single mbarrier (expect_tx = total bytes).
#pragma once
#include "mgs_common.cuh"
#include "mgs_kernels.h"

namespace mgs {

constexpr int BLEND_THREADS = 256;
constexpr int BATCH = 256;  // instances staged per round (one per thread)

// Stage instances [lo, lo+n) of the sorted list.  Must be called by all BLEND_THREADS threads, after a
// __syncthreads() that retired every read of the previous batch.  Returns after the data is visible.
template <int NQ>
__device__ __forceinline__ void stage_batch(const BlendArgs& a, uint32_t lo, int n, InstRec* s_rec, uint32_t* s_id,
	float4* s_ch, uint64_t* bar, uint32_t& phase)
{
	const int t = threadIdx.x;
	const int F = a.F;
	const bool bulk_feat = (F > 0) && ((F & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.feature) & 15) == 0);
	if (t == 0) {
		fence_proxy_async();
		uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(InstRec);
		if (bulk_feat) bytes += (uint32_t)n * (uint32_t)F * 4u;
		mbar_arrive_expect_tx(bar, bytes);
		bulk_g2s(s_rec, a.recs + lo, (uint32_t)n * (uint32_t)sizeof(InstRec), bar);
	}
	if (t < n) {
		const uint32_t id = a.point_list[lo + t];
		s_id[t] = id;
		float4* row = s_ch + (size_t)t * NQ;
		const float* c = a.rgb + 3 * (size_t)id;
		row[0] = make_float4(c[0], c[1], c[2], a.want_depth ? a.depths[id] : 0.f);
		if (NQ > 1) {
			if (bulk_feat) {
				bulk_g2s(row + 1, a.feature + (size_t)id * F, (uint32_t)F * 4u, bar);
				// zero the padding quads beyond F (only when 4 + F is not a multiple covered by NQ)
#pragma unroll
				for (int q = 1; q < NQ; q++)
					if (4 * (q - 1) >= F) row[q] = make_float4(0.f, 0.f, 0.f, 0.f);
			} else {
				float* rf = reinterpret_cast<float*>(row + 1);
				const float* f = a.feature ? a.feature + (size_t)id * F : nullptr;
#pragma unroll
				for (int k = 0; k < 4 * (NQ - 1); k++) rf[k] = (k < F) ? f[k] : 0.f;
			}
		}
	}
	__syncthreads();          // plain stores visible to the CTA
	mbar_wait(bar, phase);    // bulk copies landed
	phase ^= 1;
}

}  // namespace mgs
