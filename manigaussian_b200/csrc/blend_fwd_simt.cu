// Forward blend: front-to-back alpha compositing of RGB + depth + F feature channels.
//
// Replaces FORWARD::render / renderCUDA<3,F> (DGR/cuda_rasterizer/forward.cu:262-398) behind the C-ABI.
// Same per-pixel semantics (power > 0 skip, alpha = min(0.99, o*exp(power)), alpha < 1/255 skip, stop
// when T*(1-alpha) < 1e-4, colour gets + T*bg, features/depth do not; final_T and n_contrib saved).
//
// B200 design (not the reference's; see blend_common.cuh for the decomposition):
//  * one single-warp CTA per 8x4 pixel block; the warp culls its tile's work list against the block with the
//    per-Gaussian alpha >= 1/255 footprint (exact-conservative; decided by ballot), so a pixel only evaluates
//    Gaussians that can reach its block (about a third of the tile's list on the benchmark workload);
//  * two-level software pipeline, all data movement asynchronous:
//      records     : RING 64-record batches in flight/resident, one contiguous 1-D TMA bulk copy each (UBLKCP + mbarrier);
//      channel rows: while chunk g is blended, the {r,g,b,depth} quad and the F-float feature row of every SURVIVOR of
//                    chunk g+1 are gathered by Gaussian id into a double-buffered row array with 16-byte cp.async
//                    (LDGSTS) pieces spread over the lanes, one commit group per chunk.  (A bulk copy per survivor was
//                    the first design; UBLKCP takes uniform operands, so the compiler serialised it into a loop trip per
//                    survivor and copy -- 15 % of the kernel's stall samples.  LDGSTS: 0.426 -> 0.374 ms per c3 view.)
//  * channel rows are consumed from shared memory as 128-bit broadcasts instead of per-pair scalar global
//    gathers (forward.cu:364-371);
//  * feature width is a run-time value dispatched to NQ = ceil((4+F)/4) in {1,2,3,5,9};
//  * two survivors are evaluated per step of the walk: their footprint evaluations are independent of the transmittance
//    recurrence, so the second one's expf latency overlaps with the first one's channel FMAs.
// Tried and rejected on B200 (all bit-identical; numbers per c3 view, this kernel 0.427 ms): "sub-block lockstep" (lanes of
// the two 4x4 halves walk their own survivor lists in lockstep: 25 % fewer steps, 0.443 ms -- the per-lane bookkeeping
// costs what the steps save) and a two-phase walk (phase A evaluates alpha for all survivors into shared memory, phase B
// lets every lane blend only its own contributors: ncu shows 7 of 32 threads active in the channel FMAs here, yet
// 0.429 ms -- the walk is bound by per-warp dependent-issue latency with 2-3 warps per scheduler, not by instruction
// count).
#include "blend_common.cuh"

#ifndef MGS_FWD_PREDICATED
#define MGS_FWD_PREDICATED 0
#endif

namespace mgs {
static_assert(REC_BATCH == 64, "the round-1 SIMT blends assume 64-record batches");

template <int NQ, bool VEC>
__global__ void __launch_bounds__(32, 16) blend_fwd_simt_kernel(BlendArgs a)
{
	__shared__ __align__(128) InstRec s_rec[RING * REC_BATCH];
	__shared__ __align__(16) float4 s_rows[2][32 * NQ];
	__shared__ __align__(8) uint64_t s_bar_rec[RING];
	__shared__ uint32_t s_ids[32];  // Gaussian ids of the chunk's survivors by rank (row gather addressing)

	const int lane = threadIdx.x;
	const int tile = blockIdx.x >> 3, sub = blockIdx.x & 7;
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int bx0 = tile_x * TILE_X + (sub & 1) * WARP_BX;
	const int by0 = tile_y * TILE_Y + (sub >> 1) * WARP_BY;
	const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
	const bool inside = pxi < a.W && pyi < a.H;
	if (__all_sync(0xffffffffu, !inside)) return;  // block entirely outside the image (ragged right/bottom tiles)
	const float pfx = (float)pxi, pfy = (float)pyi;
	const float fbx0 = (float)bx0, fbx1 = (float)(bx0 + WARP_BX - 1), fby0 = (float)by0, fby1 = (float)(by0 + WARP_BY - 1);
	const int F = a.F;

	const uint2 range = a.ranges[tile];
	WarpRecRing ring;
	ring.init(s_rec, s_bar_rec, a.recs + range.x, (int)(range.y - range.x), false);
	const int nb = ring.num_batches();
	const int nchunks = (ring.total + 31) >> 5;
	int issued = 0, waited = 0;      // record batches

	float T = 1.0f;
	uint32_t last_contributor = 0;
	bool done = !inside;
	float acc[4 * NQ];
#pragma unroll
	for (int i = 0; i < 4 * NQ; i++) acc[i] = 0.f;

	// cull chunk g and start gathering the survivors' channel rows into s_rows[g & 1]
	auto prefetch_chunk = [&](int g) -> uint32_t {
		const int k = g >> 1;
		if ((g & 1) == 0) { ring.wait(k); waited = k + 1; }
		__syncwarp();
		const float4* rec4 = ring.buffer(k);
		const int j = ((g & 1) << 5) + lane;
		const bool hit = (g * 32 + lane < ring.total) && rec_hits_block(rec4[2 * j], rec4[2 * j + 1], fbx0, fbx1, fby0, fby1);
		const uint32_t mask = __ballot_sync(0xffffffffu, hit);
		if (mask) {
			float4* rows = s_rows[g & 1];
			const int rank = __popc(mask & ((1u << lane) - 1u));
			if (VEC) {
				// cooperative gather: the rows of all survivors are cut into 16-byte pieces and every lane copies pieces
				// lane, lane+32, ... with per-lane addresses (LDGSTS).  A per-survivor bulk copy (UBLKCP) takes uniform
				// operands, so the compiler serialises it into one loop trip per survivor and copy.
				if (hit) s_ids[rank] = rec_id(rec4[2 * j + 1]);
				__syncwarp();
				const int npieces = __popc(mask) * NQ;
				for (int idx = lane; idx < npieces; idx += 32) {
					const int r = idx / NQ, q = idx - r * NQ;
					const uint32_t id = s_ids[r];
					const float4* src = (q == 0) ? (a.rgbd + id) : (reinterpret_cast<const float4*>(a.feature + (size_t)id * F) + (q - 1));
					cp_async16(rows + idx, src);
				}
			} else if (hit) {  // rows that are not 16-byte multiples (e.g. F = 3): plain loads
				const uint32_t id = rec_id(rec4[2 * j + 1]);
				rows[rank * NQ] = a.rgbd[id];
				if (NQ > 1) {
					float* rf = reinterpret_cast<float*>(rows + rank * NQ + 1);
					const float* f = a.feature + (size_t)id * F;
#pragma unroll
					for (int i = 0; i < 4 * (NQ - 1); i++) rf[i] = (i < F) ? f[i] : 0.f;
				}
			}
		}
		if (VEC) cp_async_commit();  // one group per prefetched chunk, empty or not: wait_group counts stay uniform
		return mask;
	};

	uint32_t mask_cur = 0, mask_next = 0;
	if (nchunks > 0) {
		for (; issued < min(nb, RING); issued++) ring.issue(issued);
		mask_cur = prefetch_chunk(0);
	}
	for (int g = 0; g < nchunks; g++) {
		mask_next = 0;
		if (g + 1 < nchunks) mask_next = prefetch_chunk(g + 1);
		if (mask_cur) {
			const int rb = g & 1;
			if (VEC) {
				// groups complete in order: all but the newest one (chunk g+1, if it was prefetched) must have landed
				if (g + 1 < nchunks) cp_async_wait<1>(); else cp_async_wait<0>();
			}
			__syncwarp();
			const float4* rec4 = ring.buffer(g >> 1) + (rb << 6);
			const float4* rows = s_rows[rb];
			const uint32_t pos0 = (uint32_t)(g * 32) + 1u;  // 1-based position of the chunk's first record in the tile list
			// two survivors per step (see the header)
#if MGS_FWD_PREDICATED
			// EXPERIMENT (off; not yet measured on a GPU): the SASS view of the round-1 capture puts ~25 % of this kernel's
			// stall samples on the divergent skip tests (BRA wait + branch_resolving after each BSYNC).  Here the per-lane
			// decisions become selects, the channel FMAs run for the whole warp behind ONE warp-uniform branch with w = 0 on
			// lanes that do not contribute (fma(v, 0, acc) == acc for finite v), and the two survivors of a step can
			// interleave freely.  Differs from the committed walk only for non-finite channel values.
			auto blend_one = [&](bool act, int b, float power, float alpha_raw, const float4* row) {
				const float alpha = min(ALPHA_MAX, alpha_raw);
				const bool cand = act && !done && !(power > 0.0f) && !(alpha < ALPHA_MIN);
				const float test_T = T * (1 - alpha);
				const bool stop = cand && (test_T < T_STOP);
				const bool use = cand && !stop;
				done = done || stop;
				if (__any_sync(0xffffffffu, use)) {
					const float w = use ? alpha * T : 0.f;
#pragma unroll
					for (int q = 0; q < NQ; q++) {
						const float4 v = row[q];
						acc[4 * q + 0] += v.x * w; acc[4 * q + 1] += v.y * w;
						acc[4 * q + 2] += v.z * w; acc[4 * q + 3] += v.w * w;
					}
				}
				T = use ? test_T : T;
				last_contributor = use ? pos0 + (uint32_t)b : last_contributor;
			};
#else
			auto blend_one = [&](int b, float power, float alpha_raw, const float4* row) {
				if (done || power > 0.0f) return;
				const float alpha = min(ALPHA_MAX, alpha_raw);
				if (alpha < ALPHA_MIN) return;
				const float test_T = T * (1 - alpha);
				if (test_T < T_STOP) { done = true; return; }
				const float w = alpha * T;
#pragma unroll
				for (int q = 0; q < NQ; q++) {
					const float4 v = row[q];
					acc[4 * q + 0] += v.x * w; acc[4 * q + 1] += v.y * w;
					acc[4 * q + 2] += v.z * w; acc[4 * q + 3] += v.w * w;
				}
				T = test_T;
				last_contributor = pos0 + (uint32_t)b;
			};
#endif
			int i = 0;
			uint32_t mask = mask_cur;
			while (mask) {
				const int b0 = __ffs(mask) - 1;
				mask &= mask - 1;
				const bool two = mask != 0;
				const int b1 = two ? __ffs(mask) - 1 : b0;
				if (two) mask &= mask - 1;
				const float4 p0 = rec4[2 * b0], q0 = rec4[2 * b0 + 1];  // {x, y, ca, cb}, {cc, op, ext, id}
				const float4 p1 = rec4[2 * b1], q1 = rec4[2 * b1 + 1];
				const float dx0 = p0.x - pfx, dy0 = p0.y - pfy, dx1 = p1.x - pfx, dy1 = p1.y - pfy;
				const float power0 = -0.5f * (p0.z * dx0 * dx0 + q0.x * dy0 * dy0) - p0.w * dx0 * dy0;
				const float power1 = -0.5f * (p1.z * dx1 * dx1 + q1.x * dy1 * dy1) - p1.w * dx1 * dy1;
				const float a0 = q0.y * expf(power0);
				const float a1 = q1.y * expf(power1);
#if MGS_FWD_PREDICATED
				blend_one(true, b0, power0, a0, rows + i * NQ);
				blend_one(two, b1, power1, a1, rows + (i + 1) * NQ);
#else
				blend_one(b0, power0, a0, rows + i * NQ);
				if (two) blend_one(b1, power1, a1, rows + (i + 1) * NQ);
#endif
				i += two ? 2 : 1;
			}
		}
		if (__all_sync(0xffffffffu, done)) {
			// drain whatever is still in flight before the CTA exits
			if (VEC) cp_async_wait<0>();
			for (int k = waited; k < issued; k++) ring.wait(k);
			break;
		}
		// the batch whose last chunk was just blended frees its buffer for the batch RING ahead
		if ((g & 1) && issued < nb) { ring.issue(issued); issued++; }
		mask_cur = mask_next;
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


int nq_for(int F)
{
	const int need = (4 + F + 3) / 4;
	if (need <= 1) return 1;
	if (need <= 2) return 2;
	if (need <= 3) return 3;
	if (need <= 5) return 5;
	return 9;
}

bool feature_rows_vectorizable(const float* feature, int F);

template <int NQ>
static void launch_fwd_t(const BlendArgs& a, cudaStream_t s)
{
	const int grid = a.grid_x * a.grid_y * 8;
	// a feature row shorter than its padded NQ-1 quads would leave stale shared memory in the tail quads: only the
	// exact fits take the bulk-copy path
	const bool vec = (NQ == 1) || (feature_rows_vectorizable(a.feature, a.F) && a.F == 4 * (NQ - 1));
	if (vec) blend_fwd_simt_kernel<NQ, true><<<grid, 32, 0, s>>>(a);
	else blend_fwd_simt_kernel<NQ, false><<<grid, 32, 0, s>>>(a);
}

void launch_blend_fwd_simt(const BlendArgs& a, cudaStream_t s)
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
