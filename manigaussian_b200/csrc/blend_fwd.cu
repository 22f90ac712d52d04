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
//  * two-level software pipeline, all data movement asynchronous TMA (UBLKCP) tracked by mbarriers:
//      records     : RING 64-record batches in flight/resident (one contiguous bulk copy each),
//      channel rows: for the SURVIVORS of chunk g+1, per survivor a 16-byte {r,g,b,depth} copy and an F*4-byte
//                    feature-row copy, gathered by Gaussian id into a double-buffered row array while chunk g is blended;
//  * channel rows are consumed from shared memory as 128-bit broadcasts instead of per-pair scalar global
//    gathers (forward.cu:364-371);
//  * feature width is a run-time value dispatched to NQ = ceil((4+F)/4) in {1,2,3,5,9};
//  * optional SUB-BLOCK LOCKSTEP (MGS_FWD_NSUB = 2 or 4; default 1 = off): the 8x4 block is split into two 4x4 or four
//    4x2 sub-blocks, each chunk is also culled against every sub-block (one ballot each) and a lane walks only the
//    survivors of ITS sub-block, all sub-blocks in lockstep -- Gaussians whose footprints lie in different sub-blocks
//    commute.  Bit-identical results and ~25 % fewer walk steps on the benchmark cloud, but the per-lane list
//    bookkeeping (ffs/popc/address math that is warp-uniform otherwise) costs as much as the steps save: measured on
//    B200 at c3, 0.427 ms (off) vs 0.443 ms (2) vs 0.481 ms (4) per view.  Kept as a compile-time experiment.
#include "blend_common.cuh"

#ifndef MGS_FWD_NSUB
#define MGS_FWD_NSUB 1
#endif

namespace mgs {

constexpr int FWD_NSUB = MGS_FWD_NSUB;
static_assert(FWD_NSUB == 1 || FWD_NSUB == 2 || FWD_NSUB == 4, "sub-block split of the 8x4 block");

template <int NQ, bool VEC>
__global__ void __launch_bounds__(32, 16) blend_fwd_kernel(BlendArgs a)
{
	__shared__ __align__(128) InstRec s_rec[RING * REC_BATCH];
	__shared__ __align__(16) float4 s_rows[2][32 * NQ];
	__shared__ __align__(8) uint64_t s_bar_rec[RING];
	__shared__ __align__(8) uint64_t s_bar_row[2];

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
	// this lane's sub-block: 4 columns wide when split in two, 4x2 when split in four
	const int mysub = FWD_NSUB == 1 ? 0 : (((lane & 7) >> 2) + (FWD_NSUB == 4 ? 2 * (lane >> 4) : 0));
	const int F = a.F;
	const uint32_t row_bytes = 16u + (NQ > 1 ? (uint32_t)F * 4u : 0u);

	const uint2 range = a.ranges[tile];
	WarpRecRing ring;
	ring.init(s_rec, s_bar_rec, a.recs + range.x, (int)(range.y - range.x), false);
	if (lane == 0) {
		mbar_init(&s_bar_row[0], 1);
		mbar_init(&s_bar_row[1], 1);
		mbar_fence_init();
	}
	__syncwarp();
	const int nb = ring.num_batches();
	const int nchunks = (ring.total + 31) >> 5;
	int issued = 0, waited = 0;      // record batches
	uint32_t row_parity = 0;          // bit b = parity of the next phase to wait for on s_bar_row[b]

	float T = 1.0f;
	uint32_t last_contributor = 0;
	bool done = !inside;
	float acc[4 * NQ];
#pragma unroll
	for (int i = 0; i < 4 * NQ; i++) acc[i] = 0.f;

	uint32_t my_next = 0;  // survivors of this lane's sub-block in the chunk being prefetched
	int steps_next = 0;    // max over the sub-blocks of their survivor counts
	// cull chunk g and start gathering the survivors' channel rows into s_rows[g & 1]
	auto prefetch_chunk = [&](int g) -> uint32_t {
		const int k = g >> 1;
		if ((g & 1) == 0) { ring.wait(k); waited = k + 1; }
		__syncwarp();
		const float4* rec4 = ring.buffer(k);
		const int j = ((g & 1) << 5) + lane;
		const bool hit = (g * 32 + lane < ring.total) && rec_hits_block(rec4[2 * j], rec4[2 * j + 1], fbx0, fbx1, fby0, fby1);
		const uint32_t mask = __ballot_sync(0xffffffffu, hit);
		my_next = mask; steps_next = __popc(mask);
		if (FWD_NSUB > 1 && mask) {
			steps_next = 0;
#pragma unroll
			for (int sb = 0; sb < FWD_NSUB; sb++) {
				const float sx0 = fbx0 + 4.f * (sb & 1), sy0 = FWD_NSUB == 4 ? fby0 + 2.f * (sb >> 1) : fby0;
				const float sy1 = FWD_NSUB == 4 ? sy0 + 1.f : fby1;
				const bool hs = hit && rec_hits_block(rec4[2 * j], rec4[2 * j + 1], sx0, sx0 + 3.f, sy0, sy1);
				const uint32_t ms = __ballot_sync(0xffffffffu, hs);
				if (mysub == sb) my_next = ms;
				steps_next = max(steps_next, __popc(ms));
			}
		}
		if (mask) {
			float4* rows = s_rows[g & 1];
			const int rank = __popc(mask & ((1u << lane) - 1u));
			if (VEC) {
				if (lane == 0) {
					fence_proxy_async();
					mbar_arrive_expect_tx(&s_bar_row[g & 1], (uint32_t)__popc(mask) * row_bytes);
				}
				__syncwarp();
				if (hit) {
					const uint32_t id = rec_id(rec4[2 * j + 1]);
					bulk_g2s(rows + rank * NQ, a.rgbd + id, 16u, &s_bar_row[g & 1]);
					if (NQ > 1) bulk_g2s(rows + rank * NQ + 1, a.feature + (size_t)id * F, (uint32_t)F * 4u, &s_bar_row[g & 1]);
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
		return mask;
	};

	uint32_t mask_cur = 0, mask_next = 0, my_cur = 0;
	int steps_cur = 0;
	if (nchunks > 0) {
		for (; issued < min(nb, RING); issued++) ring.issue(issued);
		mask_cur = prefetch_chunk(0);
		my_cur = my_next; steps_cur = steps_next;
	}
	for (int g = 0; g < nchunks; g++) {
		mask_next = 0;
		if (g + 1 < nchunks) mask_next = prefetch_chunk(g + 1);
		if (mask_cur) {
			const int rb = g & 1;
			if (VEC) { mbar_wait(&s_bar_row[rb], (row_parity >> rb) & 1u); row_parity ^= 1u << rb; }
			else __syncwarp();
			const float4* rec4 = ring.buffer(g >> 1) + (rb << 6);
			const float4* rows = s_rows[rb];
			const uint32_t pos0 = (uint32_t)(g * 32) + 1u;  // 1-based position of the chunk's first record in the tile list
			// Two survivors per step: their footprint evaluations (power, expf) are independent of the transmittance
			// recurrence, so issuing both up front overlaps the second one's latency with the first one's channel FMAs.
			auto blend_one = [&](bool act, int b, float power, float alpha_raw, const float4* row) {
				if (!act || done || power > 0.0f) return;
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
			uint32_t mm = my_cur;  // per lane: the survivors of this lane's sub-block, nearest first
			int i = 0;             // FWD_NSUB == 1: rank of the next survivor among the chunk's survivors (warp-uniform)
			for (int it = 0; it < steps_cur; it += 2) {
				const bool act0 = mm != 0;
				const int b0 = act0 ? __ffs(mm) - 1 : 0;
				mm &= mm - 1;
				const bool act1 = mm != 0;
				const int b1 = act1 ? __ffs(mm) - 1 : b0;
				mm &= mm - 1;
				const float4 p0 = rec4[2 * b0], q0 = rec4[2 * b0 + 1];  // {x, y, ca, cb}, {cc, op, ext, id}
				const float4 p1 = rec4[2 * b1], q1 = rec4[2 * b1 + 1];
				const float dx0 = p0.x - pfx, dy0 = p0.y - pfy, dx1 = p1.x - pfx, dy1 = p1.y - pfy;
				const float power0 = -0.5f * (p0.z * dx0 * dx0 + q0.x * dy0 * dy0) - p0.w * dx0 * dy0;
				const float power1 = -0.5f * (p1.z * dx1 * dx1 + q1.x * dy1 * dy1) - p1.w * dx1 * dy1;
				const float a0 = q0.y * expf(power0);
				const float a1 = q1.y * expf(power1);
				// a survivor's channel row sits at its rank among the chunk's (union) survivors
				const int r0i = FWD_NSUB == 1 ? i : __popc(mask_cur & ((1u << b0) - 1u));
				const int r1i = FWD_NSUB == 1 ? i + 1 : __popc(mask_cur & ((1u << b1) - 1u));
				blend_one(act0, b0, power0, a0, rows + r0i * NQ);
				blend_one(act1, b1, power1, a1, rows + r1i * NQ);
				i += 2;
			}
		}
		if (__all_sync(0xffffffffu, done)) {
			// drain whatever is still in flight before the CTA exits
			if (VEC && mask_next) mbar_wait(&s_bar_row[(g + 1) & 1], (row_parity >> ((g + 1) & 1)) & 1u);
			for (int k = waited; k < issued; k++) ring.wait(k);
			break;
		}
		// the batch whose last chunk was just blended frees its buffer for the batch RING ahead
		if ((g & 1) && issued < nb) { ring.issue(issued); issued++; }
		mask_cur = mask_next; my_cur = my_next; steps_cur = steps_next;
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

template <int NQ>
static void launch_fwd_t(const BlendArgs& a, cudaStream_t s)
{
	const int grid = a.grid_x * a.grid_y * 8;
	// a feature row shorter than its padded NQ-1 quads would leave stale shared memory in the tail quads: only the
	// exact fits take the bulk-copy path
	const bool vec = (NQ == 1) || (feature_rows_vectorizable(a.feature, a.F) && a.F == 4 * (NQ - 1));
	if (vec) blend_fwd_kernel<NQ, true><<<grid, 32, 0, s>>>(a);
	else blend_fwd_kernel<NQ, false><<<grid, 32, 0, s>>>(a);
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
