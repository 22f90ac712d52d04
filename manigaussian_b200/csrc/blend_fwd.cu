// Forward blend: front-to-back alpha compositing of RGB + depth + F feature channels.
//
// Replaces FORWARD::render / renderCUDA<3,F> (DGR/cuda_rasterizer/forward.cu:262-398) behind the C-ABI.
// Same per-pixel semantics (power > 0 skip, alpha = min(0.99, o*exp(power)), alpha < 1/255 skip, stop
// when T*(1-alpha) < 1e-4, colour gets + T*bg, features/depth do not; final_T and n_contrib saved).
//
// B200 design (not the reference's; see blend_common.cuh for the decomposition):
//  * one single-warp CTA per 8x4 pixel block; the warp streams its tile's sorted records through a ring of 1-D TMA
//    bulk copies, culls every 32-record chunk against the block with the per-Gaussian alpha >= 1/255 footprint
//    (exact-conservative, one ballot) and QUEUES the survivors in shared memory;
//  * whenever FWD_CH survivors are queued the warp runs one chunk in two phases:
//      walk       lane = pixel, sequential over the chunk's survivors: the exact transmittance recurrence of the
//                 reference (same expf, same skip/stop tests, so final_T and n_contrib agree bit for bit); instead of
//                 36 channel FMAs per survivor (7 of 32 lanes active on the benchmark cloud) the lane only stores its
//                 blend weight w = alpha*T (0 when the pair does not contribute) into a [survivor][pixel] tile;
//      contract   acc[pixel][channel] += W[pixel][survivor] * C[survivor][channel] on the tensor cores (mma.sync
//                 m16n8k8 TF32, 3xTF32 split for fp32 accuracy, blend_mma.cuh), accumulators in the D fragments;
//    the channel rows C of the chunk ({r,g,b,depth} quad + feature row, gathered by Gaussian id with 16-byte cp.async
//    pieces, SASS LDGSTS) travel while the walk runs;
//  * index assignment that lets every MMA operand arrive in place (an A fragment is four consecutive registers, and a
//    register move per fragment would cost more than the MMA saves): MMA row 16 mt + 8 h + gid is pixel (x = gid,
//    y = 2 mt + h) = lane 4 gid + y of the walk, and the weight tile is stored per k-step as [tig][gid][mt][e][h]
//    (survivor 8 ks + tig + 4 e), so {a0..a3} of one 16-pixel tile is ONE LDS.128 and the hi/lo split writes the second
//    quad; the B pair {b0, b1} (survivors g0, g1, feature 8 nt + gid) is two conflict-free scalar loads of the channel rows;
//    with x = gid the output planes are written as full 32-byte sectors straight from the D fragments.
#include "blend_common.cuh"
#include "blend_mma.cuh"
#include "loss_heads.cuh"

namespace mgs {

#ifndef MGS_FWD_CH
#define MGS_FWD_CH 32
#endif
#ifndef MGS_FWD_MIN_CTAS
#define MGS_FWD_MIN_CTAS 14
#endif
#ifndef MGS_FWD_BATCH
#define MGS_FWD_BATCH 64
#endif
constexpr int FWD_CH = MGS_FWD_CH;        // survivors per tensor-core chunk (multiple of 8, <= 32)
constexpr int FWD_QCAP = FWD_CH + 32;     // survivor queue: a chunk plus the survivors of one more 32-record cull
constexpr int FWD_BATCH = MGS_FWD_BATCH;  // records per bulk copy (a multiple of the 32-record cull chunk)
static_assert(FWD_CH % 8 == 0 && FWD_CH <= 32 && FWD_BATCH % 32 == 0, "forward chunk size");

template <int NFT, bool VEC>
__global__ void __launch_bounds__(32, MGS_FWD_MIN_CTAS) blend_fwd_kernel(BlendArgs a)
{
	using L = RowLayout<NFT>;
	constexpr int RS = L::RS;
	constexpr int NT = NFT + 1;  // column tiles: NFT feature tiles + the {r,g,b,depth} tile
	__shared__ __align__(128) InstRec s_rec[RING * FWD_BATCH];
	__shared__ __align__(16) float4 s_q[FWD_QCAP * 2];    // survivor queue, linear: {x,y,ca,cb}, {cc,op,pos,id}; the chunk is its head
	__shared__ __align__(16) float s_rows[FWD_CH * RS];   // channel rows of the current chunk
	__shared__ __align__(16) float s_w[FWD_CH * 32];      // blend weights, per k-step [tig][gid][mt][e][h]: survivor 8 ks + tig + 4 e, pixel 4 gid + 2 mt + h
	__shared__ __align__(8) uint64_t s_bar[RING];

#ifdef MGS_CTA_LOG
	const unsigned long long t_start = cta_log_now();
#endif
	const int lane = threadIdx.x;
	const int gid = lane >> 2, tig = lane & 3;
	const int sub = blockIdx.x & 7;
	const int tile = a.tile_order ? (int)a.tile_order[blockIdx.x >> 3] : (int)(blockIdx.x >> 3);  // longest lists first
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int bx0 = tile_x * TILE_X + (sub & 1) * WARP_BX;
	const int by0 = tile_y * TILE_Y + (sub >> 1) * WARP_BY;
	// walk role: lane = pixel (x = lane >> 2, y = lane & 3)
	const int pxi = bx0 + (lane >> 2), pyi = by0 + (lane & 3);
	const bool inside = pxi < a.W && pyi < a.H;
	if (__all_sync(0xffffffffu, !inside)) return;  // block entirely outside the image (ragged right/bottom tiles)
	const float pfx = (float)pxi, pfy = (float)pyi;
	const int F = a.F;

	// rows beyond a short last chunk are multiplied by zero weights: they must hold finite numbers
	for (int i = lane; i < FWD_CH * RS / 4; i += 32) reinterpret_cast<float4*>(s_rows)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

	const uint2 range = a.ranges[tile];
	WarpRecRingT<FWD_BATCH> ring;
	ring.init(s_rec, s_bar, a.recs + range.x, (int)(range.y - range.x), false);
	const int nb = ring.num_batches();
	int issued = 0, waited = 0;
	for (; issued < min(nb, RING); issued++) ring.issue(issued);

	// T is the running product of (1 - alpha) over every evaluated pair, the stopping pair included: the product only
	// shrinks, so once one test T (1 - alpha) < 1e-4 has failed every later one fails too and no "done" flag has to sit
	// in the recurrence (the reference stops looking at that point; here the later pairs get zero weight).  T_out is the
	// transmittance after the last accepted pair = the reference's final T.  Pixels outside the image start at 0.
	float T = inside ? 1.0f : 0.0f, T_out = 1.0f;
	uint32_t last_contributor = 0;
	float acc[2][NT][4];
#pragma unroll
	for (int mt = 0; mt < 2; mt++)
#pragma unroll
		for (int nt = 0; nt < NT; nt++)
#pragma unroll
			for (int i = 0; i < 4; i++) acc[mt][nt][i] = 0.f;

	int qcount = 0;
	const int wofs = (lane >> 2) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1);  // this pixel's slot in a [tig][gid][mt][e][h] weight tile

	// consume the first cnt (<= FWD_CH) queued survivors; returns true when every pixel of the block is finished
	auto process_chunk = [&](int cnt) -> bool {
		// ---- start the gather of the chunk's channel rows (lands while the walk runs) ----
		if (VEC) {
			// lane r knows the Gaussian id of survivor r; the feature rows are cut into 16-byte pieces and 32 / (2 NFT) rows
			// are copied per trip (lane = row-in-trip x piece), the id travelling by shuffle: 4 instructions per LDGSTS
			const uint32_t myid = lane < cnt ? __float_as_uint(s_q[2 * lane + 1].w) : 0u;
			if (NFT > 0) {
				constexpr int NPF = NFT > 0 ? 2 * NFT : 1, RPI = 32 / NPF;  // pieces per feature row, rows per trip
				const int sub = lane / NPF, q = lane % NPF;
#pragma unroll
				for (int it = 0; it < FWD_CH / RPI; it++) {
					const int r = it * RPI + sub;
					const uint32_t id = __shfl_sync(0xffffffffu, myid, r);
					if (r < cnt) cp_async16(s_rows + r * RS + 4 * q, reinterpret_cast<const float4*>(a.feature + (size_t)id * (8 * NFT)) + q);
				}
			}
			if (lane < cnt) cp_async16(s_rows + lane * RS + L::RGBD, a.rgbd + myid);
			cp_async_commit();
		} else if (lane < cnt) {  // feature rows that are not whole 8-wide tiles (e.g. F = 3): plain loads, zero padded
			const uint32_t id = __float_as_uint(s_q[2 * lane + 1].w);
			float* row = s_rows + lane * RS;
			*reinterpret_cast<float4*>(row + L::RGBD) = a.rgbd[id];
			if (NFT > 0) {
				const float* f = a.feature + (size_t)id * F;
#pragma unroll
				for (int i = 0; i < 8 * NFT; i++) row[i] = (i < F) ? __ldg(f + i) : 0.f;
			}
		}
		// ---- walk: one k-step (8 survivors) per trip, in two passes so that the eight footprint evaluations (independent:
		// LDS, 8 FP ops, expf) overlap in the pipeline and only the short transmittance recurrence is sequential.
		// Slots beyond cnt hold stale records and get zero weights. ----
		const int nks = (cnt + 7) >> 3;
		int ks = 0;
		for (; ks < nks; ks++) {
			if (ks > 0 && __all_sync(0xffffffffu, T < T_STOP)) break;
			const float4* q4 = s_q + 16 * ks;
			float* wrow = s_w + 256 * ks;
			float al[8];      // alpha of the pair, 0 where the pair is skipped (power > 0, alpha < 1/255, empty slot)
			uint32_t ps[8];   // 1-based position of the survivor in the tile's list
#pragma unroll
			for (int u = 0; u < 8; u++) {
				const float4 p = q4[2 * u], q = q4[2 * u + 1];
				const float dx = p.x - pfx, dy = p.y - pfy;
				const float power = -0.5f * (p.z * dx * dx + q.x * dy * dy) - p.w * dx * dy;
				const float alpha = min(ALPHA_MAX, q.y * expf(power));
				const bool ok = (8 * ks + u < cnt) && !(power > 0.0f) && !(alpha < ALPHA_MIN);
				al[u] = ok ? alpha : 0.f;
				ps[u] = __float_as_uint(q.z);
			}
#pragma unroll
			for (int u = 0; u < 8; u++) {
				const float alpha = al[u];
				const float test_T = T * (1 - alpha);
				const bool use = !(test_T < T_STOP);  // a skipped pair (alpha = 0) passes with zero weight and T unchanged
				wrow[(u & 3) * 64 + (u >> 2) * 2 + wofs] = use ? alpha * T : 0.f;
				T_out = use ? test_T : T_out;
				last_contributor = (use && alpha != 0.f) ? ps[u] : last_contributor;
				T = test_T;
			}
		}
		const int nk = ks;  // k-steps to contract
		if (VEC) cp_async_wait<0>();
		__syncwarp();
		// ---- contract on the tensor cores ----
		for (int ks = 0; ks < nk; ks++) {
			const int g0 = 8 * ks + tig, g1 = g0 + 4;
			// {a0..a3} of pixel tile mt: weights of survivors g0 (k = tig) and g1 (k = tig + 4) at pixels 4 gid + 2 mt + {0, 1}
			const float4 wa = *reinterpret_cast<const float4*>(s_w + ks * 256 + tig * 64 + gid * 8);
			const float4 wb = *reinterpret_cast<const float4*>(s_w + ks * 256 + tig * 64 + gid * 8 + 4);
			uint32_t ahi[2][4], alo[2][4];
			tf32_split(wa.x, ahi[0][0], alo[0][0]); tf32_split(wa.y, ahi[0][1], alo[0][1]);
			tf32_split(wa.z, ahi[0][2], alo[0][2]); tf32_split(wa.w, ahi[0][3], alo[0][3]);
			tf32_split(wb.x, ahi[1][0], alo[1][0]); tf32_split(wb.y, ahi[1][1], alo[1][1]);
			tf32_split(wb.z, ahi[1][2], alo[1][2]); tf32_split(wb.w, ahi[1][3], alo[1][3]);
			const float* row0 = s_rows + g0 * RS;
			const float* row1 = s_rows + g1 * RS;
			// B fragments {b0, b1} = column gid of tile nt for survivors g0 and g1: scalar, conflict-free loads straight into
			// the operand pair (a 128-bit row load would need a register move per operand)
#pragma unroll
			for (int nt = 0; nt < NFT; nt++) {
				uint32_t bh0, bl0, bh1, bl1;
				tf32_split(row0[8 * nt + gid], bh0, bl0);
				tf32_split(row1[8 * nt + gid], bh1, bl1);
				mma_3xtf32(acc[0][nt], ahi[0], alo[0], bh0, bh1, bl0, bl1);
				mma_3xtf32(acc[1][nt], ahi[1], alo[1], bh0, bh1, bl0, bl1);
			}
			{
				const float c0 = gid < 4 ? row0[L::RGBD + gid] : 0.f;
				const float c1 = gid < 4 ? row1[L::RGBD + gid] : 0.f;
				uint32_t bh0, bl0, bh1, bl1;
				tf32_split(c0, bh0, bl0);
				tf32_split(c1, bh1, bl1);
				mma_3xtf32(acc[0][NFT], ahi[0], alo[0], bh0, bh1, bl0, bl1);
				mma_3xtf32(acc[1][NFT], ahi[1], alo[1], bh0, bh1, bl0, bl1);
			}
		}
		// the (< 32) survivors behind the chunk move to the head of the queue
		const int left = qcount - cnt;
		float4 m0, m1;
		if (lane < left) { m0 = s_q[2 * (cnt + lane)]; m1 = s_q[2 * (cnt + lane) + 1]; }
		__syncwarp();  // also: s_w and s_rows are rewritten by the next chunk
		if (lane < left) { s_q[2 * lane] = m0; s_q[2 * lane + 1] = m1; }
		qcount = left;
		__syncwarp();
		return __all_sync(0xffffffffu, T < T_STOP);
	};

	// Fill the queue from the record stream, run a chunk whenever FWD_CH survivors are queued, and the remainder at the
	// end.  One call site for the chunk: the kernel's code stays within the instruction cache.
	int k = 0, c = 0, lo = 0, n = 0;
	const float4* rec4 = nullptr;
	bool open = false;  // batch k is landed and partly culled
	for (;;) {
		while (qcount < FWD_CH && k < nb) {
			if (!open) {
				rec4 = ring.wait(k);
				waited = k + 1;
				lo = ring.batch_lo(k); n = ring.batch_n(k);
				c = 0;
				open = true;
			}
			const int j = c + lane;
			float4 r0, r1;
			bool hit = false;
			if (j < n) {
				r1 = rec4[2 * j + 1];
				hit = rec_hits_block(r1, sub);
				if (hit) r0 = rec4[2 * j];
			}
			const uint32_t mask = __ballot_sync(0xffffffffu, hit);
			if (hit) {
				const int e = qcount + __popc(mask & ((1u << lane) - 1u));
				r1.z = __uint_as_float((uint32_t)(lo + j) + 1u);  // the cull extent is spent: keep the 1-based list position instead
				s_q[2 * e] = r0;
				s_q[2 * e + 1] = r1;
			}
			qcount += __popc(mask);
			__syncwarp();
			c += 32;
			if (c >= n) {
				// every survivor of this batch sits in the queue: its buffer is free for the batch RING ahead
				open = false;
				if (issued < nb) { ring.issue(issued); issued++; }
				k++;
			}
		}
		if (qcount == 0) break;
		if (process_chunk(min(qcount, FWD_CH))) break;  // every pixel of the block is finished
	}
	for (int k = waited; k < issued; k++) ring.wait(k);  // no bulk copy may be in flight when the CTA exits

	// ---- outputs ----
	const size_t HW = (size_t)a.H * a.W;
	if (inside) {
		const size_t pix = (size_t)a.W * pyi + pxi;
		T = T_out;
		a.final_T[pix] = T;
		a.n_contrib[pix] = last_contributor;
	}
	s_w[lane] = T;
	__syncwarp();
	const int ox = bx0 + gid;  // fragment role: this thread holds pixels (x = gid, y = 0..3)
	const bool heads = a.tgt_color != nullptr;              // fused loss heads (loss_heads.cuh)
	const bool embed = heads && NFT > 0 && a.tgt_feature != nullptr;
	const float invN = 1.0f / (float)HW;
	float s_rgb = 0.f, s_cos = 0.f;                         // this thread's share of the two loss sums
#pragma unroll
	for (int mt = 0; mt < 2; mt++) {
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const int y = 2 * mt + h, oy = by0 + y;
			const bool in = ox < a.W && oy < a.H;
			const size_t pix = in ? (size_t)a.W * oy + ox : 0;
			const float Tp = s_w[4 * gid + y];
#pragma unroll
			for (int nt = 0; nt < NFT; nt++) {
#pragma unroll
				for (int e = 0; e < 2; e++) {
					const int f = 8 * nt + 2 * tig + e;
					if (in && f < F) a.out_feature[(size_t)f * HW + pix] = acc[mt][nt][2 * h + e];
				}
			}
			float col[2];
#pragma unroll
			for (int e = 0; e < 2; e++) {
				const int c = 2 * tig + e;  // column of the {r,g,b,depth} tile
				col[e] = acc[mt][NFT][2 * h + e] + (c < 3 ? Tp * a.bg[c] : 0.f);
				if (in && c < 3) a.out_color[(size_t)c * HW + pix] = col[e];
				else if (in && c == 3 && a.out_depth) a.out_depth[pix] = col[e];
			}
			if (heads) {  // warp-uniform
#pragma unroll
				for (int e = 0; e < 2; e++) {
					const int c = 2 * tig + e;
					if (in && c < 3) {
						const float d = col[e] - a.tgt_color[(size_t)c * HW + pix];
						s_rgb += d * d;
						a.cot_color[(size_t)c * HW + pix] = (2.0f / 3.0f) * invN * d;
					}
				}
				if (embed) {
					// this thread holds 2 NFT of the pixel's features; the other three threads of the quad hold the rest
					float tg[NFT > 0 ? NFT : 1][2];
					float xy = 0.f, xx = 0.f, yy = 0.f;
#pragma unroll
					for (int nt = 0; nt < NFT; nt++) {
#pragma unroll
						for (int e = 0; e < 2; e++) {
							const int f = 8 * nt + 2 * tig + e;
							const float x = acc[mt][nt][2 * h + e];
							const float g = (in && f < F) ? a.tgt_feature[(size_t)f * HW + pix] : 0.f;
							tg[nt][e] = g;
							xy += x * g; xx += x * x; yy += g * g;
						}
					}
#pragma unroll
					for (int o = 1; o <= 2; o <<= 1) {
						xy += __shfl_xor_sync(0xffffffffu, xy, o);
						xx += __shfl_xor_sync(0xffffffffu, xx, o);
						yy += __shfl_xor_sync(0xffffffffu, yy, o);
					}
					const CosTerms ct = cos_terms(xy, xx, yy);
					if (in && tig == 0) s_cos += ct.cos;
#pragma unroll
					for (int nt = 0; nt < NFT; nt++) {
#pragma unroll
						for (int e = 0; e < 2; e++) {
							const int f = 8 * nt + 2 * tig + e;
							if (in && f < F) a.cot_feature[(size_t)f * HW + pix] = -invN * cos_grad(ct, acc[mt][nt][2 * h + e], tg[nt][e]);
						}
					}
				}
			}
		}
	}
	if (heads) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			s_rgb += __shfl_xor_sync(0xffffffffu, s_rgb, o);
			s_cos += __shfl_xor_sync(0xffffffffu, s_cos, o);
		}
		if (lane == 0) {
			red_add(a.loss_acc, s_rgb);
			if (embed) red_add(a.loss_acc + 1, s_cos);
		}
	}
#ifdef MGS_CTA_LOG
	cta_log_put(a, t_start, 0u, (unsigned int)sub, range.y - range.x);
#endif
}

int blend_supported(int F) { return F >= 0 && F <= 32; }

// 128-bit feature copies need 16-byte aligned rows of a 16-byte multiple
bool feature_rows_vectorizable(const float* feature, int F)
{
	return F > 0 && (F & 3) == 0 && (reinterpret_cast<uintptr_t>(feature) & 15) == 0;
}

template <int NFT>
static void launch_fwd_t(const BlendArgs& a, cudaStream_t s)
{
	const int grid = a.grid_x * a.grid_y * 8;
	// the cp.async row gather copies whole 8-wide feature tiles: exact fits only; other widths take the padded path
	const bool vec = (NFT == 0) || (a.F == 8 * NFT && feature_rows_vectorizable(a.feature, a.F));
	if (vec) blend_fwd_kernel<NFT, true><<<grid, 32, 0, s>>>(a);
	else blend_fwd_kernel<NFT, false><<<grid, 32, 0, s>>>(a);
}


void launch_blend_fwd(const BlendArgs& a, cudaStream_t s)
{
	switch (nft_for(a.F)) {
	case 0: launch_fwd_t<0>(a, s); break;
	case 1: launch_fwd_t<1>(a, s); break;
	case 2: launch_fwd_t<2>(a, s); break;
	default: launch_fwd_t<4>(a, s); break;
	}
}

}  // namespace mgs
