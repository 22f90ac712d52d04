// Backward blend: per-pixel back-to-front replay that emits per-Gaussian gradients.
//
// Replaces BACKWARD::render / renderCUDA<3,F> (DGR/cuda_rasterizer/backward.cu:399-593) behind the
// C-ABI.  Same mathematics: walking a pixel's contributors back to front,
//   T_k     = T_{k+1} / (1 - a_k)                      (T recovered by division from final_T, :521)
//   dL/dc_k = a_k T_k g                                 (g = dL/dpixel, all channels)
//   dL/da_k = T_k (c_k . g) - (S_k + T_final (bg . g_rgb)) / (1 - a_k),   S_k = sum_{j behind k} a_j T_j (c_j . g)
//   dL/dG   = o_k dL/da_k  (the 0.99 clamp is gradient-transparent, :574), then mean2D / conic / opacity.
// S_k is the scalar form of the reference's per-channel `accum_rec` recursion.
//
// B200 design.  The reference has every pixel-thread issue 9+F global float atomics per contributing pair (41 at F=32).
// Here the sums over the pixels of a warp's 8x4 block ARE matrix products, and they run on the tensor cores
// (mma.sync m16n8k8 TF32, 3xTF32 split, blend_mma.cuh).  A single-warp CTA streams the first max(n_contrib) records of
// its tile back to front through the TMA ring, culls them against its block and queues the survivors (blend_common.cuh,
// like the forward); per chunk of BWD_CH queued survivors:
//   1. gather    the chunk's channel rows C[g][ch] by Gaussian id (cp.async, 16-byte pieces)
//   2. GEMM 1    Wd[g][p] = sum_ch C[g][ch] Gpx[p][ch]          (c_k . g for every survivor x pixel; Gpx = the block's
//                                                                cotangent rows, staged once per CTA)
//   3. walk      lane = pixel, sequential over the chunk: alpha, T_k, S_k exactly as above (the per-pixel carried
//                state is two registers); writes wgt[g][p] = a_k T_k and q[g][p] = dL/dG * G, zero for pairs that did
//                not contribute (the same tests as the forward, position <= n_contrib)
//   4. GEMM 2    dC[g][ch] = sum_p wgt[g][p] Gpx[p][ch]          (dL/dcolour, dL/ddepth, dL/dfeature)
//   5. GEMM 3    M[g][m]   = sum_p q[g][p] Y[p][m],  Y[p] = {1, x, y, x^2, x y, y^2} of pixel p relative to the block:
//                the six moments from which lane = Gaussian assembles dL/dmean2D, dL/dconic, dL/dopacity
//                (sum_p q (u - x)^2 = u^2 M0 - 2 u Mx + Mxx with u = mean.x - block.x, ...; Y is exact in TF32)
//   6. flush     per (block, Gaussian): 3 red.global.add.v4.f32 of the 10 small gradients + the feature row as
//                red.v4 straight from the D fragments (column gid of feature tile nt is feature NFT*gid + nt, so a thread
//                holds 8 consecutive features of its four Gaussians).
// Index assignment.  An MMA A operand is four consecutive registers, a B operand two; a register move per operand would
// cost more than the tensor cores save, so every shared-memory tile is laid out such that operands ARRIVE in place:
//   pixels      p = 2 pi + e  (pair pi = 0..15).  K-side (GEMM 2/3): k-step pi >> 2, slots tig = pi & 3 (e = 0) and
//               tig + 4 (e = 1).  M-side (GEMM 1): tile pi >> 3, rows gid = pi & 7 (e = 0) and gid + 8 (e = 1).
//   survivors   j.  M-side (GEMM 2/3): tile j >> 4, rows gid = (j & 15) >> 1 (even j) and gid + 8 (odd j).
//               N-side (GEMM 1): tile j >> 3, column j & 7.
//   channels    K-side (GEMM 1): step ch >> 3, slots tig = (ch & 7) >> 1 (even ch) and tig + 4 (odd ch).
//               N-side (GEMM 2): column gid of feature tile nt is feature NFT gid + nt; {r,g,b,depth} tile: column gid.
//   s_g  [pi][ch][e]              cotangents: GEMM 1's A quad {(e0,ch),(e1,ch),(e0,ch+1),(e1,ch+1)} is one LDS.128, GEMM 2's
//                                 B pairs {(e0,ch),(e1,ch)} for a thread's NFT consecutive channels are LDS.128s
//   s_rows [j][ch]                channel rows as gathered: GEMM 1's B pair {ch, ch+1} is one LDS.64
//   s_x, s_wgt [pi][j >> 1][e][j & 1]   GEMM 1's D fragment is one STS.128, GEMM 2/3's A quad one LDS.128; the walk
//                                 (lane = pixel) reads Wd and writes q in place
#include "blend_common.cuh"
#include "blend_mma.cuh"

namespace mgs {

#ifndef MGS_BWD_CH
#define MGS_BWD_CH 32
#endif
#ifndef MGS_BWD_MIN_CTAS
#define MGS_BWD_MIN_CTAS 12
#endif
#ifndef MGS_BWD_BATCH
#define MGS_BWD_BATCH 64
#endif
constexpr int BWD_CH = MGS_BWD_CH;        // survivors per chunk: 16 or 32 (whole 16-row MMA tiles)
constexpr int BWD_QCAP = BWD_CH + 32;
constexpr int BWD_BATCH = MGS_BWD_BATCH;
constexpr int BWD_MT = BWD_CH / 16;       // 16-row tiles of Gaussians
static_assert(BWD_CH == 16 || BWD_CH == 32, "backward chunk size");
static_assert(BWD_BATCH % 32 == 0, "backward record batch");

template <int NFT, bool VEC>
__global__ void __launch_bounds__(32, MGS_BWD_MIN_CTAS) blend_bwd_kernel(BlendArgs a)
{
	constexpr int NT = NFT + 1;          // channel tiles of 8: NFT feature tiles + {r,g,b,depth, 4 x zero}
	constexpr int CHP = 8 * NT;          // padded channel count
	constexpr int RS = CHP;              // channel-row stride (floats): [8 NFT features | r g b depth | 0 0 0 0]
	constexpr int RGBD = 8 * NFT;
	constexpr int GS = 2 * CHP;          // s_g stride per pixel pair
	constexpr int XS = 2 * BWD_CH + 4;   // s_x / s_wgt stride per pixel pair (+4: spreads the pairs over the banks)
	constexpr int ROWS_FLOATS = (BWD_CH * RS > 16 * XS) ? BWD_CH * RS : 16 * XS;  // the rows buffer is reused for the wgt tile
	__shared__ __align__(128) InstRec s_rec[RING * BWD_BATCH];
	__shared__ __align__(16) float4 s_q[BWD_QCAP * 2];     // survivor queue (linear, back to front): {x,y,ca,cb}, {cc,op,pos,id}
	__shared__ __align__(16) float s_rows[ROWS_FLOATS];    // C[j][RS]; after GEMM 1: wgt tile
	__shared__ __align__(16) float s_x[16 * XS];           // Wd tile; the walk overwrites it with q in place; then per-Gaussian staging
	__shared__ __align__(16) float s_g[16 * GS];           // cotangent rows of the block's pixels
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
	const int pxi = bx0 + (lane >> 2), pyi = by0 + (lane & 3);  // walk role: lane = pixel (x = lane >> 2, y = lane & 3)
	const bool inside = pxi < a.W && pyi < a.H;
	const float fbx0 = (float)bx0, fby0 = (float)by0;
	const size_t HW = (size_t)a.H * a.W;
	const size_t pix = (size_t)a.W * pyi + pxi;
	const int F = a.F;

	// ---- per-pixel state and cotangent rows ----
	uint32_t nc = 0;
	float T = 0.f, bgterm = 0.f;
	{
		float g4[4] = { 0.f, 0.f, 0.f, 0.f };
		// upstream gradients of the fused loss heads (loss_heads.cuh): the cotangent planes were written for d loss = 1
		const float sc_color = a.cot_scale ? a.cot_scale[0] : 1.f, sc_feat = a.cot_scale ? a.cot_scale[1] : 1.f;
		if (inside) {
			nc = a.n_contrib[pix];
			T = a.final_T[pix];
#pragma unroll
			for (int ch = 0; ch < 3; ch++) g4[ch] = sc_color * a.dL_dcolor[ch * HW + pix];
			if (a.dL_ddepth) g4[3] = a.dL_ddepth[pix];
		}
		bgterm = T * (a.bg[0] * g4[0] + a.bg[1] * g4[1] + a.bg[2] * g4[2]);
		float* grow = s_g + (lane >> 1) * GS + (lane & 1);  // element (pixel, ch) at [pi][ch][e]
#pragma unroll
		for (int c = 0; c < 4; c++) { grow[2 * (RGBD + c)] = g4[c]; grow[2 * (RGBD + 4 + c)] = 0.f; }
		if (NFT > 0) {
#pragma unroll
			for (int i = 0; i < 8 * NFT; i++)
				grow[2 * i] = (inside && i < F && a.dL_dfeature) ? sc_feat * a.dL_dfeature[(size_t)i * HW + pix] : 0.f;
		}
	}
	uint32_t maxc = nc;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
	if (maxc == 0) return;  // nothing blended into this block
	for (int i = lane; i < ROWS_FLOATS / 4; i += 32) reinterpret_cast<float4*>(s_rows)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	__syncwarp();

	const uint2 range = a.ranges[tile];
	WarpRecRingT<BWD_BATCH> ring;
	// only the first maxc instances of the tile's list can have contributed to a pixel of this block
	ring.init(s_rec, s_bar, a.recs + range.x, (int)maxc, true);
	const int nb = ring.num_batches();
	int issued = 0;
	for (; issued < min(nb, RING); issued++) ring.issue(issued);

	const float pfx = (float)pxi, pfy = (float)pyi;
	float S = 0.f;  // sum over the contributors behind the current one of a_j T_j (c_j . g)
	const int xofs = (lane >> 1) * XS + (lane & 1) * 2;  // this pixel's slot in the [pi][j >> 1][e][j & 1] tiles
	// GEMM 3's B operand: moments basis of the two pixels of k-slots (ks, tig) and (ks, tig + 4): p = 8 ks + 2 tig + e,
	// x = p >> 2, y = p & 3; column gid selects {1, x, y, x^2, x y, y^2, 0, 0}
	uint32_t ymom[4][2];
#pragma unroll
	for (int ks = 0; ks < 4; ks++) {
#pragma unroll
		for (int e = 0; e < 2; e++) {
			const int p = 8 * ks + 2 * tig + e;
			const float x = (float)(p >> 2), y = (float)(p & 3);
			const float v = gid == 0 ? 1.f : gid == 1 ? x : gid == 2 ? y : gid == 3 ? x * x : gid == 4 ? x * y : gid == 5 ? y * y : 0.f;
			ymom[ks][e] = __float_as_uint(v);
		}
	}
	const float half_W = 0.5f * a.W, half_H = 0.5f * a.H;
	int qcount = 0;

	// consume the first cnt (<= BWD_CH) queued survivors
	auto process_chunk = [&](int cnt) {
		// ---- 1. gather the channel rows ----
		if (VEC) {
			// lane r knows the Gaussian id of survivor r; 32 / (2 NFT) feature rows are copied per trip (lane = row x 16-byte
			// piece), the id travelling by shuffle
			const uint32_t myid = lane < cnt ? __float_as_uint(s_q[2 * lane + 1].w) : 0u;
			if (NFT > 0) {
				constexpr int NPF = NFT > 0 ? 2 * NFT : 1, RPI = 32 / NPF;
				const int sub = lane / NPF, q = lane % NPF;
#pragma unroll
				for (int it = 0; it < BWD_CH / RPI; it++) {
					const int r = it * RPI + sub;
					const uint32_t id = __shfl_sync(0xffffffffu, myid, r);
					if (r < cnt) cp_async16(s_rows + r * RS + 4 * q, reinterpret_cast<const float4*>(a.feature + (size_t)id * (8 * NFT)) + q);
				}
			}
			if (lane < cnt) cp_async16(s_rows + lane * RS + RGBD, a.rgbd + myid);
			cp_async_commit();
			cp_async_wait<0>();
		} else if (lane < cnt) {
			const uint32_t id = __float_as_uint(s_q[2 * lane + 1].w);
			float* row = s_rows + lane * RS;
			*reinterpret_cast<float4*>(row + RGBD) = a.rgbd[id];
			if (NFT > 0) {
				const float* f = a.feature + (size_t)id * F;
#pragma unroll
				for (int i = 0; i < 8 * NFT; i++) row[i] = (i < F) ? __ldg(f + i) : 0.f;
			}
		}
		__syncwarp();
		// ---- 2. GEMM 1: Wd[p][j] = Gpx[p][:] . C[j][:]   (M = pixels (2 tiles), N = survivors, K = channels) ----
		{
			float wd[2][BWD_CH / 8][4];
#pragma unroll
			for (int mt = 0; mt < 2; mt++)
#pragma unroll
				for (int nt = 0; nt < BWD_CH / 8; nt++)
#pragma unroll
					for (int i = 0; i < 4; i++) wd[mt][nt][i] = 0.f;
#pragma unroll
			for (int ks = 0; ks < NT; ks++) {
				uint32_t ahi[2][4], alo[2][4];
#pragma unroll
				for (int mt = 0; mt < 2; mt++) {
					const float4 v = *reinterpret_cast<const float4*>(s_g + (8 * mt + gid) * GS + (8 * ks + 2 * tig) * 2);
					tf32_split(v.x, ahi[mt][0], alo[mt][0]); tf32_split(v.y, ahi[mt][1], alo[mt][1]);
					tf32_split(v.z, ahi[mt][2], alo[mt][2]); tf32_split(v.w, ahi[mt][3], alo[mt][3]);
				}
#pragma unroll
				for (int nt = 0; nt < BWD_CH / 8; nt++) {
					const float2 v = *reinterpret_cast<const float2*>(s_rows + (8 * nt + gid) * RS + 8 * ks + 2 * tig);
					uint32_t bh0, bl0, bh1, bl1;
					tf32_split(v.x, bh0, bl0);
					tf32_split(v.y, bh1, bl1);
					mma_3xtf32(wd[0][nt], ahi[0], alo[0], bh0, bh1, bl0, bl1);
					mma_3xtf32(wd[1][nt], ahi[1], alo[1], bh0, bh1, bl0, bl1);
				}
			}
			// D fragment (mt, nt): pixels (pi = 8 mt + gid, e = 0 | 1), survivors 8 nt + 2 tig (+1): one 16-byte store
#pragma unroll
			for (int mt = 0; mt < 2; mt++)
#pragma unroll
				for (int nt = 0; nt < BWD_CH / 8; nt++)
					*reinterpret_cast<float4*>(s_x + (8 * mt + gid) * XS + (4 * nt + tig) * 4) =
						make_float4(wd[mt][nt][0], wd[mt][nt][1], wd[mt][nt][2], wd[mt][nt][3]);
		}
		__syncwarp();  // Wd complete; the rows buffer is free for the wgt tile
		// ---- 3. walk, back to front (queue order): lane = pixel.  Per k-step of 8 survivors: first the eight footprint
		// evaluations (independent, they overlap in the pipeline), then the short sequential recurrence on T and S. ----
		float* s_wgt = s_rows;
#pragma unroll 1
		for (int j0 = 0; j0 < BWD_CH; j0 += 8) {
			float al[8], opG[8], rinv[8];  // alpha, opacity * G, 1 / (1 - alpha); alpha = opG = 0, rinv = 1 for pairs that did not contribute
#pragma unroll
			for (int u = 0; u < 8; u++) {
				const int j = j0 + u;
				const float4 r0 = s_q[2 * j], r1 = s_q[2 * j + 1];
				const float dx = r0.x - pfx, dy = r0.y - pfy;
				const float op = r1.y;
				const float power = -0.5f * (r0.z * dx * dx + r1.x * dy * dy) - r0.w * dx * dy;
				// ex2.approx-based exp (rel. error ~1e-6); the alpha >= 1/255 decision must agree with the forward's (which uses
				// expf like the reference), so the rare borderline pairs are re-evaluated exactly
				float G = __expf(power);
				float alpha = min(ALPHA_MAX, op * G);
				if (fabsf(alpha - ALPHA_MIN) < 2e-5f * ALPHA_MIN * 8.f) {
					G = expf(power);
					alpha = min(ALPHA_MAX, op * G);
				}
				const bool valid = (j < cnt) && (__float_as_uint(r1.z) <= nc) && (power <= 0.0f) && (alpha >= ALPHA_MIN);
				al[u] = valid ? alpha : 0.f;
				opG[u] = valid ? op * G : 0.f;
				rinv[u] = __fdividef(1.f, 1.f - al[u]);
			}
#pragma unroll
			for (int u = 0; u < 8; u++) {
				const int jofs = (j0 >> 1) * 4 + (u >> 1) * 4 + (u & 1);  // survivor j0 + u
				const float wdot = s_x[xofs + jofs];
				const float Tk = T * rinv[u];                                 // transmittance in front of this Gaussian
				const float wgt = al[u] * Tk;                                 // d(pixel channel) / d(colour of this Gaussian)
				const float dL_dalpha = Tk * wdot - (S + bgterm) * rinv[u];
				s_wgt[xofs + jofs] = wgt;
				s_x[xofs + jofs] = opG[u] * dL_dalpha;                        // q = dL/dG * G  (dL/dG = opacity * dL/dalpha)
				T = Tk;                                                       // unchanged where the pair is skipped (rinv = 1)
				S = fmaf(wgt, wdot, S);                                       // likewise (wgt = 0)
			}
		}
		__syncwarp();
		// ---- 4 + 5. GEMM 2: dC[g][ch] = wgt[g][:] . Gpx[:][ch];  GEMM 3: M[g][m] = q[g][:] . Y[:][m]   (K = pixels) ----
		float dC[BWD_MT][NT][4], mom[BWD_MT][4];
#pragma unroll
		for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
			for (int i = 0; i < 4; i++) mom[mt][i] = 0.f;
#pragma unroll
			for (int nt = 0; nt < NT; nt++)
#pragma unroll
				for (int i = 0; i < 4; i++) dC[mt][nt][i] = 0.f;
		}
#pragma unroll
		for (int ks = 0; ks < 4; ks++) {
			// A quads of wgt and q: pixel pair pi = 4 ks + tig, survivors 16 mt + 2 gid (+1)
			uint32_t ahi[BWD_MT][4], alo[BWD_MT][4], qhi[BWD_MT][4], qlo[BWD_MT][4];
#pragma unroll
			for (int mt = 0; mt < BWD_MT; mt++) {
				const float4 w4 = *reinterpret_cast<const float4*>(s_wgt + (4 * ks + tig) * XS + mt * 32 + gid * 4);
				const float4 q4 = *reinterpret_cast<const float4*>(s_x + (4 * ks + tig) * XS + mt * 32 + gid * 4);
				tf32_split(w4.x, ahi[mt][0], alo[mt][0]); tf32_split(w4.y, ahi[mt][1], alo[mt][1]);
				tf32_split(w4.z, ahi[mt][2], alo[mt][2]); tf32_split(w4.w, ahi[mt][3], alo[mt][3]);
				tf32_split(q4.x, qhi[mt][0], qlo[mt][0]); tf32_split(q4.y, qhi[mt][1], qlo[mt][1]);
				tf32_split(q4.z, qhi[mt][2], qlo[mt][2]); tf32_split(q4.w, qhi[mt][3], qlo[mt][3]);
			}
			// B pairs of GEMM 2: cotangents of the pair's two pixels, channels NFT gid .. NFT gid + NFT - 1, then {r,g,b,depth}[gid]
			const float* gp = s_g + (4 * ks + tig) * GS;
			if (NFT > 0) {
				float f[2 * (NFT > 0 ? NFT : 1)];
				if (NFT == 4) {
					const float4 u0 = *reinterpret_cast<const float4*>(gp + 8 * gid), u1 = *reinterpret_cast<const float4*>(gp + 8 * gid + 4);
					f[0] = u0.x; f[1] = u0.y; f[2 % (2 * NFT)] = u0.z; f[3 % (2 * NFT)] = u0.w;
					f[4 % (2 * NFT)] = u1.x; f[5 % (2 * NFT)] = u1.y; f[6 % (2 * NFT)] = u1.z; f[7 % (2 * NFT)] = u1.w;
				} else if (NFT == 2) {
					const float4 u0 = *reinterpret_cast<const float4*>(gp + 4 * gid);
					f[0] = u0.x; f[1] = u0.y; f[2 % (2 * NFT)] = u0.z; f[3 % (2 * NFT)] = u0.w;
				} else {
					const float2 u0 = *reinterpret_cast<const float2*>(gp + 2 * gid);
					f[0] = u0.x; f[1] = u0.y;
				}
#pragma unroll
				for (int nt = 0; nt < NFT; nt++) {
					uint32_t bh0, bl0, bh1, bl1;
					tf32_split(f[2 * nt], bh0, bl0);
					tf32_split(f[2 * nt + 1], bh1, bl1);
#pragma unroll
					for (int mt = 0; mt < BWD_MT; mt++) mma_3xtf32(dC[mt][nt], ahi[mt], alo[mt], bh0, bh1, bl0, bl1);
				}
			}
			{
				const float2 c = *reinterpret_cast<const float2*>(gp + 2 * (RGBD + gid));  // columns 4..7 are the zero padding
				uint32_t bh0, bl0, bh1, bl1;
				tf32_split(c.x, bh0, bl0);
				tf32_split(c.y, bh1, bl1);
#pragma unroll
				for (int mt = 0; mt < BWD_MT; mt++) mma_3xtf32(dC[mt][NFT], ahi[mt], alo[mt], bh0, bh1, bl0, bl1);
			}
			// GEMM 3: the basis is exact in TF32, only q is split
#pragma unroll
			for (int mt = 0; mt < BWD_MT; mt++) {
				mma_tf32(mom[mt], qlo[mt][0], qlo[mt][1], qlo[mt][2], qlo[mt][3], ymom[ks][0], ymom[ks][1]);
				mma_tf32(mom[mt], qhi[mt][0], qhi[mt][1], qhi[mt][2], qhi[mt][3], ymom[ks][0], ymom[ks][1]);
			}
		}
		__syncwarp();  // every lane has read its q fragments: s_x becomes the per-Gaussian staging area
		// ---- 6. flush ----
		// feature gradients straight from the D fragments; moments and {r,g,b,depth} gradients to s_x[g][0..9]
		float* s_m = s_x;  // [g][12]
#pragma unroll
		for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int g = 16 * mt + 2 * gid + h;  // D rows gid (h = 0) and gid + 8 (h = 1)
				if (tig < 3) *reinterpret_cast<float2*>(s_m + g * 12 + 2 * tig) = make_float2(mom[mt][2 * h], mom[mt][2 * h + 1]);
				if (tig < 2) *reinterpret_cast<float2*>(s_m + g * 12 + 6 + 2 * tig) = make_float2(dC[mt][NFT][2 * h], dC[mt][NFT][2 * h + 1]);
				if (NFT > 0 && g < cnt && a.dL_dfeat) {
					const uint32_t id = __float_as_uint(s_q[2 * g + 1].w);
					float* df = a.dL_dfeat + (size_t)id * F;
					// column 2 tig + e of feature tile nt is feature NFT (2 tig + e) + nt
					if (VEC) {
						if (NFT == 4) {
							red_add_v4(df + 8 * tig, dC[mt][0][2 * h], dC[mt][1][2 * h], dC[mt][2][2 * h], dC[mt][3][2 * h]);
							red_add_v4(df + 8 * tig + 4, dC[mt][0][2 * h + 1], dC[mt][1][2 * h + 1], dC[mt][2][2 * h + 1], dC[mt][3][2 * h + 1]);
						} else if (NFT == 2) {
							red_add_v4(df + 4 * tig, dC[mt][0][2 * h], dC[mt][1 % NT][2 * h], dC[mt][0][2 * h + 1], dC[mt][1 % NT][2 * h + 1]);
						} else {
							red_add(df + 2 * tig, dC[mt][0][2 * h]);
							red_add(df + 2 * tig + 1, dC[mt][0][2 * h + 1]);
						}
					} else {
#pragma unroll
						for (int nt = 0; nt < NFT; nt++) {
#pragma unroll
							for (int e = 0; e < 2; e++) {
								const int f = NFT * (2 * tig + e) + nt;
								if (f < F) red_add(df + f, dC[mt][nt][2 * h + e]);
							}
						}
					}
				}
			}
		}
		__syncwarp();
		if (lane < cnt) {
			const float4 r0 = s_q[2 * lane], r1 = s_q[2 * lane + 1];
			const float4 m0 = *reinterpret_cast<const float4*>(s_m + lane * 12);      // {M0, Mx, My, Mxx}
			const float4 m1 = *reinterpret_cast<const float4*>(s_m + lane * 12 + 4);  // {Mxy, Myy, dr, dg}
			const float2 m2 = *reinterpret_cast<const float2*>(s_m + lane * 12 + 8);  // {db, ddepth}
			const float u = r0.x - fbx0, v = r0.y - fby0;  // mean relative to the block origin: d = (u - x, v - y)
			const float ca = r0.z, cb = r0.w, cc = r1.x, op = r1.y;
			const float q_dx = u * m0.x - m0.y, q_dy = v * m0.x - m0.z;
			const float q_dxx = u * (u * m0.x - 2.f * m0.y) + m0.w;
			const float q_dxy = u * (v * m0.x - m0.z) - v * m0.y + m1.x;
			const float q_dyy = v * (v * m0.x - 2.f * m0.z) + m1.y;
			const float dmx = half_W * (-ca * q_dx - cb * q_dy);
			const float dmy = half_H * (-cc * q_dy - cb * q_dx);
			float* gb = a.gb + (size_t)__float_as_uint(r1.w) * GB_STRIDE;
			red_add_v4(gb, dmx, dmy, -0.5f * q_dxx, -0.5f * q_dxy);
			red_add_v4(gb + 4, -0.5f * q_dyy, __fdividef(m0.x, op), m1.z, m1.w);
			red_add_v4(gb + 8, m2.x, m2.y, 0.f, 0.f);
		}
		// the (< 32) survivors behind the chunk move to the head of the queue
		const int left = qcount - cnt;
		float4 k0, k1;
		if (lane < left) { k0 = s_q[2 * (cnt + lane)]; k1 = s_q[2 * (cnt + lane) + 1]; }
		__syncwarp();
		if (lane < left) { s_q[2 * lane] = k0; s_q[2 * lane + 1] = k1; }
		qcount = left;
		__syncwarp();
	};

	// Fill the queue from the record stream (back to front), run a chunk whenever BWD_CH survivors are queued, and the
	// remainder at the end.  One call site for the chunk: the kernel's code stays within the instruction cache.
	int k = 0, c = 0, lo = 0, n = 0;
	const float4* rec4 = nullptr;
	bool open = false;  // batch k is landed and partly culled
	for (;;) {
		while (qcount < BWD_CH && k < nb) {
			if (!open) {
				rec4 = ring.wait(k);
				lo = ring.batch_lo(k); n = ring.batch_n(k);
				c = ((n - 1) >> 5) << 5;  // chunks of the batch, back to front
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
				// survivors are appended in back-to-front order: the highest lane first
				const uint32_t above = (lane == 31) ? 0u : (mask >> (lane + 1));
				const int e = qcount + __popc(above);
				r1.z = __uint_as_float((uint32_t)(lo + j) + 1u);  // the cull extent is spent: keep the 1-based list position instead
				s_q[2 * e] = r0;
				s_q[2 * e + 1] = r1;
			}
			qcount += __popc(mask);
			__syncwarp();
			c -= 32;
			if (c < 0) {
				// every survivor of this batch sits in the queue: refill its buffer with the batch RING ahead
				open = false;
				if (issued < nb) { ring.issue(issued); issued++; }
				k++;
			}
		}
		if (qcount == 0) break;
		process_chunk(min(qcount, BWD_CH));
	}
#ifdef MGS_CTA_LOG
	cta_log_put(a, t_start, 1u, (unsigned int)sub, maxc);
#endif
}

bool feature_rows_vectorizable(const float* feature, int F);

template <int NFT>
static void launch_bwd_t(const BlendArgs& a, cudaStream_t s)
{
	const int grid = a.grid_x * a.grid_y * 8;
	const bool vec = (NFT == 0) || (a.F == 8 * NFT && feature_rows_vectorizable(a.feature, a.F) &&
		(reinterpret_cast<uintptr_t>(a.dL_dfeat) & 15) == 0);
	if (vec) blend_bwd_kernel<NFT, true><<<grid, 32, 0, s>>>(a);
	else blend_bwd_kernel<NFT, false><<<grid, 32, 0, s>>>(a);
}

void launch_blend_bwd(const BlendArgs& a, cudaStream_t s)
{
	switch (nft_for(a.F)) {
	case 0: launch_bwd_t<0>(a, s); break;
	case 1: launch_bwd_t<1>(a, s); break;
	case 2: launch_bwd_t<2>(a, s); break;
	default: launch_bwd_t<4>(a, s); break;
	}
}

}  // namespace mgs
