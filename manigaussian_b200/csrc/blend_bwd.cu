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
// Index assignment (pixels, Gaussians, channels <-> MMA rows, columns, k) makes the fragment traffic wide: with pixel
// p = 8 ks + 2 tig + e as k-slot (ks, tig + 4 e) and the [g][pixel] tiles stored in the column order
// pi(p) = 8 tig + 2 ks + e, a thread's A values of GEMM 2/3 are two LDS.128 per Gaussian row and GEMM 1's D fragments
// store as 8-byte pairs; channel rows use stride 36 floats, which makes GEMM 1's row loads and GEMM 2's cotangent loads
// conflict free without padding the gather.
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
#define MGS_BWD_BATCH 32
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
	constexpr int NT = NFT + 1;          // channel column tiles: NFT feature tiles + {r,g,b,depth}
	constexpr int RS = 8 * NFT + 4;      // channel-row stride (floats): [8 NFT features | r g b depth]
	constexpr int RGBD = 8 * NFT;
	constexpr int ROWS_FLOATS = BWD_CH * (RS > 32 ? RS : 32);  // the rows buffer is reused for the wgt tile
	__shared__ __align__(128) InstRec s_rec[RING * BWD_BATCH];
	__shared__ __align__(16) float4 s_q[BWD_QCAP * 2];     // survivor queue (linear, back to front): {x,y,ca,cb}, {cc,op,pos,id}
	__shared__ __align__(16) float s_rows[ROWS_FLOATS];    // C[g][RS]; after GEMM 1: wgt[g][32]
	__shared__ __align__(16) float s_x[BWD_CH * 32];       // Wd[g][32]; the walk overwrites it with q in place; then moments
	__shared__ __align__(16) float s_g[32 * RS];           // Gpx[p][RS]: cotangent rows of the block's pixels
	__shared__ __align__(8) uint64_t s_bar[RING];

	const int lane = threadIdx.x;
	const int gid = lane >> 2, tig = lane & 3;
	const int tile = blockIdx.x >> 3, sub = blockIdx.x & 7;
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int bx0 = tile_x * TILE_X + (sub & 1) * WARP_BX;
	const int by0 = tile_y * TILE_Y + (sub >> 1) * WARP_BY;
	const int pxi = bx0 + (lane >> 2), pyi = by0 + (lane & 3);  // walk role: lane = pixel (x = lane >> 2, y = lane & 3)
	const bool inside = pxi < a.W && pyi < a.H;
	const float fbx0 = (float)bx0, fbx1 = (float)(bx0 + WARP_BX - 1), fby0 = (float)by0, fby1 = (float)(by0 + WARP_BY - 1);
	const size_t HW = (size_t)a.H * a.W;
	const size_t pix = (size_t)a.W * pyi + pxi;
	const int F = a.F;

	// ---- per-pixel state and cotangent rows ----
	uint32_t nc = 0;
	float T = 0.f, bgterm = 0.f;
	{
		float g4[4] = { 0.f, 0.f, 0.f, 0.f };
		if (inside) {
			nc = a.n_contrib[pix];
			T = a.final_T[pix];
#pragma unroll
			for (int ch = 0; ch < 3; ch++) g4[ch] = a.dL_dcolor[ch * HW + pix];
			if (a.dL_ddepth) g4[3] = a.dL_ddepth[pix];
		}
		bgterm = T * (a.bg[0] * g4[0] + a.bg[1] * g4[1] + a.bg[2] * g4[2]);
		float* grow = s_g + lane * RS;
		*reinterpret_cast<float4*>(grow + RGBD) = make_float4(g4[0], g4[1], g4[2], g4[3]);
		if (NFT > 0) {
#pragma unroll
			for (int i = 0; i < 8 * NFT; i++)
				grow[i] = (inside && i < F && a.dL_dfeature) ? a.dL_dfeature[(size_t)i * HW + pix] : 0.f;
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
	// column of pixel `lane` in the [g][pixel] tiles: pi(p) = 8 ((p >> 1) & 3) + 2 (p >> 3) + (p & 1); rows g with g & 1 set
	// have their 16-byte groups swapped pairwise (conflict-free 128-bit reads of two adjacent rows)
	const int pcol0 = 8 * ((lane >> 1) & 3) + 2 * (lane >> 3) + (lane & 1), pcol1 = pcol0 ^ 4;
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
			constexpr int NPR = 2 * NFT + 1;
			const int npieces = cnt * NPR;
			for (int idx = lane; idx < npieces; idx += 32) {
				const int r = idx / NPR, q = idx - r * NPR;
				const uint32_t id = __float_as_uint(s_q[2 * r + 1].w);
				const float4* src = (q == 2 * NFT) ? (a.rgbd + id) : (reinterpret_cast<const float4*>(a.feature + (size_t)id * F) + q);
				cp_async16(s_rows + r * RS + 4 * q, src);
			}
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
		// ---- 2. GEMM 1: Wd[g][p] = C[g][:] . Gpx[p][:]   (M = g, N = p (4 tiles), K = channels) ----
		{
			float wd[BWD_MT][4][4];
#pragma unroll
			for (int mt = 0; mt < BWD_MT; mt++)
#pragma unroll
				for (int nt = 0; nt < 4; nt++)
#pragma unroll
					for (int i = 0; i < 4; i++) wd[mt][nt][i] = 0.f;
			// k-slot (ks, tig + 4 e) of feature step ks is feature 2 NFT tig + 2 ks + e: a thread reads 2 NFT consecutive floats
			// of each of its rows; the last step is {r,g,b,depth}[tig] with the upper half of the slots empty
#pragma unroll
			for (int ks = 0; ks < NT; ks++) {
				uint32_t ahi[BWD_MT][4], alo[BWD_MT][4], bhi[4][2], blo[4][2];
#pragma unroll
				for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
					for (int h = 0; h < 2; h++) {
						const float* row = s_rows + (16 * mt + 8 * h + gid) * RS;
						float v0, v1;
						if (ks < NFT) { const float2 v = *reinterpret_cast<const float2*>(row + 2 * NFT * tig + 2 * ks); v0 = v.x; v1 = v.y; }
						else { v0 = row[RGBD + tig]; v1 = 0.f; }
						tf32_split(v0, ahi[mt][h], alo[mt][h]);
						tf32_split(v1, ahi[mt][2 + h], alo[mt][2 + h]);
					}
				}
#pragma unroll
				for (int nt = 0; nt < 4; nt++) {
					const float* row = s_g + (8 * nt + gid) * RS;
					float v0, v1;
					if (ks < NFT) { const float2 v = *reinterpret_cast<const float2*>(row + 2 * NFT * tig + 2 * ks); v0 = v.x; v1 = v.y; }
					else { v0 = row[RGBD + tig]; v1 = 0.f; }
					tf32_split(v0, bhi[nt][0], blo[nt][0]);
					tf32_split(v1, bhi[nt][1], blo[nt][1]);
				}
#pragma unroll
				for (int mt = 0; mt < BWD_MT; mt++)
#pragma unroll
					for (int nt = 0; nt < 4; nt++) mma_3xtf32(wd[mt][nt], ahi[mt], alo[mt], bhi[nt][0], bhi[nt][1], blo[nt][0], blo[nt][1]);
			}
			// D fragment (mt, nt): rows g = 16 mt + gid (+8), pixels p = 8 nt + 2 tig + e -> columns pi(p) = 8 tig + 2 nt + e
#pragma unroll
			for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int g = 16 * mt + 8 * h + gid;
					float* xrow = s_x + g * 32;
					const int sw = 4 * (g & 1);
#pragma unroll
					for (int nt = 0; nt < 4; nt++)
						*reinterpret_cast<float2*>(xrow + ((8 * tig + 2 * nt) ^ sw)) = make_float2(wd[mt][nt][2 * h], wd[mt][nt][2 * h + 1]);
				}
			}
		}
		__syncwarp();  // Wd complete; the rows buffer is free for the wgt tile
		// ---- 3. walk, back to front (queue order): lane = pixel ----
		float* s_wgt = s_rows;
#pragma unroll 4
		for (int j = 0; j < BWD_CH; j++) {
			const float4 r0 = s_q[2 * j], r1 = s_q[2 * j + 1];
			const int col = (j & 1) ? pcol1 : pcol0;
			const float wdot = s_x[j * 32 + col];
			const float dx = r0.x - pfx, dy = r0.y - pfy;
			const float ca = r0.z, cb = r0.w, cc = r1.x, op = r1.y;
			const float power = -0.5f * (ca * dx * dx + cc * dy * dy) - cb * dx * dy;
			// ex2.approx-based exp (rel. error ~1e-6); the alpha >= 1/255 decision must agree with the forward's (which uses
			// expf like the reference), so the rare borderline pairs are re-evaluated exactly
			float G = __expf(power);
			float alpha = min(ALPHA_MAX, op * G);
			if (fabsf(alpha - ALPHA_MIN) < 2e-5f * ALPHA_MIN * 8.f) {
				G = expf(power);
				alpha = min(ALPHA_MAX, op * G);
			}
			const bool valid = (j < cnt) && (__float_as_uint(r1.z) <= nc) && (power <= 0.0f) && (alpha >= ALPHA_MIN);
			const float rinv = __fdividef(1.f, 1.f - alpha);
			const float Tk = T * rinv;                                 // transmittance in front of this Gaussian
			const float wgt = alpha * Tk;                              // d(pixel channel) / d(colour of this Gaussian)
			const float dL_dalpha = Tk * wdot - (S + bgterm) * rinv;
			const float q = op * dL_dalpha * G;                        // dL/dG * G
			s_wgt[j * 32 + col] = valid ? wgt : 0.f;
			s_x[j * 32 + col] = valid ? q : 0.f;
			T = valid ? Tk : T;
			S = valid ? fmaf(wgt, wdot, S) : S;
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
		{
			// A fragments: rows g = 16 mt + 8 h + gid, logical columns 8 tig .. 8 tig + 7 = k-slots (ks, e) = (c >> 1, c & 1)
			float wv[BWD_MT][2][8], qv[BWD_MT][2][8];
#pragma unroll
			for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int g = 16 * mt + 8 * h + gid;
					const int sw = 4 * (g & 1);
#pragma unroll
					for (int u = 0; u < 2; u++) {
						const float4 w4 = *reinterpret_cast<const float4*>(s_wgt + g * 32 + ((8 * tig + 4 * u) ^ sw));
						const float4 q4 = *reinterpret_cast<const float4*>(s_x + g * 32 + ((8 * tig + 4 * u) ^ sw));
						wv[mt][h][4 * u] = w4.x; wv[mt][h][4 * u + 1] = w4.y; wv[mt][h][4 * u + 2] = w4.z; wv[mt][h][4 * u + 3] = w4.w;
						qv[mt][h][4 * u] = q4.x; qv[mt][h][4 * u + 1] = q4.y; qv[mt][h][4 * u + 2] = q4.z; qv[mt][h][4 * u + 3] = q4.w;
					}
				}
			}
#pragma unroll
			for (int ks = 0; ks < 4; ks++) {
				uint32_t ahi[BWD_MT][4], alo[BWD_MT][4], qhi[BWD_MT][4], qlo[BWD_MT][4];
#pragma unroll
				for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
					for (int h = 0; h < 2; h++) {
						tf32_split(wv[mt][h][2 * ks], ahi[mt][h], alo[mt][h]);              // k-slot (ks, tig)
						tf32_split(wv[mt][h][2 * ks + 1], ahi[mt][2 + h], alo[mt][2 + h]);  // k-slot (ks, tig + 4)
						tf32_split(qv[mt][h][2 * ks], qhi[mt][h], qlo[mt][h]);
						tf32_split(qv[mt][h][2 * ks + 1], qhi[mt][2 + h], qlo[mt][2 + h]);
					}
				}
				// B fragments of GEMM 2: cotangent rows of pixels p0 = 8 ks + 2 tig and p0 + 1
				const float* g0 = s_g + (8 * ks + 2 * tig) * RS;
				const float* g1 = g0 + RS;
				if (NFT > 0) {
					float f0[4], f1[4];
					load_feat<NFT>(g0, gid, f0);
					load_feat<NFT>(g1, gid, f1);
#pragma unroll
					for (int nt = 0; nt < NFT; nt++) {
						uint32_t bh0, bl0, bh1, bl1;
						tf32_split(f0[nt], bh0, bl0);
						tf32_split(f1[nt], bh1, bl1);
#pragma unroll
						for (int mt = 0; mt < BWD_MT; mt++) mma_3xtf32(dC[mt][nt], ahi[mt], alo[mt], bh0, bh1, bl0, bl1);
					}
				}
				{
					const float c0 = gid < 4 ? g0[RGBD + gid] : 0.f;
					const float c1 = gid < 4 ? g1[RGBD + gid] : 0.f;
					uint32_t bh0, bl0, bh1, bl1;
					tf32_split(c0, bh0, bl0);
					tf32_split(c1, bh1, bl1);
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
		}
		__syncwarp();  // every lane has read its q fragments: s_x becomes the per-Gaussian staging area
		// ---- 6. flush ----
		// feature gradients straight from the D fragments; moments and {r,g,b,depth} gradients to s_x[g][0..9]
		float* s_m = s_x;  // [g][12]
#pragma unroll
		for (int mt = 0; mt < BWD_MT; mt++) {
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int g = 16 * mt + 8 * h + gid;
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

	for (int k = 0; k < nb; k++) {
		const float4* rec4 = ring.wait(k);
		const int lo = ring.batch_lo(k), n = ring.batch_n(k);
		// chunks of the batch, back to front; survivors are appended in back-to-front order
		for (int c = ((n - 1) >> 5) << 5; c >= 0; c -= 32) {
			const int j = c + lane;
			float4 r0, r1;
			bool hit = false;
			if (j < n) {
				r0 = rec4[2 * j]; r1 = rec4[2 * j + 1];
				hit = rec_hits_block(r0, r1, fbx0, fbx1, fby0, fby1);
			}
			const uint32_t mask = __ballot_sync(0xffffffffu, hit);
			if (hit) {
				const uint32_t above = (lane == 31) ? 0u : (mask >> (lane + 1));
				const int e = qcount + __popc(above);
				r1.z = __uint_as_float((uint32_t)(lo + j) + 1u);  // the cull extent is spent: keep the 1-based list position instead
				s_q[2 * e] = r0;
				s_q[2 * e + 1] = r1;
			}
			qcount += __popc(mask);
			__syncwarp();
			while (qcount >= BWD_CH) process_chunk(BWD_CH);
		}
		// every survivor of this batch sits in the queue: refill its buffer with the batch RING ahead
		if (issued < nb) { ring.issue(issued); issued++; }
	}
	if (qcount > 0) process_chunk(qcount);
}

bool feature_rows_vectorizable(const float* feature, int F);
void launch_blend_bwd_simt(const BlendArgs& a, cudaStream_t s);

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
	if (blend_variant() & 2) { launch_blend_bwd_simt(a, s); return; }
	switch (nft_for(a.F)) {
	case 0: launch_bwd_t<0>(a, s); break;
	case 1: launch_bwd_t<1>(a, s); break;
	case 2: launch_bwd_t<2>(a, s); break;
	default: launch_bwd_t<4>(a, s); break;
	}
}

}  // namespace mgs
