// Backward blend: per-pixel back-to-front replay that emits per-Gaussian gradients.
//
// Replaces BACKWARD::render / renderCUDA<3,F> (DGR/cuda_rasterizer/backward.cu:399-593) behind the
// C-ABI.  Same mathematics: walking a pixel's contributors back to front,
//   T_k     = T_{k+1} / (1 - a_k)                      (T recovered by division from final_T, :521)
//   dL/dc_k = a_k T_k g                                 (g = dL/dpixel, all channels)
//   dL/da_k = T_k (c_k . g) - (S_k + T_final (bg . g_rgb)) / (1 - a_k),   S_k = sum_{j behind k} a_j T_j (c_j . g)
//   dL/dG   = o_k dL/da_k  (the 0.99 clamp is gradient-transparent, :574), then mean2D / conic / opacity.
// S_k is the scalar form of the reference's per-channel `accum_rec` recursion: sum_ch (c - A_k) g T_k with
// A_k = S_k-vector / T_{k+1}.  Using the scalar keeps the per-pixel carried state to two floats.
//
// B200 design: GAUSSIAN-parallel inside a single-warp CTA (blend_common.cuh).  The reference has every
// pixel-thread issue 9+F global float atomics per contributing pair (41 at F=32).  Here a warp owns a BWD_BW x BWD_BH
// pixel block; it streams the first max(n_contrib) records of its tile back to front through its TMA ring, culls them
// against the block and queues the survivors; whenever 32 are queued, lane l takes the l-th one and keeps its
// channel row and ALL of its gradient accumulators in registers while the warp walks the block's pixels.  The per-pixel
// sequential dependences (transmittance, S) across the 32 Gaussians of a chunk are resolved with warp prefix
// scans.  No cross-lane reduction of the 9+F gradients is needed, and each (block, Gaussian) pair costs
// ceil((12+F)/4) 128-bit red.global.add.v4.f32 instead of 9+F scalar atomics per pixel.  Pixel cotangent rows are
// 128-bit shared-memory broadcasts; each lane fetches its Gaussian's channel row with 128-bit read-only loads.
#include "blend_common.cuh"

namespace mgs {
static_assert(REC_BATCH == 64, "the round-1 SIMT blends assume 64-record batches");

#ifndef MGS_BWD_BW
#define MGS_BWD_BW 4
#endif
#ifndef MGS_BWD_BH
#define MGS_BWD_BH 4
#endif
// Pixel block of one backward CTA.  The pixel walk costs (survivors of the block) x (pixels of the block), so for splats
// of a few pixels a smaller block wastes less work on (Gaussian, pixel) pairs that do not touch; the price is one more
// cull of the tile's list and one more gradient flush per (Gaussian, block) pair.  Measured on B200 at c3 (ms per view):
// 8x4 0.641, 4x4 0.533, 8x2 0.582, 4x2 0.572, 2x2 0.768.
constexpr int BWD_BW = MGS_BWD_BW, BWD_BH = MGS_BWD_BH;
constexpr int BWD_NPX = BWD_BW * BWD_BH;                            // pixels per CTA (<= 32)
constexpr int BWD_SUBS_X = TILE_X / BWD_BW, BWD_SUBS = BWD_SUBS_X * (TILE_Y / BWD_BH);  // CTAs per 16x16 tile
static_assert(BWD_NPX <= 32 && TILE_X % BWD_BW == 0 && TILE_Y % BWD_BH == 0, "backward pixel block");
constexpr int QCAP = 64;  // survivor queue capacity (power of two, >= 63)
#ifndef MGS_BWD_PIX
#define MGS_BWD_PIX 2
#endif
#ifndef MGS_BWD_PREDICATED
#define MGS_BWD_PREDICATED 0
#endif
constexpr int PIX = MGS_BWD_PIX;  // pixels interleaved per iteration of the pixel walk (instruction-level parallelism)
#ifdef MGS_BWD_MIN_CTAS
constexpr int BWD_MIN_CTAS = MGS_BWD_MIN_CTAS;
#else
constexpr int BWD_MIN_CTAS = PIX == 4 ? 8 : (PIX == 2 ? 12 : 16);
#endif

template <int NQ, bool VEC>
__global__ void __launch_bounds__(32, BWD_MIN_CTAS) blend_bwd_simt_kernel(BlendArgs a)
{
	__shared__ __align__(128) InstRec s_rec[RING * REC_BATCH];
	__shared__ __align__(16) float4 s_queue[QCAP * 2];   // survivors: {x,y,ca,cb}, {cc,op,pos,id}
	__shared__ __align__(16) float4 s_g[BWD_NPX * NQ];    // [pixel][q] cotangent rows
	__shared__ __align__(16) float4 s_state[BWD_NPX];     // {T, S, n_contrib, Tfinal*bg.g}
	__shared__ __align__(8) uint64_t s_bar[RING];

	const int lane = threadIdx.x;
	const int tile = blockIdx.x / BWD_SUBS, sub = blockIdx.x % BWD_SUBS;
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int bx0 = tile_x * TILE_X + (sub % BWD_SUBS_X) * BWD_BW;
	const int by0 = tile_y * TILE_Y + (sub / BWD_SUBS_X) * BWD_BH;
	const int pxi = bx0 + (lane % BWD_BW), pyi = by0 + (lane / BWD_BW);
	const bool inside = lane < BWD_NPX && pxi < a.W && pyi < a.H;
	const float fbx0 = (float)bx0, fbx1 = (float)(bx0 + BWD_BW - 1), fby0 = (float)by0, fby1 = (float)(by0 + BWD_BH - 1);
	const size_t HW = (size_t)a.H * a.W;
	const size_t pix = (size_t)a.W * pyi + pxi;
	const int F = a.F;

	// ---- per-pixel cotangent rows and carried state (lane == pixel here) ----
	uint32_t nc = 0;
	{
		float g[4 * NQ];
#pragma unroll
		for (int i = 0; i < 4 * NQ; i++) g[i] = 0.f;
		float Tf = 0.f;
		if (inside) {
			nc = a.n_contrib[pix];
			Tf = a.final_T[pix];
#pragma unroll
			for (int ch = 0; ch < 3; ch++) g[ch] = a.dL_dcolor[ch * HW + pix];
			if (a.dL_ddepth) g[3] = a.dL_ddepth[pix];
			if (NQ > 1 && a.dL_dfeature) {
#pragma unroll
				for (int i = 0; i < 4 * (NQ - 1); i++)
					if (i < F) g[4 + i] = a.dL_dfeature[(size_t)i * HW + pix];
			}
		}
		if (lane < BWD_NPX) {
#pragma unroll
			for (int q = 0; q < NQ; q++) s_g[lane * NQ + q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
			const float bgdot = a.bg[0] * g[0] + a.bg[1] * g[1] + a.bg[2] * g[2];
			s_state[lane] = make_float4(Tf, 0.f, __uint_as_float(nc), Tf * bgdot);
		}
	}
	uint32_t maxc = nc;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
	if (maxc == 0) return;  // nothing blended into this block
	__syncwarp();

	const uint2 range = a.ranges[tile];
	WarpRecRing ring;
	// only the first maxc instances of the tile's list can have contributed to a pixel of this block
	ring.init(s_rec, s_bar, a.recs + range.x, (int)maxc, true);
	const int nb = ring.num_batches();
	int issued = 0;
	for (; issued < min(nb, RING); issued++) ring.issue(issued);

	const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
	int qhead = 0, qcount = 0;  // survivor queue (circular), farthest-from-camera first

	// Consume up to 32 queued survivors: lane l <- survivor l.
	auto process_chunk = [&](int cnt) {
		const bool have = lane < cnt;
		float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
		if (have) {
			const int e = (qhead + lane) & (QCAP - 1);
			r0 = s_queue[2 * e]; r1 = s_queue[2 * e + 1];
		}
		const uint32_t id = __float_as_uint(r1.w);
		const uint32_t pos = __float_as_uint(r1.z);
		float c[4 * NQ], dch[4 * NQ];
#pragma unroll
		for (int i = 0; i < 4 * NQ; i++) { c[i] = 0.f; dch[i] = 0.f; }
		if (have) {
			const float4 v = ldg_nc_v4(a.rgbd + id);
			c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
			if (NQ > 1) {
				if (VEC) {
					const float4* row = reinterpret_cast<const float4*>(a.feature + (size_t)id * F);
#pragma unroll
					for (int q = 1; q < NQ; q++) {
						if (4 * (q - 1) < F) {
							const float4 u = ldg_nc_v4(row + (q - 1));
							c[4 * q] = u.x; c[4 * q + 1] = u.y; c[4 * q + 2] = u.z; c[4 * q + 3] = u.w;
						}
					}
				} else {
					const float* row = a.feature + (size_t)id * F;
#pragma unroll
					for (int i = 0; i < 4 * (NQ - 1); i++)
						if (i < F) c[4 + i] = __ldg(row + i);
				}
			}
		}
		const float gx_ = r0.x, gy_ = r0.y, ca = r0.z, cb = r0.w, cc = r1.x, op = r1.y;
		float dmx = 0.f, dmy = 0.f, dca = 0.f, dcb = 0.f, dcc = 0.f, dop = 0.f;

		// Two horizontally adjacent pixels per iteration: their scan chains (5 dependent shuffles each, twice) are
		// independent, so interleaving them doubles the instruction-level parallelism of the latency-bound part.
		for (int p = 0; p < BWD_NPX; p += PIX) {
			float4 st[PIX];
			uint32_t ncp[PIX];
			bool live_px = false;
#pragma unroll
			for (int u = 0; u < PIX; u++) {
				st[u] = s_state[p + u];  // broadcast
				ncp[u] = __float_as_uint(st[u].z);
				live_px |= ncp[u] != 0;
			}
			if (!live_px) continue;  // uniform: pixels outside the image or without contributors
			float dx[PIX], dy[PIX], G[PIX], alpha[PIX];
			bool valid[PIX];
			bool any_valid = false;
#pragma unroll
			for (int u = 0; u < PIX; u++) {
				const float pfx = (float)(bx0 + ((p + u) % BWD_BW)), pfy = (float)(by0 + ((p + u) / BWD_BW));
				dx[u] = gx_ - pfx; dy[u] = gy_ - pfy;
				const float power = -0.5f * (ca * dx[u] * dx[u] + cc * dy[u] * dy[u]) - cb * dx[u] * dy[u];
				// ex2.approx-based exp (rel. error ~1e-6); the alpha >= 1/255 decision must agree with the forward's
				// (which uses expf like the reference), so the rare borderline pairs are re-evaluated exactly
				G[u] = __expf(power);
				alpha[u] = min(ALPHA_MAX, op * G[u]);
				if (fabsf(alpha[u] - ALPHA_MIN) < 2e-5f * ALPHA_MIN * 8.f) {
					G[u] = expf(power);
					alpha[u] = min(ALPHA_MAX, op * G[u]);
				}
				valid[u] = have && (pos <= ncp[u]) && (power <= 0.0f) && (alpha[u] >= ALPHA_MIN);
				any_valid |= valid[u];
			}
			if (!__any_sync(0xffffffffu, any_valid)) continue;
			// inclusive product scans of (1 - alpha) over the chunk, lane 0 = farthest from the camera
			float ip[PIX];
#pragma unroll
			for (int u = 0; u < PIX; u++) ip[u] = valid[u] ? (1.f - alpha[u]) : 1.f;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
				for (int u = 0; u < PIX; u++) {
					const float v = __shfl_up_sync(0xffffffffu, ip[u], o);
					if (lane >= o) ip[u] *= v;
				}
			}
			float Tk[PIX], wgt[PIX], w[PIX];
#pragma unroll
			for (int u = 0; u < PIX; u++) {
				Tk[u] = __fdividef(st[u].x, ip[u]);            // transmittance in front of this Gaussian
				wgt[u] = valid[u] ? alpha[u] * Tk[u] : 0.f;    // dchannel_dcolor
				w[u] = 0.f;
			}
			// channel work: w = c_j . g_p ; dL/dc_j += wgt * g_p
#pragma unroll
			for (int q = 0; q < NQ; q++) {
#pragma unroll
				for (int u = 0; u < PIX; u++) {
					const float4 g = s_g[(p + u) * NQ + q];
					w[u] += c[4 * q] * g.x; w[u] += c[4 * q + 1] * g.y; w[u] += c[4 * q + 2] * g.z; w[u] += c[4 * q + 3] * g.w;
					dch[4 * q] += wgt[u] * g.x; dch[4 * q + 1] += wgt[u] * g.y; dch[4 * q + 2] += wgt[u] * g.z; dch[4 * q + 3] += wgt[u] * g.w;
				}
			}
			// S_k = carried S + contributions of the lanes behind me in this chunk (exclusive prefix sum)
			float xk[PIX], is[PIX];
#pragma unroll
			for (int u = 0; u < PIX; u++) { xk[u] = wgt[u] * w[u]; is[u] = xk[u]; }
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
				for (int u = 0; u < PIX; u++) {
					const float v = __shfl_up_sync(0xffffffffu, is[u], o);
					if (lane >= o) is[u] += v;
				}
			}
#pragma unroll
			for (int u = 0; u < PIX; u++) {
				const float Sk = st[u].y + (is[u] - xk[u]);
				const float Tnew = __shfl_sync(0xffffffffu, Tk[u], 31);
				const float Snew = st[u].y + __shfl_sync(0xffffffffu, is[u], 31);
				if (lane == 0) s_state[p + u] = make_float4(Tnew, Snew, st[u].z, st[u].w);
#if MGS_BWD_PREDICATED
				// EXPERIMENT (off; not yet measured on a GPU): ~10 of 32 lanes are valid here on the benchmark cloud, so the
				// divergent region below costs a reconvergence per pixel; this form computes the addends on every lane and
				// selects them away on lanes that are not valid (a select, not a multiply: G can be inf where power > 0).
				{
					const float dL_dalpha = Tk[u] * w[u] - __fdividef(Sk + st[u].w, 1.f - alpha[u]);
					const float dL_dG = op * dL_dalpha;
					const float gdx = G[u] * dx[u], gdy = G[u] * dy[u];
					const float dG_ddelx = -gdx * ca - gdy * cb;
					const float dG_ddely = -gdy * cc - gdx * cb;
					dmx += valid[u] ? dL_dG * dG_ddelx * ddelx_dx : 0.f;
					dmy += valid[u] ? dL_dG * dG_ddely * ddely_dy : 0.f;
					dca += valid[u] ? -0.5f * gdx * dx[u] * dL_dG : 0.f;
					dcb += valid[u] ? -0.5f * gdx * dy[u] * dL_dG : 0.f;
					dcc += valid[u] ? -0.5f * gdy * dy[u] * dL_dG : 0.f;
					dop += valid[u] ? G[u] * dL_dalpha : 0.f;
				}
#else
				if (valid[u]) {
					const float dL_dalpha = Tk[u] * w[u] - __fdividef(Sk + st[u].w, 1.f - alpha[u]);
					const float dL_dG = op * dL_dalpha;
					const float gdx = G[u] * dx[u], gdy = G[u] * dy[u];
					const float dG_ddelx = -gdx * ca - gdy * cb;
					const float dG_ddely = -gdy * cc - gdx * cb;
					dmx += dL_dG * dG_ddelx * ddelx_dx;
					dmy += dL_dG * dG_ddely * ddely_dy;
					dca += -0.5f * gdx * dx[u] * dL_dG;
					dcb += -0.5f * gdx * dy[u] * dL_dG;
					dcc += -0.5f * gdy * dy[u] * dL_dG;
					dop += G[u] * dL_dalpha;
				}
#endif
			}
		}
		__syncwarp();

		// ---- flush this Gaussian's gradients: 128-bit reductions to L2 ----
		if (have) {
			float* gb = a.gb + (size_t)id * GB_STRIDE;
			red_add_v4(gb, dmx, dmy, dca, dcb);
			red_add_v4(gb + 4, dcc, dop, dch[0], dch[1]);
			red_add_v4(gb + 8, dch[2], dch[3], 0.f, 0.f);
			if (NQ > 1 && a.dL_dfeat) {
				float* df = a.dL_dfeat + (size_t)id * F;
				if (VEC) {
#pragma unroll
					for (int q = 1; q < NQ; q++)
						if (4 * (q - 1) < F) red_add_v4(df + 4 * (q - 1), dch[4 * q], dch[4 * q + 1], dch[4 * q + 2], dch[4 * q + 3]);
				} else {
#pragma unroll
					for (int i = 0; i < 4 * (NQ - 1); i++)
						if (i < F) red_add(df + i, dch[4 + i]);
				}
			}
		}
		qhead = (qhead + cnt) & (QCAP - 1);
		qcount -= cnt;
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
				const int e = (qhead + qcount + __popc(above)) & (QCAP - 1);
				r1.z = __uint_as_float((uint32_t)(lo + j) + 1u);  // the cull extent is spent: keep the 1-based list position instead
				s_queue[2 * e] = r0;
				s_queue[2 * e + 1] = r1;
			}
			qcount += __popc(mask);
			__syncwarp();
			if (qcount >= 32) process_chunk(32);
		}
		// this batch's buffer is dead: refill it with the batch RING ahead
		if (issued < nb) { ring.issue(issued); issued++; }
	}
	if (qcount > 0) process_chunk(qcount);
}

bool feature_rows_vectorizable(const float* feature, int F);

template <int NQ>
static void launch_bwd_t(const BlendArgs& a, cudaStream_t s)
{
	const int grid = a.grid_x * a.grid_y * BWD_SUBS;
	const bool vec = NQ == 1 || (feature_rows_vectorizable(a.feature, a.F) && (reinterpret_cast<uintptr_t>(a.dL_dfeat) & 15) == 0);
	if (vec) blend_bwd_simt_kernel<NQ, true><<<grid, 32, 0, s>>>(a);
	else blend_bwd_simt_kernel<NQ, false><<<grid, 32, 0, s>>>(a);
}

void launch_blend_bwd_simt(const BlendArgs& a, cudaStream_t s)
{
	switch (a.nq) {
	case 1: launch_bwd_t<1>(a, s); break;
	case 2: launch_bwd_t<2>(a, s); break;
	case 3: launch_bwd_t<3>(a, s); break;
	case 5: launch_bwd_t<5>(a, s); break;
	default: launch_bwd_t<9>(a, s); break;
	}
}

}  // namespace mgs
