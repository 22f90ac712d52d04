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
// B200 design: GAUSSIAN-parallel inside a warp.  The reference has every pixel-thread issue 9+F global
// float atomics per contributing pair (41 at F=32).  Here a warp owns an 8x4 pixel block; after culling
// the staged work list against that block, lane l takes the l-th surviving Gaussian and keeps its channel
// row and ALL of its gradient accumulators in registers; the warp then walks its 32 pixels, and the
// per-pixel sequential dependences (transmittance, S) across the 32 Gaussians of the chunk are resolved
// with warp prefix scans.  No cross-lane reduction of the 9+F gradients is needed, and each (block,
// Gaussian) pair costs ceil((12+F)/4) 128-bit red.global.add.v4.f32 instead of 9+F scalar atomics per
// pixel.  Pixel cotangent rows are read as 128-bit shared-memory broadcasts; records arrive through the
// TMA ring of blend_common.cuh; each lane fetches its Gaussian's channel row with 128-bit loads.
#include "blend_common.cuh"

namespace mgs {

template <int NQ, bool VEC>
__global__ void __launch_bounds__(BLEND_THREADS, 2) blend_bwd_kernel(BlendArgs a)
{
	constexpr int NW = BLEND_THREADS / 32;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	unsigned char* sp = smem_raw;
	InstRec* s_rec = reinterpret_cast<InstRec*>(sp); sp += (size_t)RING * BATCH * sizeof(InstRec);
	float4* s_g = reinterpret_cast<float4*>(sp); sp += (size_t)NW * 32 * NQ * sizeof(float4);   // [warp][pixel][q]
	float4* s_state = reinterpret_cast<float4*>(sp); sp += (size_t)NW * 32 * sizeof(float4);     // {T, S, n_contrib, Tfinal*bg.g}
	uint16_t* s_hit = reinterpret_cast<uint16_t*>(sp); sp += (size_t)NW * BATCH * sizeof(uint16_t);
	__shared__ __align__(8) uint64_t s_bar[RING];
	__shared__ uint32_t s_tile_max;

	const int tile = blockIdx.x;
	const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int bx0 = tile_x * TILE_X + (warp & 1) * WARP_BX;
	const int by0 = tile_y * TILE_Y + (warp >> 1) * WARP_BY;
	const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
	const bool inside = pxi < a.W && pyi < a.H;
	const float fbx0 = (float)bx0, fbx1 = (float)(bx0 + WARP_BX - 1), fby0 = (float)by0, fby1 = (float)(by0 + WARP_BY - 1);
	const size_t HW = (size_t)a.H * a.W;
	const size_t pix = (size_t)a.W * pyi + pxi;
	const int F = a.F;

	RecRing ring;
	ring.init(s_rec, s_bar, a.recs);
	if (threadIdx.x == 0) s_tile_max = 0;
	const uint2 range = a.ranges[tile];

	// ---- per-pixel cotangent rows and carried state (lane == pixel here) ----
	float4* my_g = s_g + (size_t)(warp * 32 + lane) * NQ;
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
#pragma unroll
		for (int q = 0; q < NQ; q++) my_g[q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
		const float bgdot = a.bg[0] * g[0] + a.bg[1] * g[1] + a.bg[2] * g[2];
		s_state[warp * 32 + lane] = make_float4(Tf, 0.f, __uint_as_float(nc), Tf * bgdot);
	}
	uint32_t maxc = nc;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
	__syncthreads();  // barrier init + s_tile_max = 0 visible
	if (lane == 0 && maxc > 0) atomicMax(&s_tile_max, maxc);
	__syncthreads();
	const int tile_max = (int)s_tile_max;
	if (tile_max == 0) return;

	uint16_t* my_hit = s_hit + warp * BATCH;
	const float4* wg = s_g + (size_t)warp * 32 * NQ;
	float4* wstate = s_state + warp * 32;
	const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

	// Only the first tile_max instances of the tile can have contributed to any pixel.  Batches are walked from
	// the back: batch k covers list positions [lo_k, lo_k + n_k), lo_k = BATCH * (nb - 1 - k).
	const int nb = (tile_max + BATCH - 1) / BATCH;
	auto batch_lo = [&](int k) { return BATCH * (nb - 1 - k); };
	auto batch_n = [&](int k) { return min(BATCH, tile_max - batch_lo(k)); };
	if (threadIdx.x == 0) {
		for (int k = 0; k < min(nb, RING - 1); k++) ring.issue(k, range.x + batch_lo(k), batch_n(k));
	}

	for (int k = 0; k < nb; k++) {
		__syncthreads();  // batch k-1 fully consumed: its buffer may be refilled
		if (threadIdx.x == 0 && k + RING - 1 < nb) {
			const int kk = k + RING - 1;
			ring.issue(kk, range.x + batch_lo(kk), batch_n(kk));
		}
		const float4* rec4 = ring.wait(k);
		if (maxc == 0) continue;
		const int lo = batch_lo(k), n = batch_n(k);

		// ---- cull against this warp's 8x4 block; hit list in back-to-front order ----
		int nh = 0;
		for (int c = ((n - 1) >> 5) << 5; c >= 0; c -= 32) {
			const int j = c + lane;
			bool hit = false;
			if (j < n) hit = ((uint32_t)(lo + j) < maxc) && rec_hits_block(rec4[2 * j], rec4[2 * j + 1], fbx0, fbx1, fby0, fby1);
			const uint32_t mask = __ballot_sync(0xffffffffu, hit);
			if (hit) {
				const uint32_t above = (lane == 31) ? 0u : (mask >> (lane + 1));
				my_hit[nh + __popc(above)] = (uint16_t)j;
			}
			nh += __popc(mask);
		}
		__syncwarp();

		for (int k0 = 0; k0 < nh; k0 += 32) {
			const int cnt = min(32, nh - k0);
			const bool have = lane < cnt;
			// ---- lane <- Gaussian: record, channel row, zeroed accumulators ----
			int jj = 0;
			if (have) jj = my_hit[k0 + lane];
			float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
			if (have) { r0 = rec4[2 * jj]; r1 = rec4[2 * jj + 1]; }
			const uint32_t id = rec_id(r1);
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
			const uint32_t pos = (uint32_t)(lo + jj) + 1u;
			const float gx_ = r0.x, gy_ = r0.y, ca = r0.z, cb = r0.w, cc = r1.x, op = r1.y;
			float dmx = 0.f, dmy = 0.f, dca = 0.f, dcb = 0.f, dcc = 0.f, dop = 0.f;

			for (int p = 0; p < 32; p++) {
				const float4 st = wstate[p];  // broadcast
				const uint32_t ncp = __float_as_uint(st.z);
				if (ncp == 0) continue;  // uniform: pixel outside the image or without contributors
				const float pfx = (float)(bx0 + (p & 7)), pfy = (float)(by0 + (p >> 3));
				const float dx = gx_ - pfx, dy = gy_ - pfy;
				const float power = -0.5f * (ca * dx * dx + cc * dy * dy) - cb * dx * dy;
				const float G = expf(power);
				const float alpha = min(ALPHA_MAX, op * G);
				const bool valid = have && (pos <= ncp) && (power <= 0.0f) && (alpha >= ALPHA_MIN);
				if (!__any_sync(0xffffffffu, valid)) continue;
				const float om = valid ? (1.f - alpha) : 1.f;
				// inclusive product scan of (1 - alpha) over the chunk, lane 0 = farthest from the camera
				float ip = om;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const float v = __shfl_up_sync(0xffffffffu, ip, o);
					if (lane >= o) ip *= v;
				}
				const float Tk = st.x / ip;                       // transmittance in front of this Gaussian
				const float Tnew = __shfl_sync(0xffffffffu, Tk, 31);
				const float wgt = valid ? alpha * Tk : 0.f;        // dchannel_dcolor

				// channel work: w = c_j . g_p ; dL/dc_j += wgt * g_p
				float w = 0.f;
				const float4* gp = wg + (size_t)p * NQ;
#pragma unroll
				for (int q = 0; q < NQ; q++) {
					const float4 g = gp[q];
					w += c[4 * q] * g.x; w += c[4 * q + 1] * g.y; w += c[4 * q + 2] * g.z; w += c[4 * q + 3] * g.w;
					dch[4 * q] += wgt * g.x; dch[4 * q + 1] += wgt * g.y; dch[4 * q + 2] += wgt * g.z; dch[4 * q + 3] += wgt * g.w;
				}
				// S_k = carried S + contributions of the lanes behind me in this chunk (exclusive prefix sum)
				const float xk = wgt * w;
				float is = xk;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const float v = __shfl_up_sync(0xffffffffu, is, o);
					if (lane >= o) is += v;
				}
				const float Sk = st.y + (is - xk);
				const float Snew = st.y + __shfl_sync(0xffffffffu, is, 31);
				if (lane == 0) wstate[p] = make_float4(Tnew, Snew, st.z, st.w);

				if (valid) {
					const float dL_dalpha = Tk * w - (Sk + st.w) / (1.f - alpha);
					const float dL_dG = op * dL_dalpha;
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * ca - gdy * cb;
					const float dG_ddely = -gdy * cc - gdx * cb;
					dmx += dL_dG * dG_ddelx * ddelx_dx;
					dmy += dL_dG * dG_ddely * ddely_dy;
					dca += -0.5f * gdx * dx * dL_dG;
					dcb += -0.5f * gdx * dy * dL_dG;
					dcc += -0.5f * gdy * dy * dL_dG;
					dop += G * dL_dalpha;
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
		}
	}
}

bool feature_rows_vectorizable(const float* feature, int F);

static size_t bwd_smem_bytes(int nq)
{
	constexpr int NW = BLEND_THREADS / 32;
	return (size_t)RING * BATCH * sizeof(InstRec) + (size_t)NW * 32 * nq * sizeof(float4) + (size_t)NW * 32 * sizeof(float4) +
		(size_t)NW * BATCH * sizeof(uint16_t);
}

template <int NQ, bool VEC>
static void launch_bwd_tv(const BlendArgs& a, cudaStream_t s)
{
	const size_t smem = bwd_smem_bytes(NQ);
	static bool configured = false;
	if (!configured) {
		cudaFuncSetAttribute(blend_bwd_kernel<NQ, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = true;
	}
	blend_bwd_kernel<NQ, VEC><<<a.grid_x * a.grid_y, BLEND_THREADS, smem, s>>>(a);
}

template <int NQ>
static void launch_bwd_t(const BlendArgs& a, cudaStream_t s)
{
	const bool vec = NQ == 1 || (feature_rows_vectorizable(a.feature, a.F) && (reinterpret_cast<uintptr_t>(a.dL_dfeat) & 15) == 0);
	if (vec) launch_bwd_tv<NQ, true>(a, s);
	else launch_bwd_tv<NQ, false>(a, s);
}

void launch_blend_bwd(const BlendArgs& a, cudaStream_t s)
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
