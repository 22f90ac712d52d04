// Shared by the forward and backward blend kernels: a 3-deep ring of shared-memory buffers into which one
// elected thread streams a tile's sorted 32-byte instance records with 1-D TMA bulk copies (UBLKCP), one
// copy per batch of up to BATCH records, completion tracked by one mbarrier per buffer.  Records of batch
// k+2 are in flight while batch k is consumed, so the only per-batch synchronisation is the CTA barrier that
// retires batch k-1's buffer.
//
// Channel rows ({r,g,b,depth} float4 + F feature floats per Gaussian) are NOT staged: they are read from
// global memory with 128-bit loads through the read-only path -- warp-uniform addresses in the forward
// (one transaction per load, L1-resident across the 8 warps of a tile), per-lane rows in the backward.
#pragma once
#include <cuda_fp16.h>
#include "mgs_common.cuh"
#include "mgs_kernels.h"

namespace mgs {

constexpr int BLEND_THREADS = 256;
constexpr int BATCH = 256;  // records per bulk copy (8 KB)
constexpr int RING = 3;

struct RecRing {
	InstRec* buf;      // RING * BATCH records
	uint64_t* bar;     // RING mbarriers
	const InstRec* src;

	__device__ __forceinline__ void init(InstRec* b, uint64_t* bars, const InstRec* s)
	{
		buf = b; bar = bars; src = s;
		if (threadIdx.x == 0) {
#pragma unroll
			for (int i = 0; i < RING; i++) mbar_init(&bar[i], 1);
			mbar_fence_init();
		}
	}
	// thread 0 only: start the copy of records [lo, lo+n) into buffer k % RING
	__device__ __forceinline__ void issue(int k, uint32_t lo, int n)
	{
		const int b = k % RING;
		fence_proxy_async();
		mbar_arrive_expect_tx(&bar[b], (uint32_t)n * (uint32_t)sizeof(InstRec));
		bulk_g2s(buf + b * BATCH, src + lo, (uint32_t)n * (uint32_t)sizeof(InstRec), &bar[b]);
	}
	// all threads: wait until batch k has landed; returns its buffer
	__device__ __forceinline__ const float4* wait(int k)
	{
		const int b = k % RING;
		mbar_wait(&bar[b], (uint32_t)((k / RING) & 1));
		return reinterpret_cast<const float4*>(buf + b * BATCH);
	}
};

// {x, y, ca, cb} and {cc, op, ext(half2 hx,hy), id} views of a record
__device__ __forceinline__ float2 rec_extent(const float4& r1)
{
	return __half22float2(*reinterpret_cast<const __half2*>(&r1.z));
}
__device__ __forceinline__ uint32_t rec_id(const float4& r1) { return __float_as_uint(r1.w); }

// does the alpha >= 1/255 footprint of the record overlap the pixel block [x0,x1] x [y0,y1]?
__device__ __forceinline__ bool rec_hits_block(const float4& r0, const float4& r1, float x0, float x1, float y0, float y1)
{
	const float2 e = rec_extent(r1);
	return (e.x >= 0.f) && (r0.x + e.x >= x0) && (r0.x - e.x <= x1) && (r0.y + e.y >= y0) && (r0.y - e.y <= y1);
}

}  // namespace mgs
