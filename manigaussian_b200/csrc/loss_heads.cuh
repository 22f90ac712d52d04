// Loss heads of ManiGaussian's neural-rendering objective, fused next to the rasterizer.
//
// Replaces, for the images the forward blend has just produced, the PyTorch operator chains of
//   l2_loss(render, gt)       = mean((render - gt)^2)                          agents/manigaussian_bc/loss.py:12-13
//   cosine_loss(embed, gt)    = 1 - mean_px cos_sim(embed[px, :], gt[px, :])   loss.py:18-23 (F.cosine_similarity, eps 1e-8)
// as NeuralRenderer.forward applies them (neural_rendering.py:300-318), together with their backward: per pixel the
// kernel produces the two loss partial sums and the cotangent planes dL/d(render), dL/d(embed) the backward blend reads,
// so that no image-sized PyTorch kernel runs between the forward and the backward of a training step.
//   S_rgb = sum (x - g)^2        ->  loss_rgb   = S_rgb / (3 N),     d loss_rgb / dx_c   = 2 (x_c - g_c) / (3 N)
//   S_cos = sum_px cos_px        ->  loss_embed = 1 - S_cos / N,     d loss_embed / dx_f = -(1/N) d cos / dx_f
//   cos = x.y / (max(|x|, eps) max(|y|, eps));  d cos / dx = y / (nx ny) - (x.y) x / (|x|^2 nx ny)   (second term only
//   where |x| > eps: below it the clamped norm is a constant, as in ATen's backward of clamp_min)
#pragma once
#include "mgs_common.cuh"

namespace mgs {

constexpr float COS_EPS = 1e-8f;  // F.cosine_similarity's default eps

struct CosTerms {
	float inv_nn;   // 1 / (max(|x|, eps) max(|y|, eps))
	float k;        // (x.y) / |x|^2 * inv_nn where |x| > eps, else 0
	float cos;
};
__device__ __forceinline__ CosTerms cos_terms(float xy, float xx, float yy)
{
	const float nx = sqrtf(xx), ny = sqrtf(yy);
	CosTerms t;
	t.inv_nn = 1.0f / (fmaxf(nx, COS_EPS) * fmaxf(ny, COS_EPS));
	t.cos = xy * t.inv_nn;
	t.k = nx > COS_EPS ? t.cos / xx : 0.f;
	return t;
}
// d cos / dx_f
__device__ __forceinline__ float cos_grad(const CosTerms& t, float x, float y) { return y * t.inv_nn - t.k * x; }

}  // namespace mgs
