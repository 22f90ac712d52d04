// C-ABI entry points (include/mgs_rasterizer.h): orchestration of the per-view forward and backward.
//
// Replaces CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (DGR/cuda_rasterizer/rasterizer_impl.cu:198-355, :359-463, :141-153).  Stage order and the meaning of
// every argument follow the reference; the state-buffer layout, kernels and stream handling are ours.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/mgs_rasterizer.h"
#include "mgs_common.cuh"
#include "mgs_kernels.h"

namespace mgs {

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
	g_err = msg;
	return code;
}

#define MGS_CUDA(call)                                                                                   \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess)                                                                           \
			return fail(MGS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
	} while (0)

// after each stage: launch errors always, execution errors when debug (the reference's CHECK_CUDA, auxiliary.h:166-173)
static const bool g_trace = getenv("MGS_TRACE") != nullptr;  // diagnostic: name every stage on stderr as it is enqueued
#define MGS_STAGE(name)                                                                                  \
	do {                                                                                                 \
		if (g_trace) { fprintf(stderr, "[mgs] stage %s enqueued\n", name); fflush(stderr); }            \
		cudaError_t e_ = cudaGetLastError();                                                             \
		if (e_ == cudaSuccess && debug) e_ = cudaStreamSynchronize(st);                                  \
		if (e_ != cudaSuccess)                                                                           \
			return fail(MGS_ERR_CUDA, std::string("stage ") + name + ": " + cudaGetErrorString(e_));     \
	} while (0)

// ---- optional per-stage CUDA-event timing (bench.py's roofline leg); off by default, zero cost when off ----
enum Stage { ST_PROJECT_FWD = 0, ST_DEPTH_SORT, ST_SCAN, ST_EMIT, ST_SORT, ST_RANGES_PACK, ST_BLEND_FWD, ST_BLEND_BWD, ST_PROJECT_BWD, ST_ACTIVATE_FWD, ST_ACTIVATE_BWD, ST_COUNT };
struct StageEvt { int stage; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<StageEvt> g_prof_evts;
struct StageTimer {
	cudaStream_t st; bool on; StageEvt e;
	StageTimer(int stage, cudaStream_t s) : st(s), on(g_prof_on)
	{
		if (!on) return;
		e.stage = stage;
		cudaEventCreate(&e.a); cudaEventCreate(&e.b);
		cudaEventRecord(e.a, st);
	}
	~StageTimer()
	{
		if (!on) return;
		cudaEventRecord(e.b, st);
		std::lock_guard<std::mutex> lk(g_prof_mu);
		g_prof_evts.push_back(e);
	}
};

// MGS_TILE_ORDER=0 launches the blend CTAs in tile order instead of longest-list-first (A/B switch for measurements)
static bool use_tile_order()
{
	static const bool on = [] { const char* e = getenv("MGS_TILE_ORDER"); return !(e && e[0] == '0'); }();
	return on;
}

#ifdef MGS_CTA_LOG
static unsigned long long* g_cta_log = nullptr;
static unsigned int* g_cta_log_n = nullptr;
static unsigned int g_cta_log_cap = 0;
extern "C" int mgs_debug_set_cta_log(void* log, void* counter, unsigned int cap)
{
	g_cta_log = static_cast<unsigned long long*>(log); g_cta_log_n = static_cast<unsigned int*>(counter); g_cta_log_cap = cap;
	return 0;
}
#define MGS_CTA_LOG_ARGS(ba) do { (ba).cta_log = g_cta_log; (ba).cta_log_n = g_cta_log_n; (ba).cta_log_cap = g_cta_log_cap; } while (0)
#else
#define MGS_CTA_LOG_ARGS(ba) do { } while (0)
#endif

template <typename T>
static void obtain(char*& chunk, T*& ptr, size_t count, size_t alignment = 128)
{
	size_t offset = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(alignment - 1);
	ptr = reinterpret_cast<T*>(offset);
	chunk = reinterpret_cast<char*>(ptr + count);
}

// CUB temp-storage sizes depend on the item count and the device; the queries are not free (device attribute lookups), so
// cache them.  Sizes are taken for the item count rounded up to the next multiple of 1/8 of its power of two (monotone in
// n, few distinct keys, at most 12.5 % above the exact need).
static size_t cached_temp_bytes(int which, int n)
{
	static std::mutex mu;
	static std::map<std::tuple<int, int, int>, size_t> cache;
	int p2 = 1024;
	while (p2 < n && p2 < (1 << 30)) p2 <<= 1;
	const int q = std::max(128, p2 >> 3);
	const int nb = (int)std::min<long long>(((long long)std::max(n, 1) + q - 1) / q * q, (long long)p2);
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lk(mu);
	auto key = std::make_tuple(dev, which, nb);
	auto it = cache.find(key);
	if (it != cache.end()) return it->second;
	const size_t bytes = which == 0 ? std::max(scan_temp_bytes(nb), depth_sort_temp_bytes(nb)) : tile_sort_temp_bytes(nb);
	cache[key] = bytes;
	return bytes;
}

struct GeomState {
	float* depths; uint8_t* clamped; float2* means2D; float* cov3D; float4* conic_opacity; float4* rgbd;
	float2* extent; uint32_t* tiles_touched; uint32_t* point_offsets; uint32_t* iota; uint32_t* order; uint32_t* depth_sorted;
	char* temp; size_t temp_bytes;
	static GeomState carve(char*& chunk, size_t P)
	{
		GeomState g;
		obtain(chunk, g.depths, P);
		obtain(chunk, g.clamped, P);
		obtain(chunk, g.means2D, P);
		obtain(chunk, g.cov3D, P * 6);
		obtain(chunk, g.conic_opacity, P);
		obtain(chunk, g.rgbd, P);
		obtain(chunk, g.extent, P);
		obtain(chunk, g.tiles_touched, P);
		obtain(chunk, g.point_offsets, P);  // inclusive sum of tiles_touched in DEPTH order
		obtain(chunk, g.iota, P);
		obtain(chunk, g.order, P);          // Gaussian ids sorted by depth bits (stable)
		obtain(chunk, g.depth_sorted, P);
		g.temp_bytes = cached_temp_bytes(0, (int)P);
		obtain(chunk, g.temp, g.temp_bytes);
		return g;
	}
};
struct ImageState {
	float* final_T; uint32_t* n_contrib; uint2* ranges; int* status; uint32_t* tile_order;
	static ImageState carve(char*& chunk, size_t N, size_t T)
	{
		ImageState s;
		obtain(chunk, s.final_T, N);
		obtain(chunk, s.n_contrib, N);
		obtain(chunk, s.ranges, T);
		obtain(chunk, s.status, 4);  // {instance count, overflow flag} of a forward that did not read the count back
		obtain(chunk, s.tile_order, T);
		return s;
	}
};
struct BinState {
	uint32_t* point_list; uint32_t* point_list_unsorted; uint32_t* tile_keys; uint32_t* tile_keys_unsorted; InstRec* recs;
	char* sort_temp; size_t sort_bytes;
	static BinState carve(char*& chunk, size_t R)
	{
		BinState b;
		const size_t Rp = R + 8;
		obtain(chunk, b.point_list, Rp);           // Gaussian ids in (tile, depth) order == the reference's point_list
		obtain(chunk, b.point_list_unsorted, Rp);  // ... in depth order, before the stable per-tile pass
		obtain(chunk, b.tile_keys, Rp);            // tile id of every sorted instance (high word of the reference's keys)
		obtain(chunk, b.tile_keys_unsorted, Rp);
		obtain(chunk, b.recs, Rp);
		b.sort_bytes = cached_temp_bytes(1, (int)R);
		obtain(chunk, b.sort_temp, b.sort_bytes);
		return b;
	}
};
template <typename F>
static size_t required(F carve)
{
	char* p = nullptr;
	carve(p);
	return reinterpret_cast<size_t>(p) + 128;
}
static size_t num_tiles(int W, int H) { return (size_t)ceil_div(W, TILE_X) * ceil_div(H, TILE_Y); }

// next-highest bit of the MSB (number of bits needed for tile ids); same result as the reference's
// getHigherMsb (rasterizer_impl.cu:35-50) for n >= 1
static int higher_msb(uint32_t n)
{
	int bits = 0;
	while (bits < 32 && (n >> bits)) bits++;
	return bits;
}

}  // namespace mgs

using namespace mgs;

extern "C" {

int mgs_abi_version(void) { return 200; }
const char* mgs_last_error(void) { return g_err.c_str(); }

size_t mgs_geometry_state_bytes(int P) { return required([&](char*& p) { GeomState::carve(p, (size_t)P); }); }
size_t mgs_image_state_bytes(int width, int height)
{
	return required([&](char*& p) { ImageState::carve(p, (size_t)width * height, num_tiles(width, height)); });
}
size_t mgs_binning_state_bytes(int R) { return required([&](char*& p) { BinState::carve(p, (size_t)R); }); }
size_t mgs_backward_scratch_bytes(int P) { return (size_t)P * GB_STRIDE * sizeof(float) + 128; }

// ---- forward, phase 1: per-Gaussian projection, depth order, instance offsets; the instance count R is copied to
// *host_num_rendered asynchronously on `st` (the caller synchronises before reading it) ----
static int forward_phase1(
	mgs_alloc_fn geometry_alloc, void* geometry_user, mgs_alloc_fn image_alloc, void* image_user,
	int P, int D, int M, int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
	float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
	const float* cam_pos, float tan_fovx, float tan_fovy, int* radii, int* host_num_rendered, int debug, cudaStream_t st,
	char** geometry_state, char** image_state)
{
	if (!means3D || !opacities || !viewmatrix || !projmatrix || !radii)
		return fail(MGS_ERR_INVALID_ARG, "means3D/opacities/viewmatrix/projmatrix/radii are required");
	if (!colors_precomp && !shs) return fail(MGS_ERR_INVALID_ARG, "provide SHs or precomputed colours");
	if (!colors_precomp && !cam_pos) return fail(MGS_ERR_INVALID_ARG, "cam_pos is required with SHs");
	if (!cov3D_precomp && (!scales || !rotations)) return fail(MGS_ERR_INVALID_ARG, "provide scales+rotations or a precomputed 3D covariance");
	const float focal_y = height / (2.0f * tan_fovy);  // rasterizer_impl.cu:225-226
	const float focal_x = width / (2.0f * tan_fovx);
	const int gx = ceil_div(width, TILE_X), gy = ceil_div(height, TILE_Y);

	char* gchunk = geometry_alloc(geometry_user, mgs_geometry_state_bytes(P));
	char* ichunk = image_alloc(image_user, mgs_image_state_bytes(width, height));
	if (!gchunk || !ichunk) return fail(MGS_ERR_ALLOC, "state allocation failed");
	*geometry_state = gchunk;
	*image_state = ichunk;
	GeomState geom = GeomState::carve(gchunk, (size_t)P);

	ProjectFwdArgs pa{};
	pa.P = P; pa.D = D; pa.M = M;
	pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier; pa.rotations = rotations;
	pa.opacities = opacities; pa.shs = shs; pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
	pa.viewmatrix = viewmatrix; pa.projmatrix = projmatrix; pa.cam_pos = cam_pos;
	pa.W = width; pa.H = height; pa.tan_fovx = tan_fovx; pa.tan_fovy = tan_fovy; pa.focal_x = focal_x; pa.focal_y = focal_y;
	pa.grid_x = gx; pa.grid_y = gy;
	pa.radii = radii; pa.means2D = geom.means2D; pa.depths = geom.depths; pa.cov3D = geom.cov3D; pa.rgbd = geom.rgbd;
	pa.conic_opacity = geom.conic_opacity; pa.extent = geom.extent; pa.clamped = geom.clamped; pa.tiles_touched = geom.tiles_touched;
	pa.iota = geom.iota;
	{ StageTimer t_(ST_PROJECT_FWD, st); launch_project_fwd(pa, st); }
	MGS_STAGE("project_fwd");

	// Gaussians in depth order (stable: ascending id among equal depth bits; culled ones carry +inf and emit nothing)
	{
		StageTimer t_(ST_DEPTH_SORT, st);
		launch_depth_sort(geom.temp, geom.temp_bytes, reinterpret_cast<const uint32_t*>(geom.depths), geom.depth_sorted, geom.iota,
			geom.order, P, st);
	}
	MGS_STAGE("depth_sort");
	{ StageTimer t_(ST_SCAN, st); launch_scan_sorted(geom.temp, geom.temp_bytes, geom.order, geom.tiles_touched, geom.point_offsets, P, st); }
	MGS_STAGE("scan");
	MGS_CUDA(cudaMemcpyAsync(host_num_rendered, geom.point_offsets + P - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
	return 0;
}

// ---- forward, phase 2: instance emission, per-tile order, records, blend ----
// stages: bit 0 = binning (allocates the binning state through the callback), bit 1 = blend (uses `binning_state` when
// given, else the one just allocated)
static int forward_phase2(
	mgs_alloc_fn binning_alloc, void* binning_user, char* binning_state, char* geometry_state, char* image_state,
	int P, int F, int width, int height, const float* background, const float* feature_precomp, const int* radii,
	int num_rendered, float* out_color, float* out_feature, float* out_depth, int stages, int debug, cudaStream_t st)
{
	if (num_rendered < 0) return fail(MGS_ERR_UNSUPPORTED, "more than 2^31-1 tile instances");
	const int gx = ceil_div(width, TILE_X), gy = ceil_div(height, TILE_Y);
	const size_t N = (size_t)width * height, T = (size_t)gx * gy;
	GeomState geom = GeomState::carve(geometry_state, (size_t)P);
	ImageState img = ImageState::carve(image_state, N, T);

	char* bchunk = binning_state;
	if (stages & 1) {
		if (!binning_alloc) return fail(MGS_ERR_INVALID_ARG, "binning allocator is required");
		bchunk = binning_alloc(binning_user, mgs_binning_state_bytes(num_rendered));
	}
	if (!bchunk) return fail(MGS_ERR_ALLOC, "binning state missing");
	BinState bin = BinState::carve(bchunk, (size_t)num_rendered);

	if (stages & 1) {
	{
		StageTimer t_(ST_EMIT, st);
		launch_emit_tiles(P, geom.order, geom.means2D, geom.point_offsets, radii, gx, gy, bin.tile_keys_unsorted, bin.point_list_unsorted,
			(uint32_t)num_rendered, nullptr, st);
	}
	MGS_STAGE("emit_tiles");
	if (num_rendered > 0) {
		// stable pass(es) over the tile-id bits only: instances already arrive in depth order
		const int bit = T > 1 ? higher_msb((uint32_t)(T - 1)) : 1;
		StageTimer t_(ST_SORT, st);
		launch_tile_sort(bin.sort_temp, bin.sort_bytes, bin.tile_keys_unsorted, bin.tile_keys, bin.point_list_unsorted, bin.point_list,
			num_rendered, bit, st);
		MGS_STAGE("tile_sort");
	}
	{
		StageTimer t_(ST_RANGES_PACK, st);
		launch_ranges_and_pack(num_rendered, nullptr, num_rendered, (int)T, gx, bin.tile_keys, bin.point_list, geom.means2D, geom.conic_opacity,
			geom.extent, img.ranges, bin.recs, img.tile_order, st);
	}
	MGS_STAGE("ranges_pack");
	}
	if (!(stages & 2)) return num_rendered;

	BlendArgs ba{};
	ba.W = width; ba.H = height; ba.grid_x = gx; ba.grid_y = gy; ba.F = F;
	ba.ranges = img.ranges; ba.tile_order = use_tile_order() ? img.tile_order : nullptr; ba.point_list = bin.point_list; ba.recs = bin.recs;
	ba.rgbd = geom.rgbd; ba.feature = F > 0 ? feature_precomp : nullptr; ba.bg = background;
	ba.want_depth = out_depth != nullptr;
	ba.final_T = img.final_T; ba.n_contrib = img.n_contrib;
	ba.out_color = out_color; ba.out_feature = out_feature; ba.out_depth = out_depth;
	{ StageTimer t_(ST_BLEND_FWD, st); MGS_CTA_LOG_ARGS(ba); launch_blend_fwd(ba, st); }
	MGS_STAGE("blend_fwd");
	return num_rendered;
}

static int forward_check(int P, int width, int height, int& F, const float* background, const float* feature_precomp,
	float* out_color, float* out_feature)
{
	if (P < 0 || width <= 0 || height <= 0) return fail(MGS_ERR_INVALID_ARG, "bad P/width/height");
	if (!out_color || !background) return fail(MGS_ERR_INVALID_ARG, "out_color/background are required");
	if (F < 0 || F > MGS_MAX_FEATURE_CHANNELS || !blend_supported(F))
		return fail(MGS_ERR_UNSUPPORTED, "feature channel count must be in [0, 32]");
	if (F > 0 && (!feature_precomp || !out_feature)) F = 0;  // include_feature == false
	return 0;
}

int mgs_forward(
	mgs_alloc_fn geometry_alloc, void* geometry_user,
	mgs_alloc_fn binning_alloc, void* binning_user,
	mgs_alloc_fn image_alloc, void* image_user,
	int P, int D, int M, int F,
	const float* background, int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp, const float* feature_precomp,
	const float* opacities, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	float tan_fovx, float tan_fovy, int prefiltered,
	float* out_color, float* out_feature, float* out_depth, int* radii, int debug, void* stream)
{
	(void)prefiltered;  // the reference only uses it to trap on inconsistent input (auxiliary.h:156-160)
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	if (!geometry_alloc || !binning_alloc || !image_alloc) return fail(MGS_ERR_INVALID_ARG, "allocator callbacks are required");
	int rc = forward_check(P, width, height, F, background, feature_precomp, out_color, out_feature);
	if (rc < 0) return rc;
	const size_t N = (size_t)width * height;
	if (P == 0) {
		// reference: outputs keep their zero fill when there are no Gaussians (rasterize_points.cu:70-92)
		MGS_CUDA(cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), st));
		if (F > 0) MGS_CUDA(cudaMemsetAsync(out_feature, 0, (size_t)F * N * sizeof(float), st));
		if (out_depth) MGS_CUDA(cudaMemsetAsync(out_depth, 0, N * sizeof(float), st));
		return 0;
	}
	char *gstate = nullptr, *istate = nullptr;
	int num_rendered = 0;
	rc = forward_phase1(geometry_alloc, geometry_user, image_alloc, image_user, P, D, M, width, height, means3D, shs, colors_precomp,
		opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, radii,
		&num_rendered, debug, st, &gstate, &istate);
	if (rc < 0) return rc;
	MGS_CUDA(cudaStreamSynchronize(st));  // the one host sync of the forward, as in the reference (rasterizer_impl.cu:284)
	return forward_phase2(binning_alloc, binning_user, nullptr, gstate, istate, P, F, width, height, background, feature_precomp, radii,
		num_rendered, out_color, out_feature, out_depth, 3, debug, st);
}

int mgs_forward_begin(
	mgs_alloc_fn geometry_alloc, void* geometry_user, mgs_alloc_fn image_alloc, void* image_user,
	int P, int D, int M, int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
	float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
	const float* cam_pos, float tan_fovx, float tan_fovy, int* radii, int* host_num_rendered, int debug, void* stream)
{
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	if (P <= 0 || width <= 0 || height <= 0) return fail(MGS_ERR_INVALID_ARG, "bad P/width/height (P must be > 0)");
	if (!geometry_alloc || !image_alloc || !host_num_rendered) return fail(MGS_ERR_INVALID_ARG, "allocators and host_num_rendered are required");
	char *gstate = nullptr, *istate = nullptr;
	return forward_phase1(geometry_alloc, geometry_user, image_alloc, image_user, P, D, M, width, height, means3D, shs, colors_precomp,
		opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, radii,
		host_num_rendered, debug, st, &gstate, &istate);
}

int mgs_forward_finish(
	mgs_alloc_fn binning_alloc, void* binning_user, char* binning_state, char* geometry_state, char* image_state,
	int P, int F, int width, int height, const float* background, const float* feature_precomp, const int* radii,
	int num_rendered, float* out_color, float* out_feature, float* out_depth, int stages, int debug, void* stream)
{
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	if (!geometry_state || !image_state || !radii || !(stages & 3)) return fail(MGS_ERR_INVALID_ARG, "required argument is missing");
	int rc = forward_check(P, width, height, F, background, feature_precomp, out_color, out_feature);
	if (rc < 0) return rc;
	if (P == 0) return fail(MGS_ERR_INVALID_ARG, "P must be > 0");
	return forward_phase2(binning_alloc, binning_user, binning_state, geometry_state, image_state, P, F, width, height, background,
		feature_precomp, radii, num_rendered, out_color, out_feature, out_depth, stages, debug, st);
}

int mgs_backward(
	int P, int D, int M, int F, int R,
	const float* background, int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp, const float* feature_precomp,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
	const int* radii, char* geometry_state, char* binning_state, char* image_state,
	const float* dL_dpix, const float* dL_dpix_F, const float* dL_dpix_depth,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dfeature,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
	char* blend_scratch, int accumulate, int debug, void* stream)
{
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	if (P < 0 || width <= 0 || height <= 0 || R < 0) return fail(MGS_ERR_INVALID_ARG, "bad P/R/width/height");
	if (P == 0) return 0;
	if (F < 0 || F > MGS_MAX_FEATURE_CHANNELS) return fail(MGS_ERR_UNSUPPORTED, "feature channel count must be in [0, 32]");
	if (F > 0 && (!feature_precomp || !dL_dpix_F || !dL_dfeature)) F = 0;
	if (!geometry_state || !binning_state || !image_state || !blend_scratch)
		return fail(MGS_ERR_INVALID_ARG, "state buffers and scratch are required");
	if (!dL_dpix || !dL_dmean2D || !dL_dopacity || !dL_dmean3D || !means3D || !radii || !background)
		return fail(MGS_ERR_INVALID_ARG, "required pointer is NULL");
	const float focal_y = height / (2.0f * tan_fovy);
	const float focal_x = width / (2.0f * tan_fovx);
	const int gx = ceil_div(width, TILE_X), gy = ceil_div(height, TILE_Y);
	const size_t N = (size_t)width * height, T = (size_t)gx * gy;

	GeomState geom = GeomState::carve(geometry_state, (size_t)P);
	BinState bin = BinState::carve(binning_state, (size_t)R);
	ImageState img = ImageState::carve(image_state, N, T);

	float* gb = nullptr;
	obtain(blend_scratch, gb, (size_t)P * GB_STRIDE);
	MGS_CUDA(cudaMemsetAsync(gb, 0, (size_t)P * GB_STRIDE * sizeof(float), st));
	if (F > 0 && !accumulate) MGS_CUDA(cudaMemsetAsync(dL_dfeature, 0, (size_t)P * F * sizeof(float), st));

	BlendArgs ba{};
	ba.W = width; ba.H = height; ba.grid_x = gx; ba.grid_y = gy; ba.F = F;
	ba.ranges = img.ranges; ba.tile_order = use_tile_order() ? img.tile_order : nullptr; ba.point_list = bin.point_list; ba.recs = bin.recs;
	ba.rgbd = geom.rgbd; ba.feature = F > 0 ? feature_precomp : nullptr; ba.bg = background;
	ba.want_depth = dL_dpix_depth != nullptr;
	ba.final_T = img.final_T; ba.n_contrib = img.n_contrib;
	ba.dL_dcolor = dL_dpix; ba.dL_dfeature = F > 0 ? dL_dpix_F : nullptr; ba.dL_ddepth = dL_dpix_depth;
	ba.gb = gb; ba.dL_dfeat = F > 0 ? dL_dfeature : nullptr;
	if (R > 0) {
		{ StageTimer t_(ST_BLEND_BWD, st); MGS_CTA_LOG_ARGS(ba); launch_blend_bwd(ba, st); }
		MGS_STAGE("blend_bwd");
	}

	ProjectBwdViewsArgs pb{};
	pb.P = P; pb.D = D; pb.M = M; pb.V = 1;
	pb.means3D = means3D; pb.shs = (shs && dL_dsh) ? shs : nullptr;
	const bool have_sr = scales && rotations && dL_dscale && dL_drot;
	pb.scales = have_sr ? scales : nullptr; pb.rotations = rotations; pb.scale_modifier = scale_modifier;
	pb.cov3D_precomp = cov3D_precomp ? cov3D_precomp : geom.cov3D;
	if (!have_sr && !pb.cov3D_precomp) return fail(MGS_ERR_INVALID_ARG, "provide scales+rotations or a precomputed 3D covariance");
	pb.accumulate = accumulate; pb.shared_mean2D = 1;
	pb.dL_dmean3D = dL_dmean3D; pb.dL_dopacity = dL_dopacity; pb.dL_dcolor = dL_dcolor; pb.dL_dcov3D = dL_dcov3D;
	pb.dL_dsh = pb.shs ? dL_dsh : nullptr; pb.dL_dscale = have_sr ? dL_dscale : nullptr; pb.dL_drot = have_sr ? dL_drot : nullptr;
	pb.dL_dconic = accumulate ? nullptr : dL_dconic;
	ProjectBwdView& pv = pb.view[0];
	pv.radii = radii; pv.gb = gb; pv.clamped = geom.clamped; pv.viewmatrix = viewmatrix; pv.projmatrix = projmatrix; pv.cam_pos = campos;
	pv.dL_dmean2D = dL_dmean2D; pv.tan_fovx = tan_fovx; pv.tan_fovy = tan_fovy; pv.focal_x = focal_x; pv.focal_y = focal_y;
	{ StageTimer t_(ST_PROJECT_BWD, st); launch_project_bwd_views(pb, st); }
	MGS_STAGE("project_bwd");
	return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// Multi-view step entry points: V views of ONE Gaussian cloud, no host synchronisation anywhere.
// ---------------------------------------------------------------------------------------------------------------------
namespace mgs {
// fork/join events between the caller's stream and the per-view streams (created once, reused; never destroyed)
struct EventPool {
	std::mutex mu;
	std::vector<cudaEvent_t> free_;
	cudaEvent_t get()
	{
		{
			std::lock_guard<std::mutex> lk(mu);
			if (!free_.empty()) { cudaEvent_t e = free_.back(); free_.pop_back(); return e; }
		}
		cudaEvent_t e = nullptr;
		cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
		return e;
	}
	void put(cudaEvent_t e) { std::lock_guard<std::mutex> lk(mu); free_.push_back(e); }
};
static EventPool g_events;

// every view stream waits for what the join stream has enqueued so far
static int fork_streams(int V, const mgs_view* views, cudaStream_t join)
{
	bool need = false;
	for (int v = 0; v < V; v++) need |= reinterpret_cast<cudaStream_t>(views[v].stream) != join;
	if (!need) return 0;
	cudaEvent_t e = g_events.get();
	MGS_CUDA(cudaEventRecord(e, join));
	for (int v = 0; v < V; v++) {
		cudaStream_t st = reinterpret_cast<cudaStream_t>(views[v].stream);
		if (st != join) MGS_CUDA(cudaStreamWaitEvent(st, e, 0));
	}
	g_events.put(e);  // recorded and waited on: safe to re-record later (waits captured the earlier record)
	return 0;
}
// every view stream waits for what every OTHER view stream has enqueued so far
static int cross_wait_streams(int V, const mgs_view* views)
{
	std::vector<cudaStream_t> uniq;
	for (int v = 0; v < V; v++) {
		cudaStream_t st = reinterpret_cast<cudaStream_t>(views[v].stream);
		bool seen = false;
		for (cudaStream_t u : uniq) seen |= u == st;
		if (!seen) uniq.push_back(st);
	}
	if (uniq.size() < 2) return 0;
	std::vector<cudaEvent_t> ev(uniq.size());
	for (size_t i = 0; i < uniq.size(); i++) {
		ev[i] = g_events.get();
		MGS_CUDA(cudaEventRecord(ev[i], uniq[i]));
	}
	for (size_t i = 0; i < uniq.size(); i++)
		for (size_t j = 0; j < uniq.size(); j++)
			if (i != j) MGS_CUDA(cudaStreamWaitEvent(uniq[i], ev[j], 0));
	for (cudaEvent_t e : ev) g_events.put(e);
	return 0;
}
static bool barrier_before_blend()
{
	static const bool on = [] { const char* e = getenv("MGS_BLEND_BARRIER"); return e && e[0] == '1'; }();
	return on;
}
// the join stream waits for every view stream
static int join_streams(int V, const mgs_view* views, cudaStream_t join)
{
	for (int v = 0; v < V; v++) {
		cudaStream_t st = reinterpret_cast<cudaStream_t>(views[v].stream);
		if (st == join) continue;
		bool seen = false;
		for (int u = 0; u < v; u++) seen |= views[u].stream == views[v].stream;
		if (seen) continue;
		cudaEvent_t e = g_events.get();
		MGS_CUDA(cudaEventRecord(e, st));
		MGS_CUDA(cudaStreamWaitEvent(join, e, 0));
		g_events.put(e);
	}
	return 0;
}
}  // namespace mgs

int mgs_forward_views(
	int V, const mgs_view* views,
	int P, int D, int M, int F,
	const float* means3D, const float* shs, const float* colors_precomp, const float* feature_precomp,
	const float* opacities, const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	int prefiltered, int debug, void* join_stream)
{
	(void)prefiltered;
	if (V <= 0 || !views) return fail(MGS_ERR_INVALID_ARG, "need at least one view");
	if (P <= 0) return fail(MGS_ERR_INVALID_ARG, "P must be > 0");
	if (!means3D || !opacities) return fail(MGS_ERR_INVALID_ARG, "means3D/opacities are required");
	if (!colors_precomp && !shs) return fail(MGS_ERR_INVALID_ARG, "provide SHs or precomputed colours");
	if (!cov3D_precomp && (!scales || !rotations)) return fail(MGS_ERR_INVALID_ARG, "provide scales+rotations or a precomputed 3D covariance");
	if (F < 0 || F > MGS_MAX_FEATURE_CHANNELS || !blend_supported(F)) return fail(MGS_ERR_UNSUPPORTED, "feature channel count must be in [0, 32]");
	for (int v = 0; v < V; v++) {
		const mgs_view& w = views[v];
		if (!w.viewmatrix || !w.projmatrix || !w.background || !w.geometry_state || !w.binning_state || !w.image_state || !w.out_color || !w.radii)
			return fail(MGS_ERR_INVALID_ARG, "view: matrices, background, the three state buffers, out_color and radii are required");
		if (!colors_precomp && !w.cam_pos) return fail(MGS_ERR_INVALID_ARG, "view: cam_pos is required with SHs");
		if (w.width <= 0 || w.height <= 0 || w.binning_capacity < 0) return fail(MGS_ERR_INVALID_ARG, "view: bad size or capacity");
	}
	cudaStream_t join = reinterpret_cast<cudaStream_t>(join_stream);
	int rc = fork_streams(V, views, join);
	if (rc < 0) return rc;
	// stage by stage across the views: the short per-Gaussian and binning kernels of every view are enqueued before any
	// view's long blend kernel, so no view's chain queues behind another view's blend
	for (int phase = 0; phase < 3; phase++) {
		// MGS_BLEND_BARRIER=1: all blends start together, after the binning chain of EVERY view.  A blend launch is one
		// wave of single-warp CTAs that takes every shared-memory slot of the GPU, and the short binning kernels of the
		// other views crawl behind it (CTA timeline, tools/cta_timeline.py: blends 0 and 1 each run alone, 790 us for the
		// four forward blends of the bench step; started together 590 us) -- but four binning chains side by side take
		// 0.38 ms instead of 0.22 ms for the first alone, and the step time comes out the same (2.21 ms either way,
		// DESIGN.md section 7), so the default keeps the dependency-free schedule.
		if (phase == 2 && barrier_before_blend()) {
			rc = cross_wait_streams(V, views);
			if (rc < 0) return rc;
		}
		for (int v = 0; v < V; v++) {
			const mgs_view& w = views[v];
			cudaStream_t st = reinterpret_cast<cudaStream_t>(w.stream);
			const int gx = ceil_div(w.width, TILE_X), gy = ceil_div(w.height, TILE_Y);
			const size_t N = (size_t)w.width * w.height, T = (size_t)gx * gy;
			char* gchunk = w.geometry_state; char* ichunk = w.image_state; char* bchunk = w.binning_state;
			GeomState geom = GeomState::carve(gchunk, (size_t)P);
			ImageState img = ImageState::carve(ichunk, N, T);
			BinState bin = BinState::carve(bchunk, (size_t)w.binning_capacity);
			const int Fv = (F > 0 && feature_precomp && w.out_feature) ? F : 0;
			if (phase == 0) {
				ProjectFwdArgs pa{};
				pa.P = P; pa.D = D; pa.M = M;
				pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier; pa.rotations = rotations;
				pa.opacities = opacities; pa.shs = shs; pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
				pa.viewmatrix = w.viewmatrix; pa.projmatrix = w.projmatrix; pa.cam_pos = w.cam_pos;
				pa.W = w.width; pa.H = w.height; pa.tan_fovx = w.tan_fovx; pa.tan_fovy = w.tan_fovy;
				pa.focal_x = w.width / (2.0f * w.tan_fovx); pa.focal_y = w.height / (2.0f * w.tan_fovy);
				pa.grid_x = gx; pa.grid_y = gy;
				pa.radii = w.radii; pa.means2D = geom.means2D; pa.depths = geom.depths; pa.cov3D = geom.cov3D; pa.rgbd = geom.rgbd;
				pa.conic_opacity = geom.conic_opacity; pa.extent = geom.extent; pa.clamped = geom.clamped; pa.tiles_touched = geom.tiles_touched;
				pa.iota = geom.iota;
				{ StageTimer t_(ST_PROJECT_FWD, st); launch_project_fwd(pa, st); }
				MGS_STAGE("project_fwd");
				{
					StageTimer t_(ST_DEPTH_SORT, st);
					launch_depth_sort(geom.temp, geom.temp_bytes, reinterpret_cast<const uint32_t*>(geom.depths), geom.depth_sorted, geom.iota,
						geom.order, P, st);
				}
				MGS_STAGE("depth_sort");
				{ StageTimer t_(ST_SCAN, st); launch_scan_sorted(geom.temp, geom.temp_bytes, geom.order, geom.tiles_touched, geom.point_offsets, P, st); }
				MGS_STAGE("scan");
			} else if (phase == 1) {
				const uint32_t cap = (uint32_t)w.binning_capacity;
				{
					StageTimer t_(ST_EMIT, st);
					launch_emit_tiles(P, geom.order, geom.means2D, geom.point_offsets, w.radii, gx, gy, bin.tile_keys_unsorted, bin.point_list_unsorted,
						cap, img.status, st);
					// the slots behind the R real instances sort to the very end: last tile id, ids never read
					launch_fill_tail(geom.point_offsets, P, cap, (uint32_t)(T - 1), bin.tile_keys_unsorted, bin.point_list_unsorted, st);
				}
				MGS_STAGE("emit_tiles");
				if (cap > 0) {
					const int bit = T > 1 ? higher_msb((uint32_t)(T - 1)) : 1;
					StageTimer t_(ST_SORT, st);
					launch_tile_sort(bin.sort_temp, bin.sort_bytes, bin.tile_keys_unsorted, bin.tile_keys, bin.point_list_unsorted, bin.point_list,
						(int)cap, bit, st);
					MGS_STAGE("tile_sort");
				}
				{
					StageTimer t_(ST_RANGES_PACK, st);
					launch_ranges_and_pack(-1, geom.point_offsets + P - 1, (int)cap, (int)T, gx, bin.tile_keys, bin.point_list, geom.means2D,
						geom.conic_opacity, geom.extent, img.ranges, bin.recs, img.tile_order, st);
				}
				MGS_STAGE("ranges_pack");
				if (w.status) MGS_CUDA(cudaMemcpyAsync(w.status, img.status, 2 * sizeof(int), cudaMemcpyDefault, st));
			} else {
				BlendArgs ba{};
				ba.W = w.width; ba.H = w.height; ba.grid_x = gx; ba.grid_y = gy; ba.F = Fv;
				ba.ranges = img.ranges; ba.tile_order = use_tile_order() ? img.tile_order : nullptr; ba.point_list = bin.point_list; ba.recs = bin.recs;
				ba.rgbd = geom.rgbd; ba.feature = Fv > 0 ? feature_precomp : nullptr; ba.bg = w.background;
				ba.want_depth = w.out_depth != nullptr;
				ba.final_T = img.final_T; ba.n_contrib = img.n_contrib;
				ba.out_color = w.out_color; ba.out_feature = w.out_feature; ba.out_depth = w.out_depth;
				if (w.target_color) {
					if (!w.cot_color || !w.loss_acc || (w.target_feature && Fv > 0 && !w.cot_feature))
						return fail(MGS_ERR_INVALID_ARG, "view: loss heads need cot_color, loss_acc (and cot_feature with target_feature)");
					ba.tgt_color = w.target_color; ba.tgt_feature = Fv > 0 ? w.target_feature : nullptr;
					ba.cot_color = w.cot_color; ba.cot_feature = w.cot_feature; ba.loss_acc = w.loss_acc;
					MGS_CUDA(cudaMemsetAsync(w.loss_acc, 0, 2 * sizeof(float), st));
				}
				{ StageTimer t_(ST_BLEND_FWD, st); MGS_CTA_LOG_ARGS(ba); launch_blend_fwd(ba, st); }
				MGS_STAGE("blend_fwd");
			}
		}
	}
	return join_streams(V, views, join);
}

int mgs_backward_views(
	int V, const mgs_view* views,
	int P, int D, int M, int F,
	const float* means3D, const float* shs, const float* colors_precomp, const float* feature_precomp,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	float* dL_dmean3D, float* dL_dopacity, float* dL_dcolor, float* dL_dfeature, float* dL_dcov3D, float* dL_dsh,
	float* dL_dscale, float* dL_drot, int shared_mean2D, int accumulate, int stages, int debug, void* join_stream)
{
	(void)colors_precomp;
	if (!(stages & 3)) return fail(MGS_ERR_INVALID_ARG, "stages must name the blend stage (1), the per-Gaussian stage (2) or both (3)");
	if (V <= 0 || !views) return fail(MGS_ERR_INVALID_ARG, "need at least one view");
	if (P <= 0) return fail(MGS_ERR_INVALID_ARG, "P must be > 0");
	if (!means3D || !dL_dmean3D || !dL_dopacity) return fail(MGS_ERR_INVALID_ARG, "means3D, dL_dmean3D and dL_dopacity are required");
	if (F < 0 || F > MGS_MAX_FEATURE_CHANNELS) return fail(MGS_ERR_UNSUPPORTED, "feature channel count must be in [0, 32]");
	if (F > 0 && (!feature_precomp || !dL_dfeature)) F = 0;
	for (int v = 0; v < V; v++) {
		const mgs_view& w = views[v];
		if (!w.viewmatrix || !w.projmatrix || !w.background || !w.geometry_state || !w.binning_state || !w.image_state || !w.radii ||
			!w.dL_dpix || !w.blend_scratch)
			return fail(MGS_ERR_INVALID_ARG, "view: matrices, background, state buffers, radii, dL_dpix and blend_scratch are required");
		if (F > 0 && !w.dL_dpix_F) return fail(MGS_ERR_INVALID_ARG, "view: dL_dpix_F is required with features");
	}
	cudaStream_t join = reinterpret_cast<cudaStream_t>(join_stream);
	int rc = 0;
	if (stages & 1) {
	{
		cudaStream_t st = join;  // for MGS_CUDA's messages
		(void)st;
		if (F > 0 && !accumulate) MGS_CUDA(cudaMemsetAsync(dL_dfeature, 0, (size_t)P * F * sizeof(float), join));
	}
	rc = fork_streams(V, views, join);
	if (rc < 0) return rc;
	// blend stage of every view on its own stream: per-view blend-stage records, feature gradients summed with 128-bit
	// reductions straight into dL_dfeature
	for (int v = 0; v < V; v++) {
		const mgs_view& w = views[v];
		cudaStream_t st = reinterpret_cast<cudaStream_t>(w.stream);
		const int gx = ceil_div(w.width, TILE_X), gy = ceil_div(w.height, TILE_Y);
		const size_t N = (size_t)w.width * w.height, T = (size_t)gx * gy;
		char* gchunk = w.geometry_state; char* ichunk = w.image_state; char* bchunk = w.binning_state; char* schunk = w.blend_scratch;
		GeomState geom = GeomState::carve(gchunk, (size_t)P);
		ImageState img = ImageState::carve(ichunk, N, T);
		BinState bin = BinState::carve(bchunk, (size_t)w.binning_capacity);
		float* gb = nullptr;
		obtain(schunk, gb, (size_t)P * GB_STRIDE);
		MGS_CUDA(cudaMemsetAsync(gb, 0, (size_t)P * GB_STRIDE * sizeof(float), st));
		BlendArgs ba{};
		ba.W = w.width; ba.H = w.height; ba.grid_x = gx; ba.grid_y = gy; ba.F = F;
		ba.ranges = img.ranges; ba.tile_order = use_tile_order() ? img.tile_order : nullptr; ba.point_list = bin.point_list; ba.recs = bin.recs;
		ba.rgbd = geom.rgbd; ba.feature = F > 0 ? feature_precomp : nullptr; ba.bg = w.background;
		ba.want_depth = w.dL_dpix_depth != nullptr;
		ba.final_T = img.final_T; ba.n_contrib = img.n_contrib;
		ba.dL_dcolor = w.dL_dpix; ba.dL_dfeature = F > 0 ? w.dL_dpix_F : nullptr; ba.dL_ddepth = w.dL_dpix_depth;
		ba.gb = gb; ba.dL_dfeat = F > 0 ? dL_dfeature : nullptr;
		ba.cot_scale = w.cot_scale;
		if (w.binning_capacity > 0) {
			{ StageTimer t_(ST_BLEND_BWD, st); MGS_CTA_LOG_ARGS(ba); launch_blend_bwd(ba, st); }
			MGS_STAGE("blend_bwd");
		}
	}
	rc = join_streams(V, views, join);
	if (rc < 0) return rc;
	}  // blend stage: dL_dfeature is final here -- a caller may start exchanging it while the per-Gaussian stage runs
	if (!(stages & 2)) return 0;
	// one launch per MAX_BWD_VIEWS views: the per-Gaussian chain rule, summed over the views in registers
	const bool have_sr = scales && rotations && dL_dscale && dL_drot;
	for (int v0 = 0; v0 < V; v0 += MAX_BWD_VIEWS) {
		cudaStream_t st = join;
		ProjectBwdViewsArgs pb{};
		pb.P = P; pb.D = D; pb.M = M; pb.V = std::min(V - v0, (int)MAX_BWD_VIEWS);
		pb.means3D = means3D; pb.shs = (shs && dL_dsh) ? shs : nullptr;
		pb.scales = have_sr ? scales : nullptr; pb.rotations = rotations; pb.scale_modifier = scale_modifier;
		pb.cov3D_precomp = cov3D_precomp;
		if (!have_sr && !cov3D_precomp) {
			// gradients w.r.t. scale/rotation not requested: the forward's own covariance (any view's copy) serves
			char* gchunk = views[0].geometry_state;
			pb.cov3D_precomp = GeomState::carve(gchunk, (size_t)P).cov3D;
		}
		pb.accumulate = (accumulate || v0 > 0) ? 1 : 0;
		pb.shared_mean2D = shared_mean2D;
		pb.dL_dmean3D = dL_dmean3D; pb.dL_dopacity = dL_dopacity; pb.dL_dcolor = dL_dcolor; pb.dL_dcov3D = dL_dcov3D;
		pb.dL_dsh = pb.shs ? dL_dsh : nullptr; pb.dL_dscale = have_sr ? dL_dscale : nullptr; pb.dL_drot = have_sr ? dL_drot : nullptr;
		pb.dL_dconic = nullptr;
		for (int u = 0; u < pb.V; u++) {
			const mgs_view& w = views[v0 + u];
			char* gchunk = w.geometry_state; char* schunk = w.blend_scratch;
			GeomState geom = GeomState::carve(gchunk, (size_t)P);
			float* gb = nullptr;
			obtain(schunk, gb, (size_t)P * GB_STRIDE);
			ProjectBwdView& pv = pb.view[u];
			pv.radii = w.radii; pv.gb = gb; pv.clamped = geom.clamped; pv.viewmatrix = w.viewmatrix; pv.projmatrix = w.projmatrix;
			pv.cam_pos = w.cam_pos; pv.dL_dmean2D = shared_mean2D ? views[0].dL_dmean2D : w.dL_dmean2D;
			pv.tan_fovx = w.tan_fovx; pv.tan_fovy = w.tan_fovy;
			pv.focal_x = w.width / (2.0f * w.tan_fovx); pv.focal_y = w.height / (2.0f * w.tan_fovy);
		}
		{ StageTimer t_(ST_PROJECT_BWD, st); launch_project_bwd_views(pb, st); }
		MGS_STAGE("project_bwd");
	}
	return 0;
}

int mgs_loss_heads(int V, int F, int N, const float* color, const float* feature, const float* target_color,
	const float* target_feature, float* cot_color, float* cot_feature, float* loss_acc, void* stream)
{
	if (V <= 0 || N <= 0 || F < 0) return fail(MGS_ERR_INVALID_ARG, "bad V/N/F");
	if (!color || !target_color || !cot_color || !loss_acc) return fail(MGS_ERR_INVALID_ARG, "color, target_color, cot_color and loss_acc are required");
	const bool emb = F > 0 && feature && target_feature;
	if (emb && !cot_feature) return fail(MGS_ERR_INVALID_ARG, "cot_feature is required with features");
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	MGS_CUDA(cudaMemsetAsync(loss_acc, 0, (size_t)V * 2 * sizeof(float), st));
	launch_loss_heads(V, emb ? F : 0, N, color, emb ? feature : nullptr, target_color, emb ? target_feature : nullptr, cot_color, cot_feature,
		loss_acc, st);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(MGS_ERR_CUDA, std::string("loss_heads: ") + cudaGetErrorString(e));
	return 0;
}

int mgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present, void* stream)
{
	(void)projmatrix;
	if (P < 0) return fail(MGS_ERR_INVALID_ARG, "bad P");
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !present) return fail(MGS_ERR_INVALID_ARG, "required pointer is NULL");
	launch_mark_visible(P, means3D, viewmatrix, present, reinterpret_cast<cudaStream_t>(stream));
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(MGS_ERR_CUDA, cudaGetErrorString(e));
	return 0;
}

static int activate_common(ActivateArgs& a, int P, int F, const float* means, const float* d_means, const float* rot, const float* d_rot,
	const float* scales, const float* d_scales, const float* opac, const float* feature,
	int scale_mode, float scale_max, int opacity_mode, int rot_normalize, int feature_normalize)
{
	if (P < 0 || F < 0) return fail(MGS_ERR_INVALID_ARG, "bad P or F");
	if (scale_mode < 0 || scale_mode > 1 || opacity_mode < 0 || opacity_mode > 1) return fail(MGS_ERR_INVALID_ARG, "unknown activation mode");
	memset(&a, 0, sizeof(a));
	a.P = P; a.F = F;
	a.means = means; a.d_means = d_means; a.rot = rot; a.d_rot = d_rot; a.scales = scales; a.d_scales = d_scales;
	a.opac = opac; a.feature = F > 0 ? feature : nullptr;
	a.scale_mode = scale_mode; a.scale_max = scale_max; a.opacity_mode = opacity_mode;
	a.rot_normalize = rot_normalize != 0; a.feature_normalize = feature_normalize != 0;
	return 0;
}

int mgs_activate(int P, int F,
	const float* means, const float* d_means, const float* rot, const float* d_rot,
	const float* scales, const float* d_scales, const float* opac, const float* feature,
	int scale_mode, float scale_max, int opacity_mode, int rot_normalize, int feature_normalize,
	float* out_means, float* out_rot, float* out_scales, float* out_opac, float* out_feature,
	void* stream)
{
	ActivateArgs a;
	int rc = activate_common(a, P, F, means, d_means, rot, d_rot, scales, d_scales, opac, feature,
		scale_mode, scale_max, opacity_mode, rot_normalize, feature_normalize);
	if (rc) return rc;
	if (P == 0) return 0;
	if ((out_means && !means) || (out_rot && !rot) || (out_scales && !scales) || (out_opac && !opac) || (out_feature && F > 0 && !feature))
		return fail(MGS_ERR_INVALID_ARG, "an output was requested for an input that is NULL");
	a.o_means = out_means; a.o_rot = out_rot; a.o_scales = out_scales; a.o_opac = out_opac; a.o_feature = F > 0 ? out_feature : nullptr;
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	{ StageTimer t_(ST_ACTIVATE_FWD, st); launch_activate_fwd(a, st); }
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(MGS_ERR_CUDA, std::string("activate: ") + cudaGetErrorString(e));
	return 0;
}

int mgs_activate_backward(int P, int F,
	const float* means, const float* d_means, const float* rot, const float* d_rot,
	const float* scales, const float* d_scales, const float* opac, const float* feature,
	int scale_mode, float scale_max, int opacity_mode, int rot_normalize, int feature_normalize,
	const float* g_means, const float* g_rot, const float* g_scales, const float* g_opac, const float* g_feature,
	float* dL_dmeans, float* dL_dd_means, float* dL_drot, float* dL_dd_rot,
	float* dL_dscales, float* dL_dd_scales, float* dL_dopac, float* dL_dfeature,
	void* stream)
{
	ActivateArgs a;
	int rc = activate_common(a, P, F, means, d_means, rot, d_rot, scales, d_scales, opac, feature,
		scale_mode, scale_max, opacity_mode, rot_normalize, feature_normalize);
	if (rc) return rc;
	if (P == 0) return 0;
	if ((g_rot && !rot) || (g_scales && scale_mode == 1 && !scales) || (g_opac && opacity_mode == 1 && !opac) || (g_feature && F > 0 && !feature))
		return fail(MGS_ERR_INVALID_ARG, "a gradient was given for an input that is NULL");
	a.g_means = g_means; a.g_rot = g_rot; a.g_scales = g_scales; a.g_opac = g_opac; a.g_feature = F > 0 ? g_feature : nullptr;
	a.dL_means = dL_dmeans; a.dL_dmeans = dL_dd_means; a.dL_rot = dL_drot; a.dL_drot = dL_dd_rot;
	a.dL_scales = dL_dscales; a.dL_dscales = dL_dd_scales; a.dL_opac = dL_dopac; a.dL_feature = F > 0 ? dL_dfeature : nullptr;
	cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
	{ StageTimer t_(ST_ACTIVATE_BWD, st); launch_activate_bwd(a, st); }
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(MGS_ERR_CUDA, std::string("activate_backward: ") + cudaGetErrorString(e));
	return 0;
}

int mgs_profile_enable(int on)
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	g_prof_on = on != 0;
	return 0;
}

// Sums the recorded stage durations (ms) and launch counts since the last call, then clears them.
// total_ms / counts: arrays of mgs_profile_num_stages() entries.  Synchronises on the recorded events.
int mgs_profile_num_stages(void) { return ST_COUNT; }
const char* mgs_profile_stage_name(int i)
{
	static const char* names[ST_COUNT] = { "project_fwd", "depth_sort", "scan", "emit_tiles", "tile_sort", "ranges_pack", "blend_fwd", "blend_bwd", "project_bwd", "activate_fwd", "activate_bwd" };
	return (i >= 0 && i < ST_COUNT) ? names[i] : "";
}
int mgs_profile_read(float* total_ms, int* counts)
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	for (int i = 0; i < ST_COUNT; i++) { total_ms[i] = 0.f; counts[i] = 0; }
	for (auto& e : g_prof_evts) {
		float ms = 0.f;
		cudaEventSynchronize(e.b);
		if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { total_ms[e.stage] += ms; counts[e.stage]++; }
		cudaEventDestroy(e.a); cudaEventDestroy(e.b);
	}
	g_prof_evts.clear();
	return 0;
}

int mgs_state_array(const char* which_state, const char* name, char* state, int a0, int a1, void** out_ptr)
{
	if (!which_state || !name || !state || !out_ptr) return fail(MGS_ERR_INVALID_ARG, "NULL argument");
	const std::string w(which_state), n(name);
	void* p = nullptr;
	if (w == "geometry") {
		GeomState g = GeomState::carve(state, (size_t)a0);
		if (n == "depths") p = g.depths; else if (n == "means2D") p = g.means2D; else if (n == "cov3D") p = g.cov3D;
		else if (n == "conic_opacity") p = g.conic_opacity; else if (n == "rgbd") p = g.rgbd;
		else if (n == "tiles_touched") p = g.tiles_touched; else if (n == "point_offsets") p = g.point_offsets;
		else if (n == "depth_order") p = g.order;
		else if (n == "clamped") p = g.clamped; else if (n == "extent") p = g.extent;
	} else if (w == "binning") {
		BinState b = BinState::carve(state, (size_t)a0);
		if (n == "point_list") p = b.point_list; else if (n == "tile_ids") p = b.tile_keys;
		else if (n == "point_list_unsorted") p = b.point_list_unsorted; else if (n == "tile_ids_unsorted") p = b.tile_keys_unsorted;
		else if (n == "records") p = b.recs;
	} else if (w == "image") {
		ImageState s = ImageState::carve(state, (size_t)a0 * a1, num_tiles(a0, a1));
		if (n == "final_T") p = s.final_T; else if (n == "n_contrib") p = s.n_contrib; else if (n == "ranges") p = s.ranges;
		else if (n == "tile_order") p = s.tile_order;
	}
	if (!p) return fail(MGS_ERR_INVALID_ARG, "unknown state array " + w + "/" + n);
	*out_ptr = p;
	return 0;
}

}  // extern "C"
