// TEST SCAFFOLDING -- force-included (nvcc -include) in front of every reference .cu
// file by oracle/build_ref.py.  It supplies, without editing the reference tree:
//   * <cstdint>, which cuda_rasterizer/rasterizer_impl.h needs under gcc >= 13;
//   * the four macros of cuda_rasterizer/config.h:15-18, with the feature width taken
//     from -DMGS_REF_FEATURE_CHANNELS (stock value 3), config.h itself being skipped via
//     its own include guard.
// CUB is parsed FIRST because it uses NUM_CHANNELS as a template-parameter name; in the
// stock translation unit <cub/cub.cuh> likewise precedes config.h (rasterizer_impl.cu:20-28).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cub/cub.cuh>
#include <cub/device/device_radix_sort.cuh>
#endif
#ifndef MGS_REF_FEATURE_CHANNELS
#define MGS_REF_FEATURE_CHANNELS 3
#endif
#define CUDA_RASTERIZER_CONFIG_H_INCLUDED
#define NUM_CHANNELS 3
#define NUM_CHANNELS_language_feature MGS_REF_FEATURE_CHANNELS
#define BLOCK_X 16
#define BLOCK_Y 16
