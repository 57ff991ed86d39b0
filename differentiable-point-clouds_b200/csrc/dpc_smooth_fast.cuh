// Shape-specialised smoothing / projection kernels (placeholder until the FFMA2 kernels land).
#pragma once
#include "dpc_common.cuh"

static inline bool dpc_conv_xy_fast_supported(int, int, int, int, int) { return false; }
static inline int dpc_conv_xy_fast_launch(const float*, float*, const float*, const float*, int, int, int, int, int,
                                          uint32_t*, const uint32_t*, void*) { return DPC_ERR_ARG; }
static inline bool dpc_conv_z_fast_supported(int, int, int, int) { return false; }
static inline int dpc_conv_z_fwd_fast_launch(const float*, const float*, int, const float*, int, float, float, float, int,
                                             int, int, int, float*, uint32_t*, float*, float*, float*, void*) { return DPC_ERR_ARG; }
static inline int dpc_conv_z_bwd_fast_launch(const float*, const uint32_t*, const float*, const float*, int, int, float,
                                             float, float, int, int, int, int, const float*, const float*, const float*,
                                             const float*, float*, float*, void*) { return DPC_ERR_ARG; }
