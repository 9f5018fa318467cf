// Prior rasterisation: event list -> per-frame prior images.
//
// Restates the scatter of stereovis/utils/frames_io.py:239-244 (reference):
//     frames = ones((n_frames, rows, cols)) * non_pixel_value
//     frames[i, rows_i, cols_i] = vals_i            # numpy fancy assignment: the LAST event wins
// Events arrive already ordered and grouped per frame by the host (frames_io.py:184-204 is index
// bookkeeping); `order` is the event's position in that order, so "last wins" becomes an
// atomicMax on the position, followed by a gather of the winning value.
#pragma once

#include <cuda_runtime.h>

namespace hsm {

__global__ void k_raster_winner(const int *__restrict__ ex, const int *__restrict__ ey,
                                const int *__restrict__ eframe, size_t n_events, int rows, int cols,
                                int *__restrict__ winner)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_events; i += (size_t)gridDim.x * blockDim.x) {
        const size_t at = ((size_t)eframe[i] * rows + ey[i]) * cols + ex[i];
        atomicMax(winner + at, (int)i);
    }
}

__global__ void k_raster_gather(const int *__restrict__ winner, const double *__restrict__ ez, size_t n_pixels,
                                double fill, double *__restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * blockDim.x) {
        const int w = winner[i];
        out[i] = w >= 0 ? ez[w] : fill;
    }
}

}  // namespace hsm

// ---------------------------------------------------------------------------
// Frame pre-processing: per-frame min-max contrast normalisation,
// stereovis/utils/frames_io.py:103-109 (reference):
//     if max - min != 0:  out = 1. / (max - min) * (img - min)      (float64, two roundings + divide)
//     else:               out = ones_like(img)
// One CTA per frame: block-wide min/max in float64 (np.max / np.min propagate NaN), then the map.
// ---------------------------------------------------------------------------
namespace hsm {

__device__ __forceinline__ double dmin_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : (b < a ? b : a)); }
__device__ __forceinline__ double dmax_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : (b > a ? b : a)); }

__global__ void __launch_bounds__(1024) k_normalise_contrast(const double *__restrict__ in, double *__restrict__ out,
                                                             size_t frame_elems)
{
    __shared__ double s_min[32], s_max[32];
    const double *src = in + (size_t)blockIdx.x * frame_elems;
    double *dst = out + (size_t)blockIdx.x * frame_elems;
    double lo = src[0], hi = src[0];
    for (size_t i = threadIdx.x; i < frame_elems; i += blockDim.x) {
        const double v = src[i];
        lo = dmin_nan(lo, v);
        hi = dmax_nan(hi, v);
    }
    for (int s = 16; s > 0; s >>= 1) {
        lo = dmin_nan(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = dmax_nan(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    const int n_warps = (blockDim.x + 31) >> 5;
    lo = s_min[0]; hi = s_max[0];
    for (int w = 1; w < n_warps; ++w) { lo = dmin_nan(lo, s_min[w]); hi = dmax_nan(hi, s_max[w]); }
    const double span = __dsub_rn(hi, lo);
    if (span != 0.0) {                                     // NaN != 0 is true, as in numpy
        const double scale = __ddiv_rn(1.0, span);
        for (size_t i = threadIdx.x; i < frame_elems; i += blockDim.x)
            dst[i] = __dmul_rn(scale, __dsub_rn(src[i], lo));
    } else {
        for (size_t i = threadIdx.x; i < frame_elems; i += blockDim.x) dst[i] = 1.0;
    }
}

}  // namespace hsm
