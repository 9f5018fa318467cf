// Prior rasterisation on the device: event list -> per-frame prior images.
//
// Restates generate_frames_from_spikes / split_frames_by_time of stereovis/utils/frames_io.py:184-248 (reference):
// frame f collects the events whose timestamp lies in its window -- [tick_f - interval, tick_f] when the caller
// gives pivots (hybrid_stereo.py:169-176 passes the frame timestamps), [T_(f-1), T_f) between the wrap-arounds
// of t % interval otherwise -- in chronological order, and
//     frames[f, rows_f, cols_f] = vals_f            # numpy fancy assignment: the LAST event wins
// i.e. a pixel shows the value of the latest event that hit it inside the window.  The reference gets the
// chronological order from np.argsort (not stable: the order of events with EQUAL timestamps is an accident of
// the sort); here no sort is needed at all:
//     pass 1  atomicMax over an order-preserving 64-bit key of t   -> the latest timestamp per (frame, pixel)
//     pass 2  among the events carrying that timestamp, atomicMax of the event index (a stable sort's order)
//     pass 3  gather the winner's value: float64 frame for the host and / or int32 label frame that stays on the
//             device as the matcher's prior input (exact integer -> that integer, anything else matches no level)
// Windows may overlap (interval > pivot spacing) and need not be sorted.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hsm {

__device__ __forceinline__ unsigned long long time_key(double t)
{
    const long long b = __double_as_longlong(t);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

struct RasterArgs {
    const int *x, *y;
    const double *t;
    size_t n_events;
    const double *lo, *hi;      // per-frame window bounds (device)
    int n_frames, hi_inclusive;
    int monotone;               // lo and hi are both non-decreasing: binary search for the first candidate frame
    int rows, cols;
};

// Calls visit(frame, pixel index inside the frame stack) for every window the event falls into.
// numpy index semantics: a negative coordinate wraps once, anything else out of range is an IndexError.
template <typename Visit>
__device__ __forceinline__ void for_each_hit(const RasterArgs &a, size_t e, int *error, Visit visit)
{
    const double t = a.t[e];
    int f = 0;
    if (a.monotone) {                      // first frame whose upper bound admits t
        int lo_i = 0, hi_i = a.n_frames;
        while (lo_i < hi_i) {
            const int mid = (lo_i + hi_i) >> 1;
            const bool below = a.hi_inclusive ? (a.hi[mid] < t) : (a.hi[mid] <= t);
            if (below) lo_i = mid + 1; else hi_i = mid;
        }
        f = lo_i;
    }
    for (; f < a.n_frames; ++f) {
        const bool in_lo = t >= a.lo[f];
        if (!in_lo) {
            if (a.monotone) break;
            continue;
        }
        if (!(a.hi_inclusive ? (t <= a.hi[f]) : (t < a.hi[f]))) continue;
        int xi = a.x[e], yi = a.y[e];
        if (xi < 0) xi += a.cols;
        if (yi < 0) yi += a.rows;
        if (xi < 0 || xi >= a.cols || yi < 0 || yi >= a.rows) {
            atomicMax(error, 1);
            return;
        }
        visit(((size_t)f * a.rows + yi) * a.cols + xi);
    }
}

static __global__ void k_raster_latest(RasterArgs a, unsigned long long *__restrict__ latest, int *__restrict__ error)
{
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < a.n_events; e += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long key = time_key(a.t[e]);
        for_each_hit(a, e, error, [&](size_t at) { atomicMax(latest + at, key); });
    }
}

static __global__ void k_raster_winner(RasterArgs a, const unsigned long long *__restrict__ latest,
                                       unsigned *__restrict__ winner, int *__restrict__ error)
{
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < a.n_events; e += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long key = time_key(a.t[e]);
        for_each_hit(a, e, error, [&](size_t at) {
            if (latest[at] == key) atomicMax(winner + at, (unsigned)e + 1u);
        });
    }
}

__device__ __forceinline__ int label_of(double v)
{
    return (v >= -2147483648.0 && v <= 2147483647.0 && v == floor(v)) ? (int)v : -1;
}

static __global__ void k_raster_emit(const unsigned *__restrict__ winner, const double *__restrict__ z, size_t n_pixels,
                                     double fill, double *__restrict__ frames, int *__restrict__ labels)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned w = winner[i];
        const double v = w ? z[w - 1u] : fill;
        if (frames) frames[i] = v;
        if (labels) labels[i] = label_of(v);
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

// ---------------------------------------------------------------------------
// Frame pre-processing (row N3), down-scaling: load_frames' `downsample` (frames_io.py:99-101) is
//     skimage.transform.rescale(img, 1.0 / scale_down_factor, preserve_range=True, mode='constant')
// with scikit-image 0.13.0 (requirements.txt:33; un-vendored).  That release's rescale -> resize -> warp path
// for a 2-D image is: output shape = round(scale * shape); an affine map from output to input pixel centres,
//     c = col_scale * (x + 0.5) - 0.5,   r = row_scale * (y + 0.5) - 0.5,   *_scale = input size / output size
// bilinear interpolation (order 1) between floor and ceil neighbours with out-of-image neighbours = cval = 0
// (mode 'constant'), NO anti-aliasing filter (added in 0.14), then clipping to the input's [min, max] where
// output pixels that are exactly cval keep cval when cval lies outside that range.  Arithmetic in float64 in
// the order of skimage's `bilinear_interpolation`.  One difference is deliberate: skimage ESTIMATES the affine
// matrix from three corner correspondences by least squares, so its coefficients carry rounding noise of a
// few ulp; here the matrix is the exact one.  PARITY UNPINNED: scikit-image is not installable in the build
// container, so no fixture of its output exists.
// One CTA per frame: min / max of the input frame, then the map.
// ---------------------------------------------------------------------------
namespace hsm {

__device__ __forceinline__ double pixel_or_cval(const double *img, int rows, int cols, long r, long c)
{
    return (r < 0 || r >= rows || c < 0 || c >= cols) ? 0.0 : img[r * cols + c];
}

static __global__ void __launch_bounds__(1024) k_rescale_bilinear(const double *__restrict__ in, double *__restrict__ out,
                                                                  int rows, int cols, int out_rows, int out_cols)
{
    __shared__ double s_min[32], s_max[32];
    const double *src = in + (size_t)blockIdx.x * rows * cols;
    double *dst = out + (size_t)blockIdx.x * out_rows * out_cols;
    const size_t n_in = (size_t)rows * cols, n_out = (size_t)out_rows * out_cols;
    double lo = src[0], hi = src[0];
    for (size_t i = threadIdx.x; i < n_in; i += blockDim.x) {
        lo = dmin_nan(lo, src[i]);
        hi = dmax_nan(hi, src[i]);
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
    const bool preserve_cval = !(lo <= 0.0 && 0.0 <= hi);
    const double row_scale = (double)rows / (double)out_rows, col_scale = (double)cols / (double)out_cols;
    const double off_r = __dsub_rn(__dmul_rn(row_scale, 0.5), 0.5), off_c = __dsub_rn(__dmul_rn(col_scale, 0.5), 0.5);
    for (size_t i = threadIdx.x; i < n_out; i += blockDim.x) {
        const int y = (int)(i / out_cols), x = (int)(i % out_cols);
        const double c = __dadd_rn(__dmul_rn(col_scale, (double)x), off_c);
        const double r = __dadd_rn(__dmul_rn(row_scale, (double)y), off_r);
        const long minr = (long)floor(r), minc = (long)floor(c), maxr = (long)ceil(r), maxc = (long)ceil(c);
        const double dr = __dsub_rn(r, (double)minr), dc = __dsub_rn(c, (double)minc);
        const double top = __dadd_rn(__dmul_rn(__dsub_rn(1.0, dc), pixel_or_cval(src, rows, cols, minr, minc)),
                                     __dmul_rn(dc, pixel_or_cval(src, rows, cols, minr, maxc)));
        const double bottom = __dadd_rn(__dmul_rn(__dsub_rn(1.0, dc), pixel_or_cval(src, rows, cols, maxr, minc)),
                                        __dmul_rn(dc, pixel_or_cval(src, rows, cols, maxr, maxc)));
        double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, dr), top), __dmul_rn(dr, bottom));
        if (!(preserve_cval && v == 0.0)) v = v < lo ? lo : (v > hi ? hi : v);          // np.clip (NaN passes through)
        dst[i] = v;
    }
}

// ---------------------------------------------------------------------------
// Prior adjustment (row N4): OpticalFlowPixelCorrection.adjust, stereovis/spiking/optical_flow_correction.py:34-88,
// REPRODUCED AS WRITTEN including the defect its own FIXME (:69) names: the fitted velocity field is never
// used -- compute_shift is called with the pixel's (col, row) COORDINATES, so every prior pixel with a value
// >= 0 moves by one step of the 8-direction compass chosen by the angle of its position vector:
//     k = int(floor(round(8 * arctan2(row, col) / (2 pi)))) % 8,   (dcol, drow) = time_arrow * shifts[k]
//     (no move when |(col, row)| <= 1), target = (row - drow, col + dcol) if inside the frame, else dropped.
// The reference scatters in np.argwhere (row-major) order, so on a collision the LAST source in row-major order
// wins; here every target pixel looks at its (at most four) possible sources and takes the one with the
// largest row-major index -- the same result without write races.  Output float32 like the reference's
// np.empty_like(event_frames, dtype=np.float32).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void compass_shift(int col, int row, int time_arrow, int &dcol, int &drow)
{
    dcol = 0; drow = 0;
    if (!(sqrt((double)col * col + (double)row * row) > 1.0)) return;
    const double turns = 8.0 * atan2((double)row, (double)col) / (2.0 * 3.141592653589793);
    int k = (int)floor(rint(turns)) % 8;                 // np.round = round half to even = rint
    if (k < 0) k += 8;
    const int sx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, sy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    dcol = time_arrow * sx[k];
    drow = time_arrow * sy[k];
}

static __global__ void k_adjust_priors(const double *__restrict__ prior, float *__restrict__ out, int n_frames, int rows,
                                       int cols, double nondata, int time_arrow)
{
    const size_t n = (size_t)n_frames * rows * cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols), r = (int)((i / cols) % rows);
        const double *frame = prior + (i / ((size_t)rows * cols)) * (size_t)rows * cols;
        double value = nondata;
        long best = -1;
        // sources whose one-step move could land here: the pixel itself (no move) and its 8 neighbours
        for (int sr = r - 1; sr <= r + 1; ++sr) {
            for (int sc = c - 1; sc <= c + 1; ++sc) {
                if (sr < 0 || sr >= rows || sc < 0 || sc >= cols) continue;
                const double v = frame[(size_t)sr * cols + sc];
                if (!(v >= 0.0)) continue;
                int dcol, drow;
                compass_shift(sc, sr, time_arrow, dcol, drow);
                if (sr - drow != r || sc + dcol != c) continue;
                const long at = (long)sr * cols + sc;
                if (at > best) { best = at; value = v; }
            }
        }
        out[i] = (float)value;
    }
}

}  // namespace hsm
