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
#include <math_constants.h>
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

// ---------------------------------------------------------------------------
// Event-based visual flow (row N4): VelocityVectorField._fit_velocity_field, stereovis/spiking/algorithms/vvf.py:32-82
// (Benosman et al. 2014 as the reference implements it).  Events (t, x, y) of one polarity, sorted by time.  For
// every event: the candidates are the events with  t_e - dt < t <= t_e + dt  (the two np.searchsorted(side='right')
// of vvf.py:46-47) inside the (2 nx + 1) x (2 ny + 1) pixel neighbourhood; with fewer than `min_events` of them the
// velocity stays 0; otherwise a plane  t ~ a x + b y + c  is fitted by least squares and refined at most
// `max_iter` times: keep the candidates with |(t, x, y) . n| <= rejection_threshold for n = (c, a, b) / |(c, a, b)|
// (the reference's expression, vvf.py:66-69, as written), stop if fewer than `min_events` or all of them remain,
// refit on the kept ones, stop when the parameters move by less than `convergence_threshold`.  Result: (a, b).
// The reference solves each fit with scipy.linalg.lstsq (LAPACK gelsd, an SVD); here the 3 x 3 normal equations
// are accumulated in float64 with coordinates taken relative to the event (so the sums stay small and the
// solution matches the SVD's to ~1e-9 relative for the well-posed neighbourhoods) and solved by Cramer's rule;
// collinear neighbourhoods (an edge's events all in one pixel column) get the minimum-norm solution, as an
// SVD-based lstsq returns it, from an eigen-decomposition of the normal matrix (min_norm_solve3).
// One thread per event; every refit is another pass over the event's time window.
// ---------------------------------------------------------------------------
// Minimum-norm least-squares solution of the 3 x 3 normal equations M p = b for a (numerically) singular M, via
// a cyclic Jacobi eigen-decomposition: p = sum over eigenvalues above the threshold of (v . b / lambda) v -- what
// an SVD-based lstsq returns for a rank-deficient system (collinear neighbourhoods: e.g. an edge's events all in
// one pixel column).
__device__ __forceinline__ void min_norm_solve3(const double m_in[6], const double b[3], double p[3])
{
    // m_in = {m00, m01, m02, m11, m12, m22}
    double a[3][3] = {{m_in[0], m_in[1], m_in[2]}, {m_in[1], m_in[3], m_in[4]}, {m_in[2], m_in[4], m_in[5]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off == 0.0) break;
        for (int pi = 0; pi < 2; ++pi) {
            for (int qi = pi + 1; qi < 3; ++qi) {
                if (a[pi][qi] == 0.0) continue;
                const double theta = (a[qi][qi] - a[pi][pi]) / (2.0 * a[pi][qi]);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
                for (int k = 0; k < 3; ++k) {                 // A <- A J
                    const double akp = a[k][pi], akq = a[k][qi];
                    a[k][pi] = c * akp - sn * akq;
                    a[k][qi] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {                 // A <- J^T A
                    const double apk = a[pi][k], aqk = a[qi][k];
                    a[pi][k] = c * apk - sn * aqk;
                    a[qi][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {                 // V <- V J
                    const double vkp = v[k][pi], vkq = v[k][qi];
                    v[k][pi] = c * vkp - sn * vkq;
                    v[k][qi] = sn * vkp + c * vkq;
                }
            }
        }
    }
    const double lmax = fmax(fmax(fabs(a[0][0]), fabs(a[1][1])), fabs(a[2][2]));
    p[0] = p[1] = p[2] = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double lambda = a[i][i];
        if (!(lambda > 1e-13 * lmax)) continue;
        const double w = (v[0][i] * b[0] + v[1][i] * b[1] + v[2][i] * b[2]) / lambda;
        p[0] += w * v[0][i]; p[1] += w * v[1][i]; p[2] += w * v[2][i];
    }
}

struct FlowFitArgs {
    const double *t, *x, *y;
    size_t n;
    double dt, nx, ny, rejection, convergence;
    int max_iter, min_events;
};

static __global__ void k_vvf_fit(FlowFitArgs a, double *__restrict__ vel)
{
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < a.n; e += (size_t)gridDim.x * blockDim.x) {
        const double te = a.t[e], xe = a.x[e], ye = a.y[e];
        double va = 0.0, vb = 0.0;
        // window [lo, hi): first index with t > te - dt, first index with t > te + dt
        size_t lo, hi;
        {
            size_t l = 0, h = a.n;
            const double key = te - a.dt;
            while (l < h) { const size_t m = (l + h) >> 1; if (a.t[m] <= key) l = m + 1; else h = m; }
            lo = l;
            l = lo; h = a.n;
            const double key2 = te + a.dt;
            while (l < h) { const size_t m = (l + h) >> 1; if (a.t[m] <= key2) l = m + 1; else h = m; }
            hi = l;
        }
        // parameters of the current plane in absolute coordinates (a, b, c); `use_plane` = refit with rejection
        double pa = 0.0, pb = 0.0, pc = 0.0;
        bool have = false;
        long n_candidates = 0;
        for (int pass = 0; pass <= a.max_iter; ++pass) {
            double nrm = 0.0, na = 0.0, nb = 0.0, nc = 0.0;
            if (have) {
                nrm = sqrt(pc * pc + pa * pa + pb * pb);
                nc = pc / nrm; na = pa / nrm; nb = pb / nrm;
            }
            double sxx = 0, sxy = 0, syy = 0, sx = 0, sy = 0, sxt = 0, syt = 0, st = 0;
            long cnt = 0;
            for (size_t j = lo; j < hi; ++j) {
                const double xj = a.x[j], yj = a.y[j];
                if (!(fabs(xj - xe) <= a.nx && fabs(yj - ye) <= a.ny)) continue;
                const double tj = a.t[j];
                if (have && !(fabs(tj * nc + xj * na + yj * nb) <= a.rejection)) continue;
                const double dx = xj - xe, dy = yj - ye, dtj = tj - te;
                sxx += dx * dx; sxy += dx * dy; syy += dy * dy; sx += dx; sy += dy;
                sxt += dx * dtj; syt += dy * dtj; st += dtj;
                ++cnt;
            }
            if (!have) {
                n_candidates = cnt;
                if (cnt < a.min_events) break;                       // velocity stays 0 (vvf.py:53-55)
            } else if (cnt < a.min_events || cnt == n_candidates) {
                break;                                               // keep the parameters found so far
            }
            // normal equations in event-relative coordinates: [sxx sxy sx; sxy syy sy; sx sy n] p = [sxt syt st]
            const double n = (double)cnt;
            const double det = sxx * (syy * n - sy * sy) - sxy * (sxy * n - sy * sx) + sx * (sxy * sy - syy * sx);
            const double scale = (sxx + syy + n);
            double qa, qb, qc;
            if (fabs(det) > 1e-12 * scale * scale * scale) {
                qa = (sxt * (syy * n - sy * sy) - sxy * (syt * n - sy * st) + sx * (syt * sy - syy * st)) / det;
                qb = (sxx * (syt * n - st * sy) - sxt * (sxy * n - sy * sx) + sx * (sxy * st - syt * sx)) / det;
                const double qc_rel =
                    (sxx * (syy * st - sy * syt) - sxy * (sxy * st - syt * sx) + sxt * (sxy * sy - syy * sx)) / det;
                qc = qc_rel + te - qa * xe - qb * ye;                // back to absolute coordinates
            } else {
                // collinear neighbourhood: the minimum-norm solution in ABSOLUTE coordinates (it is not
                // translation invariant); absolute sums from the event-relative ones
                const double X = sx + n * xe, Y = sy + n * ye, T = st + n * te;
                const double m[6] = {sxx + 2.0 * xe * sx + n * xe * xe, sxy + xe * sy + ye * sx + n * xe * ye, X,
                                     syy + 2.0 * ye * sy + n * ye * ye, Y, n};
                const double rhs[3] = {sxt + xe * st + te * sx + n * xe * te, syt + ye * st + te * sy + n * ye * te, T};
                double p[3];
                min_norm_solve3(m, rhs, p);
                qa = p[0]; qb = p[1]; qc = p[2];
            }
            if (have) {
                const double da = pa - qa, db = pb - qb, dc = pc - qc;
                const double eps = sqrt(da * da + db * db + dc * dc);
                pa = qa; pb = qb; pc = qc;
                if (!(eps > a.convergence)) break;
            } else {
                pa = qa; pb = qb; pc = qc;
                have = true;
            }
        }
        if (have) { va = pa; vb = pb; }
        vel[2 * e] = va;
        vel[2 * e + 1] = vb;
    }
}

}  // namespace hsm
