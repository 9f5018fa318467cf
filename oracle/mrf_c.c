/*
 * CPU oracle (plain C) for the frame-based MRF stereo matcher.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product
 * library; used by tests/ as a fast checker at sizes where the numpy oracle
 * is too slow, after being validated itself against the golden fixtures the
 * real reference produced (tests/test_oracle.py).
 *
 * Restates stereovis/framed/algorithms/mrf.py (reference) with a
 * disparity-INNERMOST layout: every plane is H x W x D float32, element
 * (y, x, l) at ((y * W) + x) * D + l.  The reference stores (D, H, W); the
 * arithmetic per element is identical, so results agree bit for bit after a
 * transpose.
 *
 * Build (see oracle/Makefile):
 *   gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math mrf_c.c -o _build/libmrf_oracle.so
 * -ffp-contract=off matters: the adaptive prior blend is (1 - d) * d with two
 * roundings (mrf.py:60,66) and must not become an fma.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define AT(y, x) (((size_t)(y) * (size_t)W + (size_t)(x)) * (size_t)D)

/* np.max semantics: NaN propagates. */
static inline float max_np(float acc, float v)
{
    if (acc != acc) return acc;
    if (v != v) return v;
    return v > acc ? v : acc;
}

/*
 * Data term, mrf.py:50-71.
 *   mode 0: no prior            Dt = |L - R(x-l)|
 *   mode 1: const, f32 factor   Dt = one_minus_tau * diff      where prior == l
 *   mode 2: adaptive            Dt = (1 - diff) * diff         where prior == l
 *   mode 3: const, f64 factor   Dt = (float)(omt64 * (double)diff)  (numpy float64 scalar factor)
 * prior_label: int32 per pixel, the label the prior equals exactly, or -1.
 * Columns x < l stay 0 (never written in the reference).
 */
void mrf_c_cost_volume(const float *L, const float *R, const int32_t *prior_label, int mode,
                       float one_minus_tau, double one_minus_tau64, int H, int W, int D, float *Dt)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float *out = Dt + AT(y, x);
            const float left = L[(size_t)y * W + x];
            const int hit = (prior_label && mode != 0) ? prior_label[(size_t)y * W + x] : -1;
            for (int l = 0; l < D; ++l) {
                if (x < l) { out[l] = 0.0f; continue; }
                float diff = fabsf(left - R[(size_t)y * W + (x - l)]);
                if (hit == l) {
                    if (mode == 1) {
                        diff = one_minus_tau * diff;
                    } else if (mode == 2) {
                        float keep = 1.0f - diff;
                        diff = keep * diff;
                    } else if (mode == 3) {
                        diff = (float)(one_minus_tau64 * (double)diff);
                    }
                }
                out[l] = diff;
            }
        }
    }
}

/* Normalise one pixel of a plane in place: divide by max over l, 0 -> 1 (mrf.py:92-94). */
static inline void normalise_pixel(float *v, int D)
{
    float m = v[0];
    for (int l = 1; l < D; ++l) m = max_np(m, v[l]);
    if (m == 0.0f) m = 1.0f;
    for (int l = 0; l < D; ++l) v[l] = v[l] / m;
}

/*
 * One sweep, mrf.py:82-94.  a, b, c are the three message planes other than
 * `target`, in dict order; the sum is ((a + b) + c) + Dt.  (dy, dx) is where
 * the message goes: target[y + dy, x + dx] = U[y, x].  The target plane is not
 * read by its own sweep, so the shifted write can go straight into it; the
 * border line that receives nothing keeps its old content and is normalised
 * like the rest (it is all zeros in practice).
 */
static void sweep(float *target, const float *a, const float *b, const float *c, const float *Dt,
                  int dy, int dx, int H, int W, int D)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        const int sy = y - dy;
        for (int x = 0; x < W; ++x) {
            const int sx = x - dx;
            float *out = target + AT(y, x);
            if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
                const size_t s = AT(sy, sx);
                for (int l = 0; l < D; ++l) {
                    float u = a[s + l] + b[s + l];
                    u = u + c[s + l];
                    u = u + Dt[s + l];
                    out[l] = u;
                }
            }
            normalise_pixel(out, D);
        }
    }
}

/* n_iter iterations of south, west, north, east (mrf.py:73-94, 130-133). */
void mrf_c_iterate(float *S, float *Wp, float *N, float *E, const float *Dt,
                   int H, int W, int D, int n_iter)
{
    for (int it = 0; it < n_iter; ++it) {
        sweep(S, Wp, N, E, Dt, +1, 0, H, W, D);    /* S[y+1] = U[y]   mrf.py:84 */
        sweep(Wp, S, N, E, Dt, 0, -1, H, W, D);    /* W[x-1] = U[x]   mrf.py:86 */
        sweep(N, S, Wp, E, Dt, -1, 0, H, W, D);    /* N[y-1] = U[y]   mrf.py:88 */
        sweep(E, S, Wp, N, Dt, 0, +1, H, W, D);    /* E[x+1] = U[x]   mrf.py:90 */
    }
}

/*
 * Energy = (((S + W) + N) + E) + Dt and first-minimum argmin (mrf.py:103-104).
 * np.argmin treats NaN as the minimum (first NaN wins).  energy may be NULL.
 */
void mrf_c_readout(const float *S, const float *Wp, const float *N, const float *E, const float *Dt,
                   int H, int W, int D, int64_t *labels, float *energy)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const size_t s = AT(y, x);
            float best = 0.0f;
            int arg = 0;
            int best_nan = 0;
            for (int l = 0; l < D; ++l) {
                float e = S[s + l] + Wp[s + l];
                e = e + N[s + l];
                e = e + E[s + l];
                e = e + Dt[s + l];
                if (energy) energy[s + l] = e;
                if (l == 0) { best = e; best_nan = (e != e); continue; }
                if (best_nan) continue;
                if (e != e) { best_nan = 1; arg = l; best = e; }
                else if (e < best) { best = e; arg = l; }
            }
            labels[(size_t)y * W + x] = arg;
        }
    }
}
