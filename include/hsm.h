/*
 * hsm.h -- C ABI of the B200-native frame-based MRF stereo matcher.
 *
 * This is the drop-in boundary for ONE hot path of gdikov/hybrid-stereo-matching:
 * stereovis/framed/algorithms/mrf.py, class StereoMRF (reference mrf.py:8-167).
 * Every entry point names the reference interface it replaces.  Signatures use
 * plain pointers and sizes only (no torch / numpy types); the Python class
 * `StereoMRF` in hybrid-stereo-matching_b200/mrf.py binds them with ctypes and
 * keeps the reference's method surface.  INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns HSM_OK (0) or an HSM_ERR_* code; the message is
 *     available from hsm_last_error() (thread-local);
 *   - images are row-major (rows x cols) arrays, frames stacked outermost;
 *   - message/data planes live on the device as frames x rows x cols x Dp float32,
 *     disparity INNERMOST, Dp = n_levels rounded up to a multiple of 4 (the
 *     reference keeps (D, H, W); hsm_get_plane returns rows x cols x D and the
 *     Python layer transposes for `_message_field`);
 *   - plane indices follow the reference's dict order (mrf.py:15-19):
 *     0 south, 1 west, 2 north, 3 east, 4 data;
 *   - no CUDA call is made before the first compute call (hsm_create only
 *     records the geometry), because the reference constructs the matcher in a
 *     parent process and runs it in a fork()ed child (hybrid_stereo.py:107-111,
 *     online_processing.py:148).
 *   - one matcher = one CUDA stream; calls on one matcher are not thread-safe.
 */
#ifndef HSM_H_
#define HSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSM_OK 0
#define HSM_ERR_ARG 1    /* bad argument / shape contract violated           */
#define HSM_ERR_CUDA 2   /* CUDA runtime error (Python raises RuntimeError)   */
#define HSM_ERR_STATE 3  /* call sequence error (e.g. readout of a stale S)   */

/* element types of caller arrays */
#define HSM_F32 0
#define HSM_F64 1
#define HSM_I32 2
#define HSM_I64 3

/* prior blending, reference mrf.py:52-67 */
#define HSM_PRIOR_NONE 0      /* prior is ignored / absent                                    */
#define HSM_PRIOR_CONST 1     /* Dt = keep_f32 * diff where prior == l   (keep = (float)(1 - tau)) */
#define HSM_PRIOR_ADAPTIVE 2  /* Dt = (1 - diff) * diff where prior == l                       */
#define HSM_PRIOR_CONST_F64 3 /* Dt = (float)(keep_f64 * (double)diff): tau was a numpy float64 scalar */

/* hsm_set_option keys */
#define HSM_OPT_ALGORITHM 0   /* 0 = one kernel per direction (T80), 1 = fused iteration fed by LDG,
                                 2 = fused iteration fed by TMA (falls back to 1 if n_levels > 256),
                                 3 = two iterations per launch through a shared-memory row ring
                                     (falls back to 2 for an odd leftover / unsupported shapes),
                                 4 = like 2 with a TMA producer warp and warp-to-warp exchange (no CTA barrier),
                                 5 = like 2 with TMA stores (results staged in shared memory; DEFAULT; falls
                                     back to 2 when the staging does not fit in shared memory, e.g. n_levels 256),
                                 6 = persistent dataflow: all iterations of an hsm_iterate call in ONE launch of
                                     kernel 5's code; (iteration, band, strip) tasks are ordered by per-tile
                                     progress counters in global memory (falls back to 5).  With the default 5
                                     it is chosen automatically for a single frame of >= 8M elements         */
#define HSM_OPT_BAND_ROWS 1   /* rows per CTA band in the fused kernel; 0 = heuristic                */
#define HSM_OPT_DATA_TERM 2   /* 0 = read the data plane (T28), 1 = recompute it in-kernel (T24, default) */
#define HSM_OPT_VALID_LO 3    /* row-band mode: first/last local row that is inside the image ...    */
#define HSM_OPT_VALID_HI 4
#define HSM_OPT_OUT_LO 5      /* ... and first/last local row this matcher owns (computes)           */
#define HSM_OPT_OUT_HI 6
#define HSM_OPT_STAGES 7      /* shared-memory row stages of the TMA pipeline; 0 = automatic (2..3)          */
#define HSM_OPT_LANE_CONFIG 8 /* tuning: 0 = default lanes-per-pixel layout for n_levels and width, 1..16 = table entry
                                 (10..13: 5 / 6 lanes per pixel, 14..16: 7- / 9- / 10-warp CTAs; default kernel family only) */
#define HSM_OPT_FRAME_GROUPS 9 /* 2 = launch every iteration as two kernels (halves of the frames) on two streams so
                                  that each fills the other's tail; 1 = one kernel; 0 = automatic (2 whenever there are >= 2 frames) */

typedef struct hsm_matcher hsm_matcher;

typedef struct hsm_stats {
    uint64_t kernel_launches;      /* kernels of this library launched by this matcher        */
    uint64_t iteration_launches;   /* message-kernel launches (4 per iteration if algorithm 0) */
    uint64_t iterations;           /* BP iterations executed                                   */
    double iteration_ms;           /* device time of all hsm_iterate loops (CUDA events)       */
    uint64_t device_bytes;         /* device memory owned by the matcher                       */
} hsm_stats;

int hsm_version(void);
const char *hsm_last_error(void);
/* number of CUDA devices visible; fails with HSM_ERR_CUDA when there is no usable GPU */
int hsm_device_count(int *count);

/*
 * Replaces StereoMRF.__init__ (mrf.py:12-19): a matcher for `capacity` frames of
 * rows x cols pixels and n_levels disparities.  Requires n_levels <= cols + 1 (the
 * reference's slicing fails beyond that) and n_levels <= 512.  No CUDA call.
 */
int hsm_create(int rows, int cols, int n_levels, int capacity, int device, hsm_matcher **out);
int hsm_destroy(hsm_matcher *m);

/* Use the caller's CUDA stream (a cudaStream_t passed as void*) instead of the matcher's own
 * non-blocking stream; NULL goes back to the matcher's own stream.  (To run on the legacy
 * default stream pass cudaStreamLegacy, not NULL.) */
int hsm_set_stream(hsm_matcher *m, void *cuda_stream);
int hsm_set_option(hsm_matcher *m, int key, int value);
int hsm_get_option(hsm_matcher *m, int key, int *value);
/* Bytes of device memory the matcher needs (allocated on first use). */
int hsm_device_bytes(int rows, int cols, int n_levels, int capacity, size_t *bytes);

/*
 * Replaces StereoMRF._init_fields (mrf.py:21-71) for n_frames <= capacity frames:
 * casts the images to float32 (mrf.py:41-42), zeroes the four message planes if
 * reinit_messages (mrf.py:43-48) and defines the data term (+ prior blend).
 * left/right: n_frames x rows x cols of image_dtype (HSM_F32 / HSM_F64).
 * prior: NULL or n_frames x rows x cols of prior_dtype (any HSM_* type); an entry
 * matches level l iff it equals l exactly (mrf.py:57); -1 / NaN match nothing.
 * on_device: 0 = host arrays; 1 = all pointers are device pointers on the matcher's device;
 * 2 = images on the host, prior on the device (a prior rasterised by hsm_rasterise_events_device).
 * The inputs are snapshotted (the online prior is a live shared buffer,
 * online_processing.py:121).
 */
int hsm_init_fields(hsm_matcher *m, int n_frames, const void *left, const void *right,
                    int image_dtype, const void *prior, int prior_dtype, int prior_mode,
                    float keep_f32, double keep_f64, int reinit_messages, int on_device);

/*
 * Replaces the loop over StereoMRF._update_message_fields (mrf.py:73-94, 130-133).
 * finalize != 0 makes the last iteration also store the south plane (it is a
 * transient inside a fused iteration); readout needs a finalized state.
 */
int hsm_iterate(hsm_matcher *m, int n_iter, int finalize);

/*
 * Replaces StereoMRF._update_belief_field / get_map_belief (mrf.py:96-104, 136-144):
 * labels[f, y, x] = first argmin over l of (((S + W) + N) + E) + Dt, int64 like
 * np.argmin.  labels: n_frames x rows x cols int64, host (or device if on_device).
 * async != 0: do not wait for the copy (call hsm_sync before reading a host buffer;
 * the buffer should be pinned, see hsm_host_alloc).
 */
int hsm_readout(hsm_matcher *m, int64_t *labels, int on_device, int async);

/* Replaces StereoMRF.lbp (mrf.py:146-167): init (reinit) + n_iter iterations + readout.
 * async: 0 = returns when the labels are in `labels`; 1 = returns when the INPUT host buffers may be reused,
 * labels arrive by hsm_sync; 2 = nothing is waited for: every host buffer stays untouched until hsm_sync. */
int hsm_lbp(hsm_matcher *m, int n_frames, const void *left, const void *right, int image_dtype,
            const void *prior, int prior_dtype, int prior_mode, float keep_f32, double keep_f64,
            int n_iter, int64_t *labels, int on_device, int async);

int hsm_sync(hsm_matcher *m);

/*
 * Read-back of `StereoMRF._message_field[...]` (mrf.py:15-19) for one frame:
 * out is rows x cols x n_levels float32 on the host (disparity innermost).
 */
int hsm_get_plane(hsm_matcher *m, int frame, int which, float *out);
/* Energy field of mrf.py:103, same layout as hsm_get_plane. */
int hsm_get_energy(hsm_matcher *m, int frame, float *out);
/* Overwrite one plane of one frame from the host (tests, warm start from saved state). */
int hsm_set_plane(hsm_matcher *m, int frame, int which, const float *in);

/*
 * Row-band mode (one large frame split across GPUs).  The
 * matcher holds local rows [0, rows) of which [out_lo, out_hi] are owned and the
 * rest are ghost rows.  After each iteration the owned boundary rows of the three
 * planes that cross an iteration (west, north, east) are exchanged:
 *   pack:   top    <- row out_lo of (W, N, E), bottom <- row out_hi   (3 x cols x Dp floats each)
 *   unpack: row out_lo - 1 <- from_above, row out_hi + 1 <- from_below (NULL = skip)
 * All pointers are device pointers; copies are ordered on the matcher's stream.
 */
int hsm_halo_floats(hsm_matcher *m, size_t *count);
int hsm_halo_pack(hsm_matcher *m, float *top, float *bottom);
int hsm_halo_unpack(hsm_matcher *m, const float *from_above, const float *from_below);

/*
 * Row-band mode, fused variant: instead of pack -> NCCL -> unpack, the iteration kernel itself
 * stores its first / last owned row of the new W, N, E into the neighbouring GPU's ghost rows
 * through peer-mapped memory (NVLink), and iterations are ordered between neighbours with
 * stream-ordered flag writes / waits (cuStreamWriteValue32 / cuStreamWaitValue32) on a mailbox in
 * each matcher's arena -- no collective call and no host round trip per iteration.
 *   hsm_peer_export: fills `out` (CUDA IPC handle of the arena + plane offsets); send it to the neighbours.
 *   hsm_peer_attach: side 0 = the band above, 1 = the band below; same_process != 0 uses the address
 *                    directly (two matchers in one process, e.g. tests on a single GPU).
 *   hsm_peer_reset:  zero the iteration counters; every rank calls it after hsm_init_fields and
 *                    synchronises with its neighbours (any host barrier) before the first hsm_iterate.
 */
typedef struct hsm_peer_info {
    unsigned char ipc_handle[64];
    uint64_t off_w[2], off_n[2], off_e[2], off_mailbox, local_base;
    int rows, cols, dp, out_lo, out_hi, device, current, reserved;
} hsm_peer_info;
int hsm_peer_export(hsm_matcher *m, hsm_peer_info *out);
int hsm_peer_attach(hsm_matcher *m, int side, const hsm_peer_info *peer, int same_process);
int hsm_peer_reset(hsm_matcher *m);
int hsm_peer_detach(hsm_matcher *m);

int hsm_get_stats(hsm_matcher *m, hsm_stats *out);
int hsm_reset_stats(hsm_matcher *m);

/*
 * Prior rasterisation on the device (next row N2): replaces split_frames_by_time + generate_frames_from_spikes
 * (stereovis/utils/frames_io.py:184-248), the producer of the matcher's `prior` input
 * (hybrid_stereo.py:169-176).  Frame f collects the events with lo[f] <= t <= hi[f] (hi_inclusive != 0; with
 * pivots: lo = tick - time_interval, hi = tick, frames_io.py:200-202) or lo[f] <= t < hi[f] (the wrap-around
 * frames of frames_io.py:193-198); a pixel gets the value z of the LATEST such event that hit it (numpy's fancy
 * assignment in chronological order, frames_io.py:239-244; among equal timestamps the later array position
 * wins), `fill` if none did.  x / y follow numpy index semantics (negative wraps once, otherwise out of range
 * -> HSM_ERR_ARG, the reference's IndexError).  All inputs are host arrays.  Outputs (either may be NULL):
 *   frames_out     host, n_frames x rows x cols float64 -- what the reference function returns;
 *   labels_device  DEVICE, n_frames x rows x cols int32 -- exact integers as they are, everything else -1
 *                  (`prior == l` of mrf.py:57 is preserved); pass it to hsm_lbp / hsm_init_fields as the prior
 *                  with prior_dtype HSM_I32 and on_device 2 (or 1): the priors never visit the host.
 */
int hsm_rasterise_events(size_t n_events, const int32_t *x, const int32_t *y, const double *t, const double *z,
                         int n_frames, const double *lo, const double *hi, int hi_inclusive, int rows, int cols,
                         double fill, double *frames_out, int32_t *labels_device, int device);

/* Plain device memory for buffers that travel between entry points (e.g. labels_device above). */
int hsm_device_alloc(size_t bytes, int device, void **out);
int hsm_device_free(void *p, int device);
int hsm_device_read(void *dst_host, const void *src_device, size_t bytes, int device);

/*
 * Frame pre-processing (next row N3): per-frame min-max contrast normalisation in float64, replaces
 * normalise_contrast of stereovis/utils/frames_io.py:103-109 (used by load_frames(adjust_contrast=True),
 * hybrid_stereo.py:90-103).  in/out: n_frames x rows x cols float64 host arrays (may alias).
 */
int hsm_normalise_contrast(int n_frames, int rows, int cols, const double *in, double *out, int device);

/*
 * Frame pre-processing (next row N3), down-scaling: replaces load_frames' `downsample`
 * (stereovis/utils/frames_io.py:99-101) = skimage.transform.rescale(img, 1 / factor, preserve_range=True,
 * mode='constant') of scikit-image 0.13.0: bilinear, out-of-image neighbours count as 0, no anti-aliasing,
 * clipped to the input's range.  in: n_frames x rows x cols, out: n_frames x out_rows x out_cols, float64 host
 * arrays.  Parity with scikit-image is unpinned (not installable here); see raster.cuh for the one known
 * deviation (exact instead of least-squares-estimated affine coefficients).
 */
int hsm_rescale_frames(int n_frames, int rows, int cols, const double *in, int out_rows, int out_cols, double *out,
                       int device);

/*
 * Prior adjustment (next row N4): replaces OpticalFlowPixelCorrection.adjust
 * (stereovis/spiking/optical_flow_correction.py:34-88) as written, including the defect its FIXME (:69) names --
 * the shift of a prior pixel is taken from the compass direction of its own (col, row) position, not from the
 * fitted velocity field.  prior: n_frames x rows x cols float64 (values >= 0 move), out: float32, host arrays;
 * time_arrow = +1 / -1.
 */
int hsm_adjust_priors(int n_frames, int rows, int cols, const double *prior, double nondata, int time_arrow, float *out,
                      int device);

/*
 * Event-based visual flow (next row N4): replaces VelocityVectorField._fit_velocity_field
 * (stereovis/spiking/algorithms/vvf.py:32-82) for ONE polarity group: events sorted by time, host arrays t, x, y
 * of n_events float64; velocities: n_events x 2 float64 (x, y slope of the fitted time plane; 0 where the
 * neighbourhood has fewer than min_events events; the minimum-norm solution where it is collinear).  float64
 * throughout; agreement
 * with the reference's SVD-based least squares is to rounding (~1e-9 relative), not bit for bit.
 */
int hsm_fit_velocity_field(size_t n_events, const double *t, const double *x, const double *y, double time_interval,
                           double neighbourhood_x, double neighbourhood_y, double rejection_threshold,
                           double convergence_threshold, int max_iter_steps, int min_events, double *velocities,
                           int device);

/*
 * Self-test of the shared-reciprocal division used by the fused kernel: a4 holds n x 4
 * numerators, b holds n divisors (host arrays); *mismatches = how many of the 4n quotients
 * differ in bits from div.rn.f32 (must be 0).
 */
int hsm_selftest_division(const float *a4, const float *b, size_t n, unsigned long long *mismatches);

/* Pinned host memory for asynchronous transfers. */
int hsm_host_alloc(size_t bytes, void **out);
int hsm_host_free(void *p);
/* *pinned = 1 when `p` lies in page-locked memory known to CUDA (asynchronous copies really are asynchronous) */
int hsm_host_is_pinned(const void *p, int *pinned);

#ifdef __cplusplus
}
#endif
#endif /* HSM_H_ */
