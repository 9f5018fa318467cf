// Host side of the C ABI declared in include/hsm.h.
//
// One hsm_matcher owns: the five planes of the reference's StereoMRF
// (mrf.py:15-19) in the device layout described in kernels.cuh, with the west /
// north / east planes double-buffered (the fused iteration reads the old state
// of neighbouring pixels, so it cannot update in place); the float32 images and
// the int32 prior labels; staging space for callers' float64 / int64 arrays.
// Everything is allocated lazily on the first compute call (fork safety).
// The iteration kernels are launched from their own translation units (matcher.hpp).
#include "matcher.hpp"
#include "raster.cuh"

using namespace hsm;

namespace {
thread_local std::string g_error;

// Every entry point that touches CUDA runs on the matcher's device and puts the caller's current device back
// when it returns (a caller that drives several GPUs from one thread must not find its device changed).
struct DeviceScope {
    int prev = -1, dev;
    // `active` = false: no CUDA call at all (entry points that may run before the first compute call, i.e.
    // possibly in the parent of a fork()ed worker)
    explicit DeviceScope(int device, bool active = true) : dev(device)
    {
        if (!active) return;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceScope()
    {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};
}

namespace hsm {

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

}  // namespace hsm

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t dtype_size(int dtype)
{
    switch (dtype) {
        case HSM_F32: case HSM_I32: return 4;
        case HSM_F64: case HSM_I64: return 8;
        default: return 0;
    }
}

// Lane configuration for a disparity range: LPP lanes per pixel, VPL float4 per lane,
// NT threads per CTA of the fused kernel.
// Table of lane configurations (index = HSM_OPT_LANE_CONFIG - 1).  Entries 0..5 are the
// defaults per disparity range; 6..8 trade lanes per pixel for float4s per lane (fewer
// shuffle steps and per-pixel operations per element, more registers per thread).
const LaneCfg kLaneTable[] = {{4, 1, 256}, {8, 1, 256}, {16, 1, 512}, {32, 1, 512}, {32, 2, 512},
                              {32, 4, 256}, {4, 2, 256}, {8, 2, 256}, {16, 2, 512},
                              // entries 10..13: lanes-per-pixel counts that are not a power of two (default
                              // kernel family only): the reference's shipped disparity counts 20 and 22 are 5 and 6
                              // float4 chunks per pixel
                              {5, 1, 256}, {6, 1, 256}, {5, 2, 256}, {6, 2, 256},
                              // entries 14..16: other CTA widths (7, 9, 10 warps = 42, 45, 40 pixels) for frames whose
                              // width the 8-warp strips divide badly -- the shipped 80- and 70-column frames
                              {5, 1, 224}, {6, 1, 288}, {8, 1, 320}};
constexpr int kLaneTableSize = 16;

bool is_pow2(int v) { return (v & (v - 1)) == 0; }

bool lane_cfg_fits(const LaneCfg &c, int D) { return D <= 4 * c.lpp * c.vpl; }

bool lane_cfg(int D, LaneCfg *c)
{
    int pick;
    if (D <= 16) pick = 0;
    else if (D <= 32) pick = 1;
    else if (D <= 64) pick = 7;        // 8 lanes x 2 float4: measured faster than 16 x 1 at 640x480xD64
    else if (D <= 128) pick = 8;       // 16 lanes x 2 float4: measured faster than 32 x 1 at 480x270xD128
    else if (D <= 256) pick = 4;
    else if (D <= 512) pick = 5;
    else return false;
    *c = kLaneTable[pick];
    return true;
}

// Lane layout of the default kernel family: the power-of-two choice unless 5 or 6 chunk slots per lane group
// waste fewer lanes (D = 17..24 and 33..48; for 25..32 and 49..64 the 8-lane layouts are already as good).
// pixel columns a frame of `cols` columns costs with this layout: strips x pixels per CTA (two of them halo)
long strip_cost(const LaneCfg &c, int cols)
{
    const int px = (c.nt / 32) * (32 / c.lpp), tx = px - 2;
    return (long)((cols + tx - 1) / tx) * px;
}

LaneCfg lane_cfg_st(int D, int cols, const LaneCfg &base)
{
    LaneCfg pick = base;
    if (D > 16 && D <= 20) pick = kLaneTable[9];
    else if (D > 20 && D <= 24) pick = kLaneTable[10];
    else if (D > 32 && D <= 40) pick = kLaneTable[11];
    else if (D > 40 && D <= 48) pick = kLaneTable[12];
    // a CTA width that wastes fewer columns on this frame width, if there is one for the layout
    for (int alt = 13; alt < kLaneTableSize; ++alt) {
        const LaneCfg &c = kLaneTable[alt];
        if (c.lpp == pick.lpp && c.vpl == pick.vpl && 10 * strip_cost(c, cols) < 9 * strip_cost(pick, cols)) pick = c;
    }
    return pick;
}

size_t plane_floats(const hsm_matcher *m) { return (size_t)m->cap * m->H * m->W * m->Dp; }
size_t image_elems(const hsm_matcher *m) { return (size_t)m->cap * m->H * m->Wp; }

size_t arena_layout(hsm_matcher *m, bool assign)
{
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t at = off;
        off = align_up(off + bytes, 256);
        return at;
    };
    char *base = (char *)m->arena;
    const size_t pb = plane_floats(m) * sizeof(float), ib = image_elems(m);
    size_t o;
    o = take(pb); if (assign) m->S = (float *)(base + o);
    for (int i = 0; i < 2; ++i) {
        o = take(pb); if (assign) m->Wm[i] = (float *)(base + o);
        o = take(pb); if (assign) m->Nm[i] = (float *)(base + o);
        o = take(pb); if (assign) m->Em[i] = (float *)(base + o);
    }
    o = take(pb); if (assign) m->Dt = (float *)(base + o);
    o = take(ib * 4); if (assign) m->L = (float *)(base + o);
    o = take(ib * 4); if (assign) m->R = (float *)(base + o);
    o = take(ib * 4); if (assign) m->P = (int *)(base + o);
    for (int i = 0; i < 3; ++i) { o = take(ib * 8); if (assign) m->stage[i] = base + o; }
    o = take(ib * 8); if (assign) m->labels = (long long *)(base + o);
    o = take(hsm_matcher::kMailboxWords * sizeof(unsigned)); if (assign) m->mailbox = (unsigned *)(base + o);
    return off;
}

int ensure_ready(hsm_matcher *m)          // (the caller's DeviceScope has selected m->device)
{
    if (m->ready) return HSM_OK;
    if (m->stream == nullptr) {
        CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->own_stream = true;
    }
    CUDA_TRY(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device));
    m->arena_bytes = arena_layout(m, false);
    CUDA_TRY(cudaMalloc(&m->arena, m->arena_bytes));
    arena_layout(m, true);
    // message planes start at zero (mrf.py:15-19); the data plane is defined by init_fields
    CUDA_TRY(cudaMemsetAsync(m->arena, 0, m->arena_bytes, m->stream));
    m->stats.device_bytes = m->arena_bytes;
    m->ready = true;
    return HSM_OK;
}

}  // namespace

namespace hsm {
Geo geometry(const hsm_matcher *m)
{
    Geo g;
    g.H = m->H; g.W = m->W; g.Wp = m->Wp; g.D = m->D; g.Dp = m->Dp;
    g.vlo = m->vlo; g.vhi = m->vhi;
    g.plane_stride = (size_t)m->H * m->W * m->Dp;
    g.image_stride = (size_t)m->H * m->Wp;
    return g;
}
}  // namespace hsm

namespace {

int pixel_grid(const hsm_matcher *m, int lpp)
{
    const size_t groups = 256 / lpp;
    const size_t n_pix = (size_t)m->frames * m->H * m->W;
    size_t blocks = (n_pix + groups - 1) / groups;
    return (int)std::min<size_t>(blocks, (size_t)m->sm_count * 64);
}

}  // namespace

namespace hsm {
int check_launch(hsm_matcher *m, const char *what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(HSM_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(err));
    m->stats.kernel_launches++;
    return HSM_OK;
}
}  // namespace hsm

namespace {

int ensure_data_plane(hsm_matcher *m)
{
    if (m->dt_valid) return HSM_OK;
    if (!m->have_fields) {      // never initialised: the data plane is all zero (mrf.py:19)
        CUDA_TRY(cudaMemsetAsync(m->Dt, 0, plane_floats(m) * sizeof(float), m->stream));
        m->dt_valid = true;
        return HSM_OK;
    }
    const Geo g = geometry(m);
    const int *prior = m->have_prior ? m->P : nullptr;
    HSM_DISPATCH(m->cfg, {
        (void)NT;
        k_cost_volume<LPP, VPL><<<pixel_grid(m, LPP), 256, 0, m->stream>>>(g, m->frames, m->L, m->R, prior,
                                                                          m->blend, m->Dt);
    });
    HSM_TRY(check_launch(m, "k_cost_volume"));
    m->dt_valid = true;
    return HSM_OK;
}

// Real zeros in memory for the four message planes of the current batch, if they are only logically zero.
int materialise_zero_messages(hsm_matcher *m)
{
    if (!m->zero_messages) return HSM_OK;
    const size_t bytes = (size_t)std::max(m->frames, 1) * m->H * m->W * m->Dp * sizeof(float);
    CUDA_TRY(cudaMemsetAsync(m->S, 0, bytes, m->stream));
    CUDA_TRY(cudaMemsetAsync(m->Wm[m->cur], 0, bytes, m->stream));
    CUDA_TRY(cudaMemsetAsync(m->Nm[m->cur], 0, bytes, m->stream));
    CUDA_TRY(cudaMemsetAsync(m->Em[m->cur], 0, bytes, m->stream));
    m->zero_messages = false;
    return HSM_OK;
}

int resolve_events(hsm_matcher *m)
{
    for (auto &pr : m->pending) {
        CUDA_TRY(cudaEventSynchronize(pr.second));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        m->stats.iteration_ms += ms;
        m->pool.push_back(pr);
    }
    m->pending.clear();
    return HSM_OK;
}

int copy_in(hsm_matcher *m, void *dst, const void *src, size_t bytes, int on_device)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             m->stream));
    return HSM_OK;
}

int load_image(hsm_matcher *m, float *dst, const void *src, int dtype, int slot, size_t n, int on_device)
{
    if (dtype == HSM_F32) {
        if (m->Wp == m->W) return copy_in(m, dst, src, n * 4, on_device);
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)m->Wp * 4, src, (size_t)m->W * 4, (size_t)m->W * 4, n / m->W,
                                   on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, m->stream));
        return HSM_OK;
    }
    const void *dev_src = src;
    if (!on_device) {
        HSM_TRY(copy_in(m, m->stage[slot], src, n * 8, 0));
        dev_src = m->stage[slot];
    }
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)m->sm_count * 16);
    k_to_f32<double><<<blocks, 256, 0, m->stream>>>((const double *)dev_src, dst, n, m->W, m->Wp);
    return check_launch(m, "k_to_f32");
}

int load_prior(hsm_matcher *m, const void *src, int dtype, size_t n, int on_device)
{
    const void *dev_src = src;
    if (!on_device) {
        HSM_TRY(copy_in(m, m->stage[2], src, n * dtype_size(dtype), 0));
        dev_src = m->stage[2];
    }
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)m->sm_count * 16);
    const int W = m->W, Wp = m->Wp;
    switch (dtype) {
        case HSM_F32: k_prior_labels<float><<<blocks, 256, 0, m->stream>>>((const float *)dev_src, m->P, n, m->D, W, Wp); break;
        case HSM_F64: k_prior_labels<double><<<blocks, 256, 0, m->stream>>>((const double *)dev_src, m->P, n, m->D, W, Wp); break;
        case HSM_I32: k_prior_labels<int><<<blocks, 256, 0, m->stream>>>((const int *)dev_src, m->P, n, m->D, W, Wp); break;
        case HSM_I64: k_prior_labels<long long><<<blocks, 256, 0, m->stream>>>((const long long *)dev_src, m->P, n, m->D, W, Wp); break;
        default: return fail(HSM_ERR_ARG, "unknown prior dtype %d", dtype);
    }
    return check_launch(m, "k_prior_labels");
}

}  // namespace

namespace hsm {
int choose_band_rows(const hsm_matcher *m, int strips, int ctas_per_sm)
{
    const int rows = m->ohi - m->olo + 1;
    if (m->band_rows > 0) return std::min(m->band_rows, rows);
    // Every band re-reads two halo rows (cost 2 / band_rows) and a CTA's life is one band, so the machine
    // drains for about half a band at the end of a launch.  Measured optimum (tools/band_sweep.py): about
    // 6 resident-CTA waves per launch when an iteration is one kernel, about 4 over both kernels when the
    // frames run as two concurrent groups (the other group fills the drain, so bands can be taller);
    // beyond ~64 rows the halo saving is below 3 % and taller bands only lengthen the drain.
    // Bands are normally at least 12 rows; when even 12-row bands cannot fill the machine once (a single
    // small frame: latency-, not bandwidth-bound) they may shrink to 4 rows.
    const long per_band = (long)strips * m->frames;
    int min_rows = 12;
    while (min_rows > 4 && per_band * ((rows + min_rows - 1) / min_rows) < (long)m->sm_count * 3) min_rows -= 2;
    // a tiny single frame (the shipped 80x60 configurations): not even one CTA per SM at 4 rows, and an iteration
    // is the latency of one band -- 2-row bands measured 6.4 us per iteration against 8.7 at 4 rows
    if (min_rows == 4 && per_band * ((rows + 3) / 4) < (long)m->sm_count) min_rows = 2;
    const bool grouped = m->launch_frames < m->frames;
    const long target = (long)m->sm_count * std::max(1, ctas_per_sm) * (grouped ? 4 : 6);
    long bands = std::max((target + per_band - 1) / per_band, (long)(rows + 63) / 64);
    bands = std::max(1L, std::min<long>(bands, std::max(1, rows / min_rows)));
    // short frames (the shipped 80x60 / 70x60 sizes in batches): two halo rows are 13 % of a 15-row band, and the
    // kernel is issue-bound there, so half as many bands win as long as two waves of CTAs remain
    // (80x60xD22 x 256: 0.650 -> 0.675 of peak with 30-row bands, 0.60 with one 60-row band)
    const long slots = (long)m->sm_count * std::max(1, ctas_per_sm);
    while (rows <= 64 && bands >= 2 && rows / bands < 30 && per_band * (bands / 2) >= 2 * slots) bands /= 2;
    return (int)((rows + bands - 1) / bands);
}

// Persistent dataflow kernel: there is no drain per iteration and no launch per iteration, so the only costs
// of a band are its two halo rows against the parallelism it removes.  A tile of iteration k+1 needs its
// neighbours of iteration k, handed out one iteration's worth of tiles earlier, so an iteration must be
// several rounds of resident CTAs long or CTAs wait: aim at ~3.5 tiles per resident CTA per iteration,
// bands of 4..64 rows.  (~2.25 tiles, i.e. 16-row instead of 11-row bands at 640x480xD64, wins in short runs -- 0.92
// against 0.89 of peak over 25 iterations -- and loses in long ones: 0.75 against 0.81 over 100 iterations, 0.84
// against 0.85 over 200, and 0.73 against 0.79 in bench.py's 100-iteration `lbp` leg; the reference's C3 runs 100.
// profiles/r2_probe_ab_configs.txt.)
int choose_flow_band_rows(const hsm_matcher *m, int strips, int ctas_per_sm)
{
    const int rows = m->ohi - m->olo + 1;
    if (m->band_rows > 0) return std::min(m->band_rows, rows);
    const long per_band = (long)strips * m->frames;
    const long target = 7L * m->sm_count * std::max(1, ctas_per_sm) / 2;
    long bands = std::max((target + per_band - 1) / per_band, (long)(rows + 63) / 64);
    bands = std::max(1L, std::min<long>(bands, std::max(1, rows / 4)));
    return (int)((rows + bands - 1) / bands);
}
}  // namespace hsm

namespace {

#ifndef HSM_L2_PROMOTION
#define HSM_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_fn(EncodeTiledFn *out)
{
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult query;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &query));
        if (!fn || query != cudaDriverEntryPointSuccess)
            return fail(HSM_ERR_CUDA, "driver does not provide cuTensorMapEncodeTiled");
        cached = (EncodeTiledFn)fn;
    }
    *out = cached;
    return HSM_OK;
}

// 4-D map (Dp, cols, valid rows, frames) over one plane; box = one row segment of `px` pixels.
int make_plane_map(const hsm_matcher *m, CUtensorMap *map, float *plane, int px)
{
    EncodeTiledFn encode;
    HSM_TRY(encode_fn(&encode));
    const cuuint64_t dims[4] = {(cuuint64_t)m->Dp, (cuuint64_t)m->W, (cuuint64_t)(m->vhi - m->vlo + 1),
                                (cuuint64_t)m->cap};
    const cuuint64_t strides[3] = {(cuuint64_t)m->Dp * 4, (cuuint64_t)m->W * m->Dp * 4,
                                   (cuuint64_t)m->H * m->W * m->Dp * 4};
    const cuuint32_t box[4] = {(cuuint32_t)m->Dp, (cuuint32_t)px, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    float *base = plane + (size_t)m->vlo * m->W * m->Dp;
    CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, HSM_L2_PROMOTION,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(HSM_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return HSM_OK;
}

// 4-D map (Dp, cols, 1, 1) over ONE row of a neighbouring GPU's plane (its ghost row; peer-mapped memory)
int make_ghost_map(const hsm_matcher *m, CUtensorMap *map, float *row, int px)
{
    EncodeTiledFn encode;
    HSM_TRY(encode_fn(&encode));
    const cuuint64_t dims[4] = {(cuuint64_t)m->Dp, (cuuint64_t)m->W, 1, 1};
    const cuuint64_t strides[3] = {(cuuint64_t)m->Dp * 4, (cuuint64_t)m->W * m->Dp * 4, (cuuint64_t)m->W * m->Dp * 4};
    const cuuint32_t box[4] = {(cuuint32_t)m->Dp, (cuuint32_t)px, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, row, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(HSM_ERR_CUDA, "cuTensorMapEncodeTiled (ghost row) failed with code %d", (int)rc);
    return HSM_OK;
}

// 3-D map (cols, valid rows, frames) over a device image with row pitch Wp; box = `box` columns.
int make_image_map(const hsm_matcher *m, CUtensorMap *map, void *image, bool is_int, int box)
{
    EncodeTiledFn encode;
    HSM_TRY(encode_fn(&encode));
    const cuuint64_t dims[3] = {(cuuint64_t)m->W, (cuuint64_t)(m->vhi - m->vlo + 1), (cuuint64_t)m->cap};
    const cuuint64_t strides[2] = {(cuuint64_t)m->Wp * 4, (cuuint64_t)m->H * m->Wp * 4};
    const cuuint32_t boxd[3] = {(cuuint32_t)box, 1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    char *base = (char *)image + (size_t)m->vlo * m->Wp * 4;
    CUresult rc = encode(map, is_int ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims,
                         strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(HSM_ERR_CUDA, "cuTensorMapEncodeTiled (image) failed with code %d", (int)rc);
    return HSM_OK;
}

}  // namespace

namespace hsm {
bool tma_supported(const hsm_matcher *m) { return m->Dp <= 256 && m->cfg.lpp * m->cfg.vpl * 4 <= 256; }

IterArgs iter_args(hsm_matcher *m, int strips, bool store_s, int ctas_per_sm)
{
    IterArgs a;
    a.Win = m->Wm[m->cur]; a.Nin = m->Nm[m->cur]; a.Ein = m->Em[m->cur];
    a.Wout = m->Wm[m->cur ^ 1]; a.Nout = m->Nm[m->cur ^ 1]; a.Eout = m->Em[m->cur ^ 1];
    a.Sout = m->S;
    a.Dt = m->Dt; a.L = m->L; a.R = m->R; a.P = m->have_prior ? m->P : nullptr;
    a.blend = m->blend;
    a.olo = m->olo; a.ohi = m->ohi;
    a.band_rows = choose_band_rows(m, strips, ctas_per_sm);
    a.store_s = store_s ? 1 : 0;
    a.has_prior = m->have_prior ? 1 : 0;
    a.frame0 = m->launch_frame0;
    a.zero_in = m->zero_messages ? 1 : 0;
    a.labels = nullptr;
    // -1 = automatic (the launch code decides per kernel); HSM_L2_HINTS overrides it for experiments
    static const int l2_hints_env = getenv("HSM_L2_HINTS") ? atoi(getenv("HSM_L2_HINTS")) : -1;
    a.l2_hints = l2_hints_env;
    for (int side = 0; side < 2; ++side) {
        const hsm_matcher::Peer &p = m->peer[side];
        a.pushW[side] = p.attached ? p.Wm[m->cur ^ 1] : nullptr;
        a.pushN[side] = p.attached ? p.Nm[m->cur ^ 1] : nullptr;
        a.pushE[side] = p.attached ? p.Em[m->cur ^ 1] : nullptr;
        a.push_off[side] = p.ghost_off;
    }
    return a;
}
}  // namespace hsm

namespace {

}  // namespace

namespace hsm {
int ensure_maps(hsm_matcher *m)
{
    if (m->maps_valid) return HSM_OK;
    const int px = m->cfg.nt / m->cfg.lpp;
    for (int i = 0; i < 2; ++i) {
        HSM_TRY(make_plane_map(m, &m->tmW[i], m->Wm[i], px));
        HSM_TRY(make_plane_map(m, &m->tmN[i], m->Nm[i], px));
        HSM_TRY(make_plane_map(m, &m->tmE[i], m->Em[i], px));
    }
    HSM_TRY(make_plane_map(m, &m->tmD, m->Dt, px));
    for (int i = 0; i < 2; ++i) {
        HSM_TRY(make_plane_map(m, &m->tsW[i], m->Wm[i], px - 2));
        HSM_TRY(make_plane_map(m, &m->tsN[i], m->Nm[i], px - 2));
        HSM_TRY(make_plane_map(m, &m->tsE[i], m->Em[i], px - 2));
    }
    // image segments: see StageLayout in iter_tma.cuh (LBOX, RSEG, RBOX)
    const int dcov = 4 * m->cfg.lpp * m->cfg.vpl, rseg = px + dcov + 4;
    const int rbox = rseg <= 256 ? rseg : ((rseg / 2 + 31) / 32 * 32);
    HSM_TRY(make_image_map(m, &m->tmI.L, m->L, false, px + 4));
    HSM_TRY(make_image_map(m, &m->tmI.R, m->R, false, rbox));
    HSM_TRY(make_image_map(m, &m->tmI.P, m->P, true, px + 4));
    // the default kernel family's own set (the same boxes unless its lane layout differs)
    {
        const int px5 = (m->cfg5.nt / 32) * (32 / m->cfg5.lpp);
        for (int i = 0; i < 2; ++i) {
            HSM_TRY(make_plane_map(m, &m->st.tmW[i], m->Wm[i], px5));
            HSM_TRY(make_plane_map(m, &m->st.tmN[i], m->Nm[i], px5));
            HSM_TRY(make_plane_map(m, &m->st.tmE[i], m->Em[i], px5));
            HSM_TRY(make_plane_map(m, &m->st.tsW[i], m->Wm[i], px5 - 2));
            HSM_TRY(make_plane_map(m, &m->st.tsN[i], m->Nm[i], px5 - 2));
            HSM_TRY(make_plane_map(m, &m->st.tsE[i], m->Em[i], px5 - 2));
        }
        HSM_TRY(make_plane_map(m, &m->st.tmD, m->Dt, px5));
        const int px54 = (px5 + 3) / 4 * 4;                     // StageLayout::PX4
        const int dcov5 = 4 * m->cfg5.lpp * m->cfg5.vpl, rseg5 = px54 + dcov5 + 4;
        const int rbox5 = rseg5 <= 256 ? rseg5 : ((rseg5 / 2 + 31) / 32 * 32);
        HSM_TRY(make_image_map(m, &m->st.tmI.L, m->L, false, px54 + 4));
        HSM_TRY(make_image_map(m, &m->st.tmI.R, m->R, false, rbox5));
        HSM_TRY(make_image_map(m, &m->st.tmI.P, m->P, true, px54 + 4));
        for (int side = 0; side < 2; ++side) {
            hsm_matcher::Peer &p = m->peer[side];
            if (!p.attached) continue;
            for (int i = 0; i < 2; ++i) {
                HSM_TRY(make_ghost_map(m, &p.psW[i], p.Wm[i] + p.ghost_off, px5 - 2));
                HSM_TRY(make_ghost_map(m, &p.psN[i], p.Nm[i] + p.ghost_off, px5 - 2));
                HSM_TRY(make_ghost_map(m, &p.psE[i], p.Em[i] + p.ghost_off, px5 - 2));
            }
        }
    }
    m->maps_valid = true;
    return HSM_OK;
}
}  // namespace hsm

namespace {

int launch_directional(hsm_matcher *m)
{
    const Geo g = geometry(m);
    float *S = m->S, *Wc = m->Wm[m->cur], *Nc = m->Nm[m->cur], *Ec = m->Em[m->cur];
    HSM_DISPATCH(m->cfg, {
        (void)NT;
        const int grid = pixel_grid(m, LPP);
        k_sweep<LPP, VPL><<<grid, 256, 0, m->stream>>>(g, m->frames, S, Wc, Nc, Ec, m->Dt, +1, 0);
        HSM_TRY(check_launch(m, "k_sweep south"));
        k_sweep<LPP, VPL><<<grid, 256, 0, m->stream>>>(g, m->frames, Wc, S, Nc, Ec, m->Dt, 0, -1);
        HSM_TRY(check_launch(m, "k_sweep west"));
        k_sweep<LPP, VPL><<<grid, 256, 0, m->stream>>>(g, m->frames, Nc, S, Wc, Ec, m->Dt, -1, 0);
        HSM_TRY(check_launch(m, "k_sweep north"));
        k_sweep<LPP, VPL><<<grid, 256, 0, m->stream>>>(g, m->frames, Ec, S, Wc, Nc, m->Dt, 0, +1);
        HSM_TRY(check_launch(m, "k_sweep east"));
    });
    m->stats.iteration_launches += 4;
    return HSM_OK;
}

int check_frame(const hsm_matcher *m, int frame, int which, int n_planes)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    if (frame < 0 || frame >= m->cap) return fail(HSM_ERR_ARG, "frame %d outside capacity %d", frame, m->cap);
    if (which < 0 || which >= n_planes) return fail(HSM_ERR_ARG, "plane index %d out of range", which);
    return HSM_OK;
}

float *plane_ptr(hsm_matcher *m, int which)
{
    switch (which) {
        case 0: return m->S;
        case 1: return m->Wm[m->cur];
        case 2: return m->Nm[m->cur];
        case 3: return m->Em[m->cur];
        default: return m->Dt;
    }
}

}  // namespace
namespace {

typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

int stream_value_fns(StreamValueFn *wait_fn, StreamValueFn *write_fn)
{
    static StreamValueFn cached_wait = nullptr, cached_write = nullptr;
    if (!cached_wait) {
        void *fw = nullptr, *fr = nullptr;
        cudaDriverEntryPointQueryResult q1, q2;
        CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fw, cudaEnableDefault, &q1));
        CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &fr, cudaEnableDefault, &q2));
        if (!fw || !fr || q1 != cudaDriverEntryPointSuccess || q2 != cudaDriverEntryPointSuccess)
            return fail(HSM_ERR_CUDA, "driver does not provide stream memory operations");
        cached_wait = (StreamValueFn)fw;
        cached_write = (StreamValueFn)fr;
    }
    *wait_fn = cached_wait;
    *write_fn = cached_write;
    return HSM_OK;
}

bool p2p_active(const hsm_matcher *m) { return m->peer[0].attached || m->peer[1].attached; }

// before an iteration: every attached neighbour has completed as many iterations as we have
// (its halo rows for our next input have landed, and it no longer reads the planes we are about
// to push into); after it: tell the neighbours.  Stream-ordered, no kernel spins.
int p2p_before_iteration(hsm_matcher *m)
{
    StreamValueFn wait_fn, write_fn;
    HSM_TRY(stream_value_fns(&wait_fn, &write_fn));
    for (int side = 0; side < 2; ++side)
        if (m->peer[side].attached) {
            CUresult rc = wait_fn((CUstream)m->stream, (CUdeviceptr)(m->mailbox + side), m->p2p_iterations,
                                  CU_STREAM_WAIT_VALUE_GEQ);
            if (rc != CUDA_SUCCESS) return fail(HSM_ERR_CUDA, "cuStreamWaitValue32 failed with code %d", (int)rc);
        }
    return HSM_OK;
}

int p2p_after_iteration(hsm_matcher *m)
{
    StreamValueFn wait_fn, write_fn;
    HSM_TRY(stream_value_fns(&wait_fn, &write_fn));
    m->p2p_iterations++;
    for (int side = 0; side < 2; ++side)
        if (m->peer[side].attached) {
            // default flags: a system-scope fence orders the kernel's peer stores before the flag
            CUresult rc = write_fn((CUstream)m->stream, (CUdeviceptr)m->peer[side].mailbox, m->p2p_iterations,
                                   CU_STREAM_WRITE_VALUE_DEFAULT);
            if (rc != CUDA_SUCCESS) return fail(HSM_ERR_CUDA, "cuStreamWriteValue32 failed with code %d", (int)rc);
        }
    return HSM_OK;
}

}  // namespace

extern "C" {

int hsm_version(void) { return 100; }

const char *hsm_last_error(void) { return g_error.c_str(); }

int hsm_device_count(int *count)
{
    if (!count) return fail(HSM_ERR_ARG, "null count");
    CUDA_TRY(cudaGetDeviceCount(count));
    return HSM_OK;
}

int hsm_create(int rows, int cols, int n_levels, int capacity, int device, hsm_matcher **out)
{
    if (!out) return fail(HSM_ERR_ARG, "null output handle");
    *out = nullptr;
    if (rows < 1 || cols < 1 || n_levels < 1 || capacity < 1)
        return fail(HSM_ERR_ARG, "rows, cols, n_levels and capacity must be positive (got %d, %d, %d, %d)", rows,
                    cols, n_levels, capacity);
    if (n_levels > cols + 1)
        return fail(HSM_ERR_ARG, "n_levels %d > cols + 1 = %d: the reference cannot slice such a pair (mrf.py:55)",
                    n_levels, cols + 1);
    LaneCfg cfg;
    if (!lane_cfg(n_levels, &cfg)) return fail(HSM_ERR_ARG, "n_levels %d > 512 is not supported", n_levels);
    if (capacity > 65535) return fail(HSM_ERR_ARG, "capacity %d > 65535", capacity);
    // element offsets inside one frame are 32-bit in the kernels (row offsets, peer ghost rows)
    if ((unsigned long long)rows * cols * ((n_levels + 3) / 4 * 4) >= (1ull << 32))
        return fail(HSM_ERR_ARG, "rows x cols x n_levels = %d x %d x %d needs >= 2^32 elements per plane", rows, cols,
                    n_levels);
    hsm_matcher *m = new hsm_matcher();
    m->H = rows; m->W = cols; m->Wp = (cols + 3) / 4 * 4; m->D = n_levels; m->Dp = (n_levels + 3) / 4 * 4;
    m->cap = capacity; m->device = device; m->cfg = cfg; m->cfg5 = lane_cfg_st(n_levels, cols, cfg);
    m->vlo = 0; m->vhi = rows - 1; m->olo = 0; m->ohi = rows - 1;
    m->frames = 0;
    *out = m;
    return HSM_OK;
}

int hsm_destroy(hsm_matcher *m)
{
    if (!m) return HSM_OK;
    if (m->ready) {
        DeviceScope scope(m->device);
        cudaStreamSynchronize(m->stream);
        for (int side = 0; side < 2; ++side)
            if (m->peer[side].attached && m->peer[side].ipc_base) cudaIpcCloseMemHandle(m->peer[side].ipc_base);
        for (auto &pr : m->pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        for (auto &pr : m->pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        cudaFree(m->arena);
        if (m->host_stage) cudaFreeHost(m->host_stage);
        if (m->flow_state) cudaFree(m->flow_state);
        if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
        if (m->fork_event) cudaEventDestroy(m->fork_event);
        if (m->join_event) cudaEventDestroy(m->join_event);
        if (m->own_stream) cudaStreamDestroy(m->stream);
    }
    delete m;
    return HSM_OK;
}

int hsm_set_stream(hsm_matcher *m, void *cuda_stream)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device, m->ready);
    if (m->ready && m->own_stream) {
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        CUDA_TRY(cudaStreamDestroy(m->stream));
    }
    m->stream = (cudaStream_t)cuda_stream;
    m->own_stream = false;
    if (m->ready && m->stream == nullptr) {
        CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->own_stream = true;
    }
    return HSM_OK;
}

int hsm_set_option(hsm_matcher *m, int key, int value)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    switch (key) {
        case HSM_OPT_ALGORITHM:
            if (value < 0 || value > 6) return fail(HSM_ERR_ARG, "algorithm must be 0..6");
            m->algorithm = value; return HSM_OK;
        case HSM_OPT_BAND_ROWS:
            if (value < 0) return fail(HSM_ERR_ARG, "band_rows must be >= 0");
            m->band_rows = value; return HSM_OK;
        case HSM_OPT_DATA_TERM:
            if (value != 0 && value != 1) return fail(HSM_ERR_ARG, "data_term must be 0 or 1");
            m->data_term = value; return HSM_OK;
        case HSM_OPT_VALID_LO: case HSM_OPT_VALID_HI: case HSM_OPT_OUT_LO: case HSM_OPT_OUT_HI:
            if (value < 0 || value >= m->H) return fail(HSM_ERR_ARG, "row %d outside [0, %d)", value, m->H);
            m->maps_valid = false;
            if (key == HSM_OPT_VALID_LO) m->vlo = value;
            else if (key == HSM_OPT_VALID_HI) m->vhi = value;
            else if (key == HSM_OPT_OUT_LO) m->olo = value;
            else m->ohi = value;
            return HSM_OK;
        case HSM_OPT_LANE_CONFIG: {
            if (value < 0 || value > kLaneTableSize) return fail(HSM_ERR_ARG, "lane config %d out of range", value);
            LaneCfg c;
            if (value == 0) lane_cfg(m->D, &c); else c = kLaneTable[value - 1];
            if (!lane_cfg_fits(c, m->D)) return fail(HSM_ERR_ARG, "lane config %d covers fewer than %d levels", value, m->D);
            // 0: defaults for both families; a power-of-two 8- / 16-warp entry: both families; 10..16: default family only
            if (is_pow2(c.lpp) && value <= 9) { m->cfg = c; m->cfg5 = value == 0 ? lane_cfg_st(m->D, m->W, c) : c; }
            else m->cfg5 = c;
            m->lane_config = value; m->maps_valid = false;
            return HSM_OK;
        }
        case HSM_OPT_STAGES:
            if (value < 0 || value > 8 || value == 1) return fail(HSM_ERR_ARG, "stages must be 0 (auto) or 2..8");
            m->stages = value; return HSM_OK;
        case HSM_OPT_FRAME_GROUPS:
            if (value < 0 || value > 2) return fail(HSM_ERR_ARG, "frame groups must be 0 (auto), 1 or 2");
            m->frame_groups = value; return HSM_OK;
        default: return fail(HSM_ERR_ARG, "unknown option %d", key);
    }
}

int hsm_get_option(hsm_matcher *m, int key, int *value)
{
    if (!m || !value) return fail(HSM_ERR_ARG, "null argument");
    switch (key) {
        case HSM_OPT_ALGORITHM: *value = m->algorithm; return HSM_OK;
        case HSM_OPT_BAND_ROWS: *value = m->band_rows; return HSM_OK;
        case HSM_OPT_DATA_TERM: *value = m->data_term; return HSM_OK;
        case HSM_OPT_VALID_LO: *value = m->vlo; return HSM_OK;
        case HSM_OPT_VALID_HI: *value = m->vhi; return HSM_OK;
        case HSM_OPT_OUT_LO: *value = m->olo; return HSM_OK;
        case HSM_OPT_OUT_HI: *value = m->ohi; return HSM_OK;
        case HSM_OPT_STAGES: *value = m->stages; return HSM_OK;
        case HSM_OPT_LANE_CONFIG: *value = m->lane_config; return HSM_OK;
        case HSM_OPT_FRAME_GROUPS: *value = m->frame_groups; return HSM_OK;
        default: return fail(HSM_ERR_ARG, "unknown option %d", key);
    }
}

int hsm_device_bytes(int rows, int cols, int n_levels, int capacity, size_t *bytes)
{
    if (!bytes) return fail(HSM_ERR_ARG, "null bytes");
    if (rows < 1 || cols < 1 || n_levels < 1 || capacity < 1 || n_levels > 512 || capacity > 65535 ||
        (unsigned long long)rows * cols * ((n_levels + 3) / 4 * 4) >= (1ull << 32))
        return fail(HSM_ERR_ARG, "hsm_device_bytes: geometry %d x %d x %d x %d is not supported", rows, cols, n_levels,
                    capacity);
    hsm_matcher tmp;
    tmp.H = rows; tmp.W = cols; tmp.Wp = (cols + 3) / 4 * 4; tmp.D = n_levels; tmp.Dp = (n_levels + 3) / 4 * 4;
    tmp.cap = capacity;
    *bytes = arena_layout(&tmp, false);
    return HSM_OK;
}

}  // extern "C"

namespace {
// hsm_init_fields; `release_host` = the caller may reuse its host buffers when this returns (a stream
// synchronisation, which hsm_lbp leaves to its readout)
int init_fields_impl(hsm_matcher *m, int n_frames, const void *left, const void *right, int image_dtype,
                     const void *prior, int prior_dtype, int prior_mode, float keep_f32, double keep_f64,
                     int reinit_messages, int on_device, bool release_host)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device);
    if (!left || !right) return fail(HSM_ERR_ARG, "null image pointer");
    if (n_frames < 1 || n_frames > m->cap)
        return fail(HSM_ERR_ARG, "n_frames %d outside [1, capacity %d]", n_frames, m->cap);
    if (image_dtype != HSM_F32 && image_dtype != HSM_F64)
        return fail(HSM_ERR_ARG, "images must be float32 or float64 (dtype code %d)", image_dtype);
    if (prior_mode < HSM_PRIOR_NONE || prior_mode > HSM_PRIOR_CONST_F64)
        return fail(HSM_ERR_ARG, "unknown prior mode %d", prior_mode);
    if (prior && dtype_size(prior_dtype) == 0) return fail(HSM_ERR_ARG, "unknown prior dtype %d", prior_dtype);
    if (!reinit_messages && m->frames != 0 && n_frames != m->frames)
        return fail(HSM_ERR_ARG, "warm start with %d frames but the state holds %d", n_frames, m->frames);
    HSM_TRY(ensure_ready(m));
    const size_t n = (size_t)n_frames * m->H * m->W;
    m->frames = n_frames;
    // on_device: 0 = host arrays, 1 = everything on the device, else bit 0 = images, bit 1 = prior on the device
    const int images_dev = (on_device & 1), prior_dev = (on_device == 1 || (on_device & 2)) ? 1 : 0;
    HSM_TRY(load_image(m, m->L, left, image_dtype, 0, n, images_dev));
    HSM_TRY(load_image(m, m->R, right, image_dtype, 1, n, images_dev));
    m->have_prior = (prior != nullptr && prior_mode != HSM_PRIOR_NONE);
    m->blend.mode = m->have_prior ? prior_mode : kPriorNone;
    m->blend.keep32 = keep_f32;
    m->blend.keep64 = keep_f64;
    if (m->have_prior) HSM_TRY(load_prior(m, prior, prior_dtype, n, prior_dev));
    if (reinit_messages) {                                   // mrf.py:43-48: logically zero, see matcher.hpp
        m->zero_messages = true;
        m->s_valid = true;
    }
    m->have_fields = true;
    m->dt_valid = false;
    m->labels_valid = false;
    // a host buffer may be reused by the caller as soon as we return
    if (on_device != 1 && release_host) CUDA_TRY(cudaStreamSynchronize(m->stream));
    return HSM_OK;
}
}  // namespace

extern "C" {

int hsm_init_fields(hsm_matcher *m, int n_frames, const void *left, const void *right, int image_dtype,
                    const void *prior, int prior_dtype, int prior_mode, float keep_f32, double keep_f64,
                    int reinit_messages, int on_device)
{
    return init_fields_impl(m, n_frames, left, right, image_dtype, prior, prior_dtype, prior_mode, keep_f32, keep_f64,
                            reinit_messages, on_device, true);
}

int hsm_iterate(hsm_matcher *m, int n_iter, int finalize)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device);
    if (n_iter < 0) return fail(HSM_ERR_ARG, "n_iter %d < 0", n_iter);
    if (!m->have_fields) return fail(HSM_ERR_STATE, "hsm_iterate before hsm_init_fields");
    if (n_iter == 0) return HSM_OK;
    HSM_TRY(ensure_ready(m));
    m->labels_valid = false;
    if (m->algorithm == 0 || m->data_term == 0) HSM_TRY(ensure_data_plane(m));
    // only the TMA-fed kernels can take "all messages are zero" as an out-of-tensor load
    if (!((m->algorithm == 2 || m->algorithm == 5 || m->algorithm == 6) && tma_supported(m)))
        HSM_TRY(materialise_zero_messages(m));
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (m->pending.size() >= 512) HSM_TRY(resolve_events(m));
    if (!m->pool.empty()) { ev = m->pool.back(); m->pool.pop_back(); }
    else { CUDA_TRY(cudaEventCreate(&ev.first)); CUDA_TRY(cudaEventCreate(&ev.second)); }
    CUDA_TRY(cudaEventRecord(ev.first, m->stream));
    m->launch_frame0 = 0; m->launch_frames = m->frames; m->launch_stream = m->stream;
    // algorithm 6: the whole loop is one persistent launch (falls back to 5 when it does not apply)
    // (also chosen automatically for one large frame: with nothing to overlap an iteration's drain with, the
    // persistent kernel is 5-18 % faster from 640x480xD64 up; small frames are latency-bound and lose to it)
    const bool flow_auto = m->algorithm == 5 && m->frames == 1 &&
                           (size_t)m->H * m->W * m->Dp >= (size_t)8000000;
    // With neighbours attached (row bands over several GPUs) algorithm 6 also replaces the stream-ordered
    // flag wait / write around every iteration: the tiles along the band boundary exchange per-strip progress
    // counters with the neighbouring GPUs' kernels directly.  Measured equal to the stream-ordered scheme on
    // 2 and 8 GPUs (1080p D256: 1.11 vs 1.09 and 0.313 vs 0.311 ms per iteration), so it is not the default.
    const bool flow_p2p = p2p_active(m) && m->algorithm == 6 && m->frames == 1;
    if ((p2p_active(m) ? flow_p2p : (m->algorithm == 6 || flow_auto)) && tma_supported(m) && n_iter >= 2) {
        bool launched = false;
        if (p2p_active(m)) HSM_TRY(launch_flow_push_any(m, n_iter, finalize != 0, &launched));
        else HSM_TRY(launch_flow_any(m, n_iter, finalize != 0, &launched));
        if (launched && p2p_active(m)) {
            // whole-iteration counters for a later one-launch-per-iteration call (stream-ordered, once)
            m->p2p_iterations += (unsigned)n_iter - 1u;
            HSM_TRY(p2p_after_iteration(m));
        }
        if (launched) {
            CUDA_TRY(cudaEventRecord(ev.second, m->stream));
            m->pending.push_back(ev);
            m->stats.iterations += (uint64_t)n_iter;
            m->s_valid = finalize != 0;
            m->zero_messages = false;
            return HSM_OK;
        }
    }
    // Two frame groups on two streams (see matcher.hpp): frames are independent, so group B's CTAs fill
    // the SMs that group A's last CTAs leave idle, and vice versa, for the whole loop.
    const bool grouped = m->algorithm != 0 && m->algorithm != 3 && !p2p_active(m) &&
                         m->frames >= 2 &&
                         (m->frame_groups == 2 ||      // automatic: not for launches that are latency-bound anyway
                          (m->frame_groups == 0 && (size_t)m->frames * m->H * m->W * m->Dp >= (size_t)500000));
    if (grouped) {
        if (!m->aux_stream) CUDA_TRY(cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
        if (!m->fork_event) CUDA_TRY(cudaEventCreateWithFlags(&m->fork_event, cudaEventDisableTiming));
        if (!m->join_event) CUDA_TRY(cudaEventCreateWithFlags(&m->join_event, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(m->fork_event, m->stream));
        CUDA_TRY(cudaStreamWaitEvent(m->aux_stream, m->fork_event, 0));
    }
    const int n_groups = grouped ? 2 : 1;
    m->labels_fused = false;
    // one kernel per iteration: let the next iteration's CTAs set up while this one drains (with frame
    // groups the other group's CTAs are the better use of those SMs; the peer-to-peer flags sit between
    // the kernels of a stream, so nothing is adjacent there)
    static const bool pdl_enabled = getenv("HSM_NO_PDL") == nullptr;
    m->launch_pdl = pdl_enabled && !grouped && !p2p_active(m);
    for (int it = 0; it < n_iter; ++it) {
        if (m->algorithm == 0) {
            HSM_TRY(launch_directional(m));
            continue;
        }
        const int left = n_iter - it;
        const bool p2p = p2p_active(m);
        const int cur_before = m->cur;
      for (int group = 0; group < n_groups; ++group) {
        if (grouped) {
            const int half = (m->frames + 1) / 2;
            m->launch_frame0 = group == 0 ? 0 : half;
            m->launch_frames = group == 0 ? half : m->frames - half;
            m->launch_stream = group == 0 ? m->stream : m->aux_stream;
            m->cur = cur_before;                 // both groups read the same generation of the planes
        }
        if (p2p) HSM_TRY(p2p_before_iteration(m));
        struct AfterIteration {          // tell the neighbours when this iteration's launch has been queued
            hsm_matcher *m; bool on;
            ~AfterIteration() { if (on) p2p_after_iteration(m); }
        } after{m, p2p};
        // algorithm 3: pairs of iterations fused in one launch; an odd count starts with a single one
        if (!p2p && m->algorithm == 3 && tma_supported(m) && m->data_term == 1 && left >= 2 && left % 2 == 0) {
            bool launched = false;
            const bool store_s2 = finalize && left == 2;
            HSM_TRY(launch_tma2_any(m, store_s2, &launched));
            if (launched) { ++it; continue; }
        }
        const bool store_s = finalize && (it == n_iter - 1);
        if (p2p && !(m->algorithm == 5 || m->algorithm == 6 || m->algorithm == 2))
            return fail(HSM_ERR_STATE, "peer-to-peer halo push needs algorithm 2 or 5");
        if ((m->algorithm == 5 || m->algorithm == 6) && tma_supported(m)) {
            bool launched = false;
            HSM_TRY(p2p ? launch_tmast_push_any(m, store_s, &launched) : launch_tmast_any(m, store_s, &launched));
            if (launched) continue;
        }
        if (m->algorithm == 4 && tma_supported(m))
            HSM_TRY(launch_ws_any(m, store_s));
        else if (m->algorithm >= 2 && tma_supported(m))
            HSM_TRY(p2p ? launch_tma_push_any(m, store_s) : launch_tma_any(m, store_s));
        else
            HSM_TRY(launch_fused_any(m, store_s));
      }
      m->zero_messages = false;            // (every group of the first iteration has been queued)
    }
    m->launch_frame0 = 0; m->launch_frames = m->frames; m->launch_stream = m->stream;
    if (grouped) {
        CUDA_TRY(cudaEventRecord(m->join_event, m->aux_stream));
        CUDA_TRY(cudaStreamWaitEvent(m->stream, m->join_event, 0));
    }
    CUDA_TRY(cudaEventRecord(ev.second, m->stream));
    m->pending.push_back(ev);
    m->stats.iterations += (uint64_t)n_iter;
    m->s_valid = (m->algorithm == 0) || finalize;
    m->labels_valid = finalize != 0 && m->labels_fused;
    return HSM_OK;
}

int hsm_readout(hsm_matcher *m, int64_t *labels, int on_device, int async)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device);
    if (!labels) return fail(HSM_ERR_ARG, "null labels pointer");
    if (m->frames == 0) return fail(HSM_ERR_STATE, "hsm_readout before hsm_init_fields");
    if (!m->s_valid) return fail(HSM_ERR_STATE, "south plane is stale: last hsm_iterate was not finalized");
    HSM_TRY(ensure_ready(m));
    HSM_TRY(materialise_zero_messages(m));
    const Geo g = geometry(m);
    const bool dt_load = m->dt_valid || !m->have_fields;
    if (dt_load) HSM_TRY(ensure_data_plane(m));
    long long *dst = on_device ? (long long *)labels : m->labels;
    const int *prior = m->have_prior ? m->P : nullptr;
    const size_t label_bytes = (size_t)m->frames * m->H * m->W * sizeof(long long);
    if (m->labels_valid) {                   // fused readout: the finalising launch already wrote m->labels
        if (on_device && dst != m->labels)
            CUDA_TRY(cudaMemcpyAsync(dst, m->labels, label_bytes, cudaMemcpyDeviceToDevice, m->stream));
    } else {
    HSM_DISPATCH(m->cfg, {
        (void)NT;
        const int grid = pixel_grid(m, LPP);
        if (dt_load)
            k_readout<LPP, VPL, true><<<grid, 256, 0, m->stream>>>(g, m->frames, m->S, m->Wm[m->cur], m->Nm[m->cur],
                                                                   m->Em[m->cur], m->Dt, m->L, m->R, prior,
                                                                   m->blend, dst, nullptr);
        else
            k_readout<LPP, VPL, false><<<grid, 256, 0, m->stream>>>(g, m->frames, m->S, m->Wm[m->cur], m->Nm[m->cur],
                                                                    m->Em[m->cur], m->Dt, m->L, m->R, prior,
                                                                    m->blend, dst, nullptr);
    });
    HSM_TRY(check_launch(m, "k_readout"));
    }
    if (!on_device) CUDA_TRY(cudaMemcpyAsync(labels, m->labels, label_bytes, cudaMemcpyDeviceToHost, m->stream));
    if (!async) CUDA_TRY(cudaStreamSynchronize(m->stream));
    return HSM_OK;
}

// hsm_lbp for a SMALL batch of host arrays that the caller waits for (the reference's own call: one frame pair,
// numpy arrays, online_processing.py:128).  Such a call is made of latencies, not bandwidth: pageable copies
// stall the calling thread one by one, and every launch costs microseconds.  Here the inputs are copied into a
// page-locked block owned by the matcher (a memcpy of a few hundred KB), go up as asynchronous copies, ONE
// kernel converts both images and the prior, and the labels come back through the same block.
int lbp_small_host(hsm_matcher *m, int n_frames, const void *left, const void *right, int image_dtype,
                   const void *prior, int prior_dtype, int prior_mode, float keep_f32, double keep_f64, int n_iter,
                   int64_t *labels)
{
    HSM_TRY(ensure_ready(m));
    const size_t n = (size_t)n_frames * m->H * m->W;
    const size_t img_bytes = n * dtype_size(image_dtype);
    const bool with_prior = prior != nullptr && prior_mode != HSM_PRIOR_NONE;
    const size_t prior_bytes = with_prior ? n * dtype_size(prior_dtype) : 0;
    const size_t off_r = align_up(img_bytes, 256), off_p = off_r + align_up(img_bytes, 256);
    const size_t off_out = off_p + align_up(prior_bytes, 256), total = off_out + n * sizeof(int64_t);
    if (m->host_stage_bytes < total) {
        if (m->host_stage) CUDA_TRY(cudaFreeHost(m->host_stage));
        m->host_stage = nullptr; m->host_stage_bytes = 0;
        CUDA_TRY(cudaHostAlloc((void **)&m->host_stage, total, cudaHostAllocMapped));
        m->host_stage_bytes = total;
    }
    char *h = m->host_stage;
    memcpy(h, left, img_bytes);
    memcpy(h + off_r, right, img_bytes);
    if (with_prior) memcpy(h + off_p, prior, prior_bytes);
    // A frame or two: the conversion kernel reads the page-locked block in place (mapped host memory, one PCIe round
    // trip) instead of waiting for three copy-engine transfers of a few tens of KB each, ~7 us apiece.
    const void *src_l = m->stage[0], *src_r = m->stage[1], *src_p = m->stage[2];
    static const bool zero_copy_enabled = getenv("HSM_NO_ZERO_COPY") == nullptr;
    if (zero_copy_enabled && 2 * img_bytes + prior_bytes <= ((size_t)1 << 20)) {
        char *hd = nullptr;
        CUDA_TRY(cudaHostGetDevicePointer((void **)&hd, h, 0));
        src_l = hd; src_r = hd + off_r; src_p = hd + off_p;
    } else {
        CUDA_TRY(cudaMemcpyAsync(m->stage[0], h, img_bytes, cudaMemcpyHostToDevice, m->stream));
        CUDA_TRY(cudaMemcpyAsync(m->stage[1], h + off_r, img_bytes, cudaMemcpyHostToDevice, m->stream));
        if (with_prior) CUDA_TRY(cudaMemcpyAsync(m->stage[2], h + off_p, prior_bytes, cudaMemcpyHostToDevice, m->stream));
    }
    m->frames = n_frames;
    m->have_prior = with_prior;
    m->blend.mode = with_prior ? prior_mode : kPriorNone;
    m->blend.keep32 = keep_f32;
    m->blend.keep64 = keep_f64;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)m->sm_count * 16);
    k_convert_inputs<<<blocks, 256, 0, m->stream>>>(src_l, src_r, image_dtype, with_prior ? src_p : nullptr, prior_dtype,
                                                    m->L, m->R, m->P, n, m->D, m->W, m->Wp);
    HSM_TRY(check_launch(m, "k_convert_inputs"));
    m->zero_messages = true;                                 // mrf.py:165: lbp always re-initialises
    m->labels_valid = false;
    m->s_valid = true;
    m->have_fields = true;
    m->dt_valid = false;
    HSM_TRY(hsm_iterate(m, n_iter, 1));
    HSM_TRY(hsm_readout(m, (int64_t *)m->labels, 1, 1));     // labels stay on the device ...
    CUDA_TRY(cudaMemcpyAsync(h + off_out, m->labels, n * sizeof(int64_t), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));               // ... and come back through the page-locked block
    memcpy(labels, h + off_out, n * sizeof(int64_t));
    return HSM_OK;
}

int hsm_lbp(hsm_matcher *m, int n_frames, const void *left, const void *right, int image_dtype, const void *prior,
            int prior_dtype, int prior_mode, float keep_f32, double keep_f64, int n_iter, int64_t *labels,
            int on_device, int async)
{
    if (m && on_device == 0 && async == 0 && left && right && labels && n_frames >= 1 && n_frames <= m->cap &&
        n_iter >= 0 && (image_dtype == HSM_F32 || image_dtype == HSM_F64) &&
        prior_mode >= HSM_PRIOR_NONE && prior_mode <= HSM_PRIOR_CONST_F64 && (!prior || dtype_size(prior_dtype) != 0) &&
        (size_t)n_frames * m->H * m->W <= (size_t)1 << 19) {
        DeviceScope scope(m->device);
        return lbp_small_host(m, n_frames, left, right, image_dtype, prior, prior_dtype, prior_mode, keep_f32, keep_f64,
                              n_iter, labels);
    }
    // one queue of work, one synchronisation (in the readout, unless async): with an asynchronous readout the
    // host buffers must stay valid until hsm_sync, so only then are they released here
    // async == 2: the caller keeps ALL host buffers valid until hsm_sync -- nothing waits here at all
    HSM_TRY(init_fields_impl(m, n_frames, left, right, image_dtype, prior, prior_dtype, prior_mode, keep_f32, keep_f64,
                             1, on_device, async == 1));
    HSM_TRY(hsm_iterate(m, n_iter, 1));
    return hsm_readout(m, labels, on_device & 1, async);
}

int hsm_sync(hsm_matcher *m)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device, m->ready);
    if (!m->ready) return HSM_OK;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return HSM_OK;
}

int hsm_get_plane(hsm_matcher *m, int frame, int which, float *out)
{
    HSM_TRY(check_frame(m, frame, which, 5));
    DeviceScope scope(m->device);
    if (!out) return fail(HSM_ERR_ARG, "null output");
    HSM_TRY(ensure_ready(m));
    if (which == 0 && !m->s_valid) return fail(HSM_ERR_STATE, "south plane is stale: last hsm_iterate was not finalized");
    HSM_TRY(materialise_zero_messages(m));
    if (which == 4) HSM_TRY(ensure_data_plane(m));
    const float *src = plane_ptr(m, which) + (size_t)frame * m->H * m->W * m->Dp;
    CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)m->D * 4, src, (size_t)m->Dp * 4, (size_t)m->D * 4,
                               (size_t)m->H * m->W, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return HSM_OK;
}

int hsm_set_plane(hsm_matcher *m, int frame, int which, const float *in)
{
    HSM_TRY(check_frame(m, frame, which, 4));
    DeviceScope scope(m->device);
    if (!in) return fail(HSM_ERR_ARG, "null input");
    HSM_TRY(ensure_ready(m));
    if (m->frames <= frame) m->frames = frame + 1;
    m->labels_valid = false;
    HSM_TRY(materialise_zero_messages(m));
    float *dst = plane_ptr(m, which) + (size_t)frame * m->H * m->W * m->Dp;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)m->Dp * 4, in, (size_t)m->D * 4, (size_t)m->D * 4,
                               (size_t)m->H * m->W, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (m->frames <= frame) m->frames = frame + 1;
    return HSM_OK;
}

int hsm_get_energy(hsm_matcher *m, int frame, float *out)
{
    HSM_TRY(check_frame(m, frame, 0, 1));
    DeviceScope scope(m->device);
    if (!out) return fail(HSM_ERR_ARG, "null output");
    if (m->frames == 0) return fail(HSM_ERR_STATE, "hsm_get_energy before hsm_init_fields");
    if (!m->s_valid) return fail(HSM_ERR_STATE, "south plane is stale: last hsm_iterate was not finalized");
    HSM_TRY(ensure_ready(m));
    HSM_TRY(materialise_zero_messages(m));
    HSM_TRY(ensure_data_plane(m));
    // stream-ordered scratch for the energy planes of the current batch
    float *scratch = nullptr;
    const size_t need = (size_t)m->frames * m->H * m->W * m->Dp * sizeof(float);
    CUDA_TRY(cudaMallocAsync((void **)&scratch, need, m->stream));
    const Geo g = geometry(m);
    const int *prior = m->have_prior ? m->P : nullptr;
    HSM_DISPATCH(m->cfg, {
        (void)NT;
        k_readout<LPP, VPL, true><<<pixel_grid(m, LPP), 256, 0, m->stream>>>(
            g, m->frames, m->S, m->Wm[m->cur], m->Nm[m->cur], m->Em[m->cur], m->Dt, m->L, m->R, prior, m->blend,
            m->labels, scratch);
    });
    HSM_TRY(check_launch(m, "k_readout(energy)"));
    const float *src = scratch + (size_t)frame * m->H * m->W * m->Dp;
    CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)m->D * 4, src, (size_t)m->Dp * 4, (size_t)m->D * 4,
                               (size_t)m->H * m->W, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaFreeAsync(scratch, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return HSM_OK;
}


int hsm_peer_export(hsm_matcher *m, hsm_peer_info *out)
{
    if (!m || !out) return fail(HSM_ERR_ARG, "null argument");
    DeviceScope scope(m->device);
    if (m->cap != 1) return fail(HSM_ERR_ARG, "row-band mode needs capacity 1");
    HSM_TRY(ensure_ready(m));
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t handle;
    CUDA_TRY(cudaIpcGetMemHandle(&handle, m->arena));
    static_assert(sizeof handle <= sizeof out->ipc_handle, "ipc handle size");
    memcpy(out->ipc_handle, &handle, sizeof handle);
    const char *base = (const char *)m->arena;
    for (int i = 0; i < 2; ++i) {
        out->off_w[i] = (uint64_t)((const char *)m->Wm[i] - base);
        out->off_n[i] = (uint64_t)((const char *)m->Nm[i] - base);
        out->off_e[i] = (uint64_t)((const char *)m->Em[i] - base);
    }
    out->off_mailbox = (uint64_t)((const char *)m->mailbox - base);
    out->local_base = (uint64_t)(uintptr_t)m->arena;
    out->rows = m->H; out->cols = m->W; out->dp = m->Dp; out->out_lo = m->olo; out->out_hi = m->ohi;
    out->device = m->device; out->current = m->cur;
    return HSM_OK;
}

int hsm_peer_attach(hsm_matcher *m, int side, const hsm_peer_info *peer, int same_process)
{
    if (!m || !peer) return fail(HSM_ERR_ARG, "null argument");
    DeviceScope scope(m->device);
    if (side != 0 && side != 1) return fail(HSM_ERR_ARG, "side must be 0 (above) or 1 (below)");
    if (m->cap != 1) return fail(HSM_ERR_ARG, "row-band mode needs capacity 1");
    if (peer->cols != m->W || peer->dp != m->Dp) return fail(HSM_ERR_ARG, "neighbour has a different row geometry");
    if (peer->current != m->cur) return fail(HSM_ERR_STATE, "neighbour's double buffers are out of phase");
    HSM_TRY(ensure_ready(m));
    hsm_matcher::Peer &p = m->peer[side];
    if (p.attached) return fail(HSM_ERR_STATE, "side %d is already attached", side);
    char *base = nullptr;
    if (same_process) {
        base = (char *)(uintptr_t)peer->local_base;
    } else {
        cudaIpcMemHandle_t handle;
        memcpy(&handle, peer->ipc_handle, sizeof handle);
        CUDA_TRY(cudaIpcOpenMemHandle(&p.ipc_base, handle, cudaIpcMemLazyEnablePeerAccess));
        base = (char *)p.ipc_base;
    }
    for (int i = 0; i < 2; ++i) {
        p.Wm[i] = (float *)(base + peer->off_w[i]);
        p.Nm[i] = (float *)(base + peer->off_n[i]);
        p.Em[i] = (float *)(base + peer->off_e[i]);
    }
    // the band above receives our first owned row in its ghost row below its last owned row, and vice versa;
    // in its mailbox we are the neighbour on the opposite side
    const int ghost_row = (side == 0) ? peer->out_hi + 1 : peer->out_lo - 1;
    if (ghost_row < 0 || ghost_row >= peer->rows) return fail(HSM_ERR_ARG, "neighbour has no ghost row on that side");
    p.ghost_off = (unsigned)ghost_row * (unsigned)m->W * (unsigned)m->Dp;
    p.mailbox = (unsigned *)(base + peer->off_mailbox) + (1 - side);
    p.ghost_flags = (unsigned *)(base + peer->off_mailbox) + hsm_matcher::kGhostBase +
                    (1 - side) * hsm_matcher::kGhostStride;
    p.attached = true;
    m->maps_valid = false;               // the ghost-row store maps are built with the other tensor maps
    return HSM_OK;
}

int hsm_peer_reset(hsm_matcher *m)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device);
    HSM_TRY(ensure_ready(m));
    CUDA_TRY(cudaMemsetAsync(m->mailbox, 0, hsm_matcher::kMailboxWords * sizeof(unsigned), m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    m->p2p_iterations = 0;
    return HSM_OK;
}

int hsm_peer_detach(hsm_matcher *m)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device, m->ready);
    for (int side = 0; side < 2; ++side) {
        hsm_matcher::Peer &p = m->peer[side];
        if (p.attached && p.ipc_base) {
            CUDA_TRY(cudaStreamSynchronize(m->stream));
            CUDA_TRY(cudaIpcCloseMemHandle(p.ipc_base));
        }
        p = hsm_matcher::Peer();
    }
    return HSM_OK;
}

int hsm_halo_floats(hsm_matcher *m, size_t *count)
{
    if (!m || !count) return fail(HSM_ERR_ARG, "null argument");
    *count = (size_t)3 * m->W * m->Dp;
    return HSM_OK;
}

int hsm_halo_pack(hsm_matcher *m, float *top, float *bottom)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device);
    if (m->cap != 1) return fail(HSM_ERR_ARG, "row-band mode needs capacity 1");
    HSM_TRY(ensure_ready(m));
    HSM_TRY(materialise_zero_messages(m));
    const size_t row = (size_t)m->W * m->Dp;
    const float *planes[3] = {m->Wm[m->cur], m->Nm[m->cur], m->Em[m->cur]};
    for (int i = 0; i < 3; ++i) {
        if (top)
            CUDA_TRY(cudaMemcpyAsync(top + i * row, planes[i] + (size_t)m->olo * row, row * 4,
                                     cudaMemcpyDeviceToDevice, m->stream));
        if (bottom)
            CUDA_TRY(cudaMemcpyAsync(bottom + i * row, planes[i] + (size_t)m->ohi * row, row * 4,
                                     cudaMemcpyDeviceToDevice, m->stream));
    }
    return HSM_OK;
}

int hsm_halo_unpack(hsm_matcher *m, const float *from_above, const float *from_below)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device);
    if (m->cap != 1) return fail(HSM_ERR_ARG, "row-band mode needs capacity 1");
    if (from_above && m->olo < 1) return fail(HSM_ERR_ARG, "no ghost row above row %d", m->olo);
    if (from_below && m->ohi + 1 >= m->H) return fail(HSM_ERR_ARG, "no ghost row below row %d", m->ohi);
    HSM_TRY(ensure_ready(m));
    HSM_TRY(materialise_zero_messages(m));
    const size_t row = (size_t)m->W * m->Dp;
    float *planes[3] = {m->Wm[m->cur], m->Nm[m->cur], m->Em[m->cur]};
    for (int i = 0; i < 3; ++i) {
        if (from_above)
            CUDA_TRY(cudaMemcpyAsync(planes[i] + (size_t)(m->olo - 1) * row, from_above + i * row, row * 4,
                                     cudaMemcpyDeviceToDevice, m->stream));
        if (from_below)
            CUDA_TRY(cudaMemcpyAsync(planes[i] + (size_t)(m->ohi + 1) * row, from_below + i * row, row * 4,
                                     cudaMemcpyDeviceToDevice, m->stream));
    }
    return HSM_OK;
}

int hsm_get_stats(hsm_matcher *m, hsm_stats *out)
{
    if (!m || !out) return fail(HSM_ERR_ARG, "null argument");
    DeviceScope scope(m->device, m->ready);
    if (m->ready) HSM_TRY(resolve_events(m));
    *out = m->stats;
    return HSM_OK;
}

int hsm_reset_stats(hsm_matcher *m)
{
    if (!m) return fail(HSM_ERR_ARG, "null matcher");
    DeviceScope scope(m->device, m->ready);
    if (m->ready) HSM_TRY(resolve_events(m));
    const uint64_t bytes = m->stats.device_bytes;
    m->stats = hsm_stats{};
    m->stats.device_bytes = bytes;
    return HSM_OK;
}

namespace {
// device memory and a stream that are released on every return path
struct Scratch {
    std::vector<void *> blocks;
    cudaStream_t stream = nullptr;
    ~Scratch()
    {
        if (stream) cudaStreamSynchronize(stream);
        for (void *p : blocks) cudaFree(p);
        if (stream) cudaStreamDestroy(stream);
    }
    int alloc(void **out, size_t bytes)
    {
        *out = nullptr;
        CUDA_TRY(cudaMalloc(out, std::max<size_t>(bytes, 16)));
        blocks.push_back(*out);
        return HSM_OK;
    }
};
int device_sm_count(int device)
{
    int count = 148;
    cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
    return count;
}
}  // namespace

int hsm_rasterise_events(size_t n_events, const int32_t *x, const int32_t *y, const double *t, const double *z,
                         int n_frames, const double *lo, const double *hi, int hi_inclusive, int rows, int cols,
                         double fill, double *frames_out, int32_t *labels_device, int device)
{
    if (n_frames < 0 || rows < 1 || cols < 1) return fail(HSM_ERR_ARG, "bad raster geometry");
    if ((!frames_out && !labels_device) || (n_events && (!x || !y || !t || !z)) || (n_frames && (!lo || !hi)))
        return fail(HSM_ERR_ARG, "null argument");
    const size_t n_pixels = (size_t)n_frames * rows * cols;
    if (n_pixels == 0) return HSM_OK;
    if (n_events >= 0xffffffffull) return fail(HSM_ERR_ARG, "too many events");
    DeviceScope scope(device);
    Scratch s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    bool monotone = true;
    for (int f = 1; f < n_frames; ++f) monotone = monotone && lo[f] >= lo[f - 1] && hi[f] >= hi[f - 1];
    int *dx = nullptr, *dy = nullptr, *derr = nullptr;
    double *dt = nullptr, *dz = nullptr, *dlo = nullptr, *dhi = nullptr, *dframes = nullptr;
    unsigned long long *latest = nullptr;
    unsigned *winner = nullptr;
    HSM_TRY(s.alloc((void **)&latest, n_pixels * sizeof(unsigned long long)));
    HSM_TRY(s.alloc((void **)&winner, n_pixels * sizeof(unsigned)));
    HSM_TRY(s.alloc((void **)&derr, sizeof(int)));
    HSM_TRY(s.alloc((void **)&dlo, (size_t)n_frames * sizeof(double)));
    HSM_TRY(s.alloc((void **)&dhi, (size_t)n_frames * sizeof(double)));
    HSM_TRY(s.alloc((void **)&dx, n_events * sizeof(int)));
    HSM_TRY(s.alloc((void **)&dy, n_events * sizeof(int)));
    HSM_TRY(s.alloc((void **)&dt, n_events * sizeof(double)));
    HSM_TRY(s.alloc((void **)&dz, n_events * sizeof(double)));
    if (frames_out) HSM_TRY(s.alloc((void **)&dframes, n_pixels * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(latest, 0, n_pixels * sizeof(unsigned long long), s.stream));
    CUDA_TRY(cudaMemsetAsync(winner, 0, n_pixels * sizeof(unsigned), s.stream));
    CUDA_TRY(cudaMemsetAsync(derr, 0, sizeof(int), s.stream));
    CUDA_TRY(cudaMemcpyAsync(dlo, lo, (size_t)n_frames * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CUDA_TRY(cudaMemcpyAsync(dhi, hi, (size_t)n_frames * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    const int sms = device_sm_count(device);
    if (n_events) {
        CUDA_TRY(cudaMemcpyAsync(dx, x, n_events * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(dy, y, n_events * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(dt, t, n_events * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(dz, z, n_events * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        RasterArgs a;
        a.x = dx; a.y = dy; a.t = dt; a.n_events = n_events; a.lo = dlo; a.hi = dhi; a.n_frames = n_frames;
        a.hi_inclusive = hi_inclusive ? 1 : 0; a.monotone = monotone ? 1 : 0; a.rows = rows; a.cols = cols;
        const int blocks = (int)std::min<size_t>((n_events + 255) / 256, (size_t)sms * 16);
        k_raster_latest<<<blocks, 256, 0, s.stream>>>(a, latest, derr);
        CUDA_TRY(cudaGetLastError());
        k_raster_winner<<<blocks, 256, 0, s.stream>>>(a, latest, winner, derr);
        CUDA_TRY(cudaGetLastError());
    }
    const int blocks = (int)std::min<size_t>((n_pixels + 255) / 256, (size_t)sms * 16);
    k_raster_emit<<<blocks, 256, 0, s.stream>>>(winner, dz, n_pixels, fill, dframes, (int *)labels_device);
    CUDA_TRY(cudaGetLastError());
    int error = 0;
    CUDA_TRY(cudaMemcpyAsync(&error, derr, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    if (frames_out)
        CUDA_TRY(cudaMemcpyAsync(frames_out, dframes, n_pixels * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (error) return fail(HSM_ERR_ARG, "event coordinate out of bounds for a %d x %d frame", rows, cols);
    return HSM_OK;
}

int hsm_normalise_contrast(int n_frames, int rows, int cols, const double *in, double *out, int device)
{
    if (n_frames < 0 || rows < 1 || cols < 1) return fail(HSM_ERR_ARG, "bad frame geometry");
    if (n_frames == 0) return HSM_OK;
    if (!in || !out) return fail(HSM_ERR_ARG, "null argument");
    const size_t frame_elems = (size_t)rows * cols, bytes = (size_t)n_frames * frame_elems * sizeof(double);
    DeviceScope scope(device);
    Scratch s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    double *din = nullptr, *dout = nullptr;
    HSM_TRY(s.alloc((void **)&din, bytes));
    HSM_TRY(s.alloc((void **)&dout, bytes));
    CUDA_TRY(cudaMemcpyAsync(din, in, bytes, cudaMemcpyHostToDevice, s.stream));
    k_normalise_contrast<<<n_frames, 1024, 0, s.stream>>>(din, dout, frame_elems);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return HSM_OK;
}

int hsm_rescale_frames(int n_frames, int rows, int cols, const double *in, int out_rows, int out_cols, double *out,
                       int device)
{
    if (n_frames < 0 || rows < 1 || cols < 1 || out_rows < 1 || out_cols < 1) return fail(HSM_ERR_ARG, "bad frame geometry");
    if (n_frames == 0) return HSM_OK;
    if (!in || !out) return fail(HSM_ERR_ARG, "null argument");
    const size_t in_bytes = (size_t)n_frames * rows * cols * sizeof(double);
    const size_t out_bytes = (size_t)n_frames * out_rows * out_cols * sizeof(double);
    DeviceScope scope(device);
    Scratch s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    double *din = nullptr, *dout = nullptr;
    HSM_TRY(s.alloc((void **)&din, in_bytes));
    HSM_TRY(s.alloc((void **)&dout, out_bytes));
    CUDA_TRY(cudaMemcpyAsync(din, in, in_bytes, cudaMemcpyHostToDevice, s.stream));
    k_rescale_bilinear<<<n_frames, 1024, 0, s.stream>>>(din, dout, rows, cols, out_rows, out_cols);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return HSM_OK;
}

int hsm_adjust_priors(int n_frames, int rows, int cols, const double *prior, double nondata, int time_arrow, float *out,
                      int device)
{
    if (n_frames < 0 || rows < 1 || cols < 1) return fail(HSM_ERR_ARG, "bad frame geometry");
    if (time_arrow != 1 && time_arrow != -1) return fail(HSM_ERR_ARG, "time_arrow must be +1 or -1");
    if (n_frames == 0) return HSM_OK;
    if (!prior || !out) return fail(HSM_ERR_ARG, "null argument");
    const size_t n = (size_t)n_frames * rows * cols;
    DeviceScope scope(device);
    Scratch s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    double *din = nullptr;
    float *dout = nullptr;
    HSM_TRY(s.alloc((void **)&din, n * sizeof(double)));
    HSM_TRY(s.alloc((void **)&dout, n * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(din, prior, n * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)device_sm_count(device) * 16);
    k_adjust_priors<<<blocks, 256, 0, s.stream>>>(din, dout, n_frames, rows, cols, nondata, time_arrow);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return HSM_OK;
}

int hsm_fit_velocity_field(size_t n_events, const double *t, const double *x, const double *y, double time_interval,
                           double neighbourhood_x, double neighbourhood_y, double rejection_threshold,
                           double convergence_threshold, int max_iter_steps, int min_events, double *velocities,
                           int device)
{
    if (n_events == 0) return HSM_OK;
    if (!t || !x || !y || !velocities) return fail(HSM_ERR_ARG, "null argument");
    if (max_iter_steps < 0 || min_events < 0) return fail(HSM_ERR_ARG, "negative iteration / event count");
    DeviceScope scope(device);
    Scratch s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    double *dt = nullptr, *dx = nullptr, *dy = nullptr, *dv = nullptr;
    const size_t bytes = n_events * sizeof(double);
    HSM_TRY(s.alloc((void **)&dt, bytes));
    HSM_TRY(s.alloc((void **)&dx, bytes));
    HSM_TRY(s.alloc((void **)&dy, bytes));
    HSM_TRY(s.alloc((void **)&dv, 2 * bytes));
    CUDA_TRY(cudaMemcpyAsync(dt, t, bytes, cudaMemcpyHostToDevice, s.stream));
    CUDA_TRY(cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, s.stream));
    CUDA_TRY(cudaMemcpyAsync(dy, y, bytes, cudaMemcpyHostToDevice, s.stream));
    FlowFitArgs a;
    a.t = dt; a.x = dx; a.y = dy; a.n = n_events; a.dt = time_interval; a.nx = neighbourhood_x; a.ny = neighbourhood_y;
    a.rejection = rejection_threshold; a.convergence = convergence_threshold; a.max_iter = max_iter_steps;
    a.min_events = min_events;
    const int blocks = (int)std::min<size_t>((n_events + 127) / 128, (size_t)device_sm_count(device) * 16);
    k_vvf_fit<<<blocks, 128, 0, s.stream>>>(a, dv);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(velocities, dv, 2 * bytes, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return HSM_OK;
}

int hsm_device_alloc(size_t bytes, int device, void **out)
{
    if (!out) return fail(HSM_ERR_ARG, "null output");
    DeviceScope scope(device);
    CUDA_TRY(cudaMalloc(out, std::max<size_t>(bytes, 16)));
    return HSM_OK;
}

int hsm_device_free(void *p, int device)
{
    if (!p) return HSM_OK;
    DeviceScope scope(device);
    CUDA_TRY(cudaFree(p));
    return HSM_OK;
}

int hsm_device_read(void *dst_host, const void *src_device, size_t bytes, int device)
{
    if (!dst_host || !src_device) return fail(HSM_ERR_ARG, "null argument");
    DeviceScope scope(device);
    CUDA_TRY(cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost));
    return HSM_OK;
}

int hsm_selftest_division(const float *a4, const float *b, size_t n, unsigned long long *mismatches)
{
    if (!a4 || !b || !mismatches) return fail(HSM_ERR_ARG, "null argument");
    float4 *da = nullptr;
    float *db = nullptr;
    unsigned long long *dm = nullptr;
    CUDA_TRY(cudaMalloc((void **)&da, n * sizeof(float4)));
    CUDA_TRY(cudaMalloc((void **)&db, n * sizeof(float)));
    CUDA_TRY(cudaMalloc((void **)&dm, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemcpy(da, a4, n * sizeof(float4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(db, b, n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(dm, 0, sizeof(unsigned long long)));
    k_division_selftest<<<device_sm_count(0) * 8, 256>>>(da, db, dm, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(mismatches, dm, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dm);
    return HSM_OK;
}


int hsm_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(HSM_ERR_ARG, "null output");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return HSM_OK;
}

int hsm_host_is_pinned(const void *p, int *pinned)
{
    if (!pinned) return fail(HSM_ERR_ARG, "null output");
    cudaPointerAttributes attr;
    const cudaError_t err = cudaPointerGetAttributes(&attr, p);
    if (err != cudaSuccess) {                    // (older drivers report unregistered memory as an error)
        cudaGetLastError();
        *pinned = 0;
        return HSM_OK;
    }
    *pinned = (attr.type == cudaMemoryTypeHost) ? 1 : 0;
    return HSM_OK;
}

int hsm_host_free(void *p)
{
    if (p) CUDA_TRY(cudaFreeHost(p));
    return HSM_OK;
}

}  // extern "C"

