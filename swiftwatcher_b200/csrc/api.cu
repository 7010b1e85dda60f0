// C ABI of libswb200 (include/swb200.h): context, buffers, streams, launch order.
// Host side only; the kernels live in fg_bits.cu / morph_mask.cu / ccl.cu /
// stages.cu / synth.cu.  There is no CPU fallback anywhere in this file: every
// compute entry point needs a CUDA device and reports SWB_ERR_CUDA without one.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "swb_internal.cuh"

using namespace swb;

namespace {

thread_local std::string g_error;

constexpr int N_TIMERS = 6;
constexpr int MAX_SUB = 4;   // sub-batches of one submit (see swb_submit)
const char* const TIMER_NAMES[N_TIMERS] = {"fg_bits",   "morph_mask", "ccl_merge",
                                           "ccl_rank",  "ccl_label",  "write_labels"};
// ccl_merge = local + boundary (or init + merge); ccl_rank = root ranking + scans + seg_init;
// ccl_label = props_final (regionprops into the table)

}  // namespace

struct swb_ctx {
    swb_config cfg;
    Geom g;
    MorphCfg morph;
    int X0a;
    int label_elem;
    int cap_rows;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // host submits: sub-batches of frames are filtered on `worker` while later frames are still being copied
    cudaStream_t worker = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_h2d[MAX_SUB] = {};
    bool pipeline = true;
    long long sub_min_px = 64ll << 20;   // least work (pixels) per sub-batch
    int forced_ts = 0;                   // option "temporal_subchunk": frames per temporal sub-chunk of K1 (0 = automatic)
    // device buffers
    uint8_t* in_buf = nullptr;      // host-mode staging: [N-1 + max_frames][h][in_pitch]
    size_t in_buf_bytes = 0;
    long long in_pitch = 0, in_stride = 0;
    uint8_t* hist[2] = {nullptr, nullptr};
    int hist_cur = 0;
    bool hist_has = false;
    uint16_t* raw_bits = nullptr;
    uint32_t* fbits = nullptr;
    uint8_t* mask = nullptr;
    void* labels = nullptr;
    CclBuffers ccl{};
    // uint8 compatibility table (label_mode U8): merged on the device, see launch_u8_merge
    U8Table u8{};
    // RPCA background model (SWB_BG_RPCA)
    RpcaWork rpca{};
    uint8_t* rp_gray = nullptr;     // [n][h*w] cropped gray stack, newest frame first (the reference's column order)
    uint8_t* rp_sparse = nullptr;   // [n][h*w] clip(-E, 0, 255)
    BilateralLut* d_lut = nullptr;
    int rp_iters = 0;
    // host staging (pinned)
    int32_t* h_segoff = nullptr;
    int32_t* h_overflow = nullptr;
    // state of the last submit
    int last_T = 0;
    int last_ts = 0;                // temporal sub-chunk length the filtering kernel used (swb_last_subchunk)
    int64_t rows_guess = 256;       // rows copied speculatively by swb_collect* before the count is known
    swb_segment* fetch_rows = nullptr;   // swb_collect_begin .. swb_collect_end
    int64_t fetch_cap = 0, fetch_guess = 0;
    bool fetching = false;
    bool pending = false;
    const uint8_t* last_frames_dev = nullptr;   // frame 0 (first OUTPUT frame) of full frames on device, or null
    long long last_stride = 0, last_pitch = 0;
    int64_t launches = 0;
    // timing
    bool timing = false;
    cudaEvent_t ev[N_TIMERS + 1] = {};
    bool ev_valid = false;
    std::string error;
};

namespace {

int fail(swb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    g_error = buf;
    return code;
}

#define CU(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(ctx, SWB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

int select_device(swb_ctx* ctx, int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(ctx, SWB_ERR_CUDA, "no CUDA device available (%s); libswb200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return fail(ctx, SWB_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CU(ctx, cudaSetDevice(device));
    return SWB_OK;
}

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
}

void make_geom(int roi_x0, int roi_y0, int roi_x1, int roi_y1, Geom& g, int& X0a) {
    X0a = roi_x0 & ~31;
    g.h = roi_y1 - roi_y0;
    g.w = roi_x1 - roi_x0;
    g.dx = roi_x0 - X0a;
    g.wa = ((roi_x1 + 31) & ~31) - X0a;
    g.wpr_raw = g.wa / 32;
    g.wpr = (g.w + 31) / 32;
    g.mpitch = g.wpr * 32;
    g.BH = (g.h + 1) / 2;
    g.wpr4 = (g.wpr + 3) & ~3;
    g.BW = 16 * g.wpr4;
}

int make_morph(swb_ctx* ctx, int size, int do_open, int do_close, MorphCfg& m) {
    memset(&m, 0, sizeof(m));
    if (size == 0 || (!do_open && !do_close)) {
        m.radius = 1;
        m.n_ops = 0;
        return SWB_OK;
    }
    if (size != 3 && size != 5) return fail(ctx, SWB_ERR_INVALID, "morph_size must be 0, 3 or 5 (got %d)", size);
    m.radius = size / 2;
    int k = 0;
    if (do_open) { m.is_erode[k++] = 1; m.is_erode[k++] = 0; }
    if (do_close) { m.is_erode[k++] = 0; m.is_erode[k++] = 1; }
    m.n_ops = k;
    return SWB_OK;
}

void free_ctx_buffers(swb_ctx* c) {
    cudaFree(c->in_buf);
    cudaFree(c->hist[0]);
    cudaFree(c->hist[1]);
    cudaFree(c->raw_bits);
    cudaFree(c->fbits);
    cudaFree(c->mask);
    cudaFree(c->labels);
    cudaFree(c->ccl.parent);
    cudaFree(c->ccl.rowcount);
    cudaFree(c->ccl.nseg);
    cudaFree(c->ccl.segoff);
    cudaFree(c->ccl.rows);
    cudaFree(c->ccl.parts);
    cudaFree(c->ccl.pcount);
    cudaFree(c->ccl.big_tiles);
    cudaFree(c->ccl.rootlist);
    cudaFree(c->u8.stage);
    cudaFree(c->u8.rows);
    cudaFree(c->u8.nseg);
    cudaFree(c->u8.segoff);
    rpca_free(c->rpca);
    cudaFree(c->rp_gray);
    cudaFree(c->rp_sparse);
    cudaFree(c->d_lut);
    if (c->h_segoff) cudaFreeHost(c->h_segoff);
    if (c->h_overflow) cudaFreeHost(c->h_overflow);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_h2d)
        if (e) cudaEventDestroy(e);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->worker) cudaStreamDestroy(c->worker);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
}

int alloc_ccl(swb_ctx* ctx, CclBuffers& b, const Geom& g, int T, int cap_rows) {
    const size_t nblk = (size_t)T * g.BH * g.BW;
    CU(ctx, dalloc(&b.parent, nblk));
    CU(ctx, dalloc(&b.rowcount, (size_t)2 * T * g.BH));
    CU(ctx, dalloc(&b.nseg, (size_t)T));
    CU(ctx, dalloc(&b.segoff, (size_t)T + 1));
    CU(ctx, dalloc(&b.rows, (size_t)cap_rows));
    b.cap_rows = cap_rows;
    b.cap_parts = 2 * cap_rows + 4096;
    CU(ctx, dalloc(&b.parts, (size_t)b.cap_parts));
    CU(ctx, dalloc(&b.pcount, 2 * MAX_SUB + 1)); // {partials, listed tiles} per sub-batch, then the overflow flag
    b.overflow = b.pcount + 2 * MAX_SUB;
    CU(ctx, dalloc(&b.big_tiles, (size_t)T * ((g.BH + 7) / 8)));   // tiles are at least 8 block rows tall
    CU(ctx, dalloc(&b.rootlist, (size_t)cap_rows));
    return SWB_OK;
}

}  // namespace

namespace {
struct DevTmp {
    std::vector<void*> ptrs;
    ~DevTmp() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t alloc(T** p, size_t count) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};
}  // namespace


extern "C" {

const char* swb_version(void) { return "swb200 0.2 (sm_100a)"; }

const char* swb_last_error(const swb_ctx* ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

int swb_device_count(int32_t* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess) return fail(nullptr, SWB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return SWB_OK;
}

int swb_create(const swb_config* cfg, swb_ctx** out) {
    if (!cfg || !out) return fail(nullptr, SWB_ERR_INVALID, "null argument");
    *out = nullptr;
    swb_config c = *cfg;
    if (c.frame_h <= 0 || c.frame_w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad frame size %dx%d", c.frame_w, c.frame_h);
    if (c.channels != 1 && c.channels != 3) return fail(nullptr, SWB_ERR_INVALID, "channels must be 1 or 3");
    if (c.frame_pitch == 0) c.frame_pitch = (int64_t)c.frame_w * c.channels;
    if (c.frame_stride == 0) c.frame_stride = (int64_t)c.frame_h * c.frame_pitch;
    if (c.frame_pitch < (int64_t)c.frame_w * c.channels || c.frame_stride < c.frame_pitch * c.frame_h)
        return fail(nullptr, SWB_ERR_INVALID, "frame_pitch / frame_stride too small");
    if (c.roi_x0 == 0 && c.roi_x1 == 0 && c.roi_y0 == 0 && c.roi_y1 == 0) {
        c.roi_x1 = c.frame_w;
        c.roi_y1 = c.frame_h;
    }
    if (c.roi_x0 < 0 || c.roi_y0 < 0 || c.roi_x1 > c.frame_w || c.roi_y1 > c.frame_h || c.roi_x1 <= c.roi_x0 ||
        c.roi_y1 <= c.roi_y0)
        return fail(nullptr, SWB_ERR_INVALID, "ROI [(%d,%d),(%d,%d)] not inside the %dx%d frame", c.roi_x0, c.roi_y0,
                    c.roi_x1, c.roi_y1, c.frame_w, c.frame_h);
    if (c.median_n < 1 || c.median_n > 9 || (c.median_n & 1) == 0)
        return fail(nullptr, SWB_ERR_INVALID, "median_n must be odd and in 1..9 (got %d)", c.median_n);
    if (c.threshold < 0 || c.threshold > 255) return fail(nullptr, SWB_ERR_INVALID, "threshold must be in 0..255");
    if (c.label_mode != SWB_LABELS_I32 && c.label_mode != SWB_LABELS_U8)
        return fail(nullptr, SWB_ERR_INVALID, "bad label_mode %d", c.label_mode);
    if (c.max_frames <= 0 || c.max_frames > 32768) return fail(nullptr, SWB_ERR_INVALID, "max_frames must be in 1..32768");
    if (c.bg_model != SWB_BG_MEDIAN && c.bg_model != SWB_BG_RPCA) return fail(nullptr, SWB_ERR_INVALID, "bad bg_model %d", c.bg_model);
    if (c.bg_model == SWB_BG_RPCA && c.max_frames > 32)
        return fail(nullptr, SWB_ERR_INVALID, "bg_model RPCA decomposes one batch per submit: max_frames must be <= 32 (got %d)", c.max_frames);
    if (c.max_segments <= 0) c.max_segments = 1024 * c.max_frames;

    swb_ctx* ctx = new swb_ctx();
    ctx->cfg = c;
    int rc = make_morph(ctx, c.morph_size, c.do_open, c.do_close, ctx->morph);
    if (rc != SWB_OK) { delete ctx; return rc; }
    make_geom(c.roi_x0, c.roi_y0, c.roi_x1, c.roi_y1, ctx->g, ctx->X0a);
    ctx->label_elem = (c.label_mode == SWB_LABELS_U8) ? 1 : 4;
    ctx->cap_rows = c.max_segments;

    auto bail = [&](int code) {
        std::string msg = ctx->error;
        free_ctx_buffers(ctx);
        delete ctx;
        g_error = msg;
        return code;
    };
    rc = select_device(ctx, c.device);
    if (rc != SWB_OK) return bail(rc);

    const Geom& g = ctx->g;
    const int T = c.max_frames;
    const int nh = c.median_n - 1;
#define CUB(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            fail(ctx, SWB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));               \
            return bail(SWB_ERR_CUDA);                                                              \
        }                                                                                           \
    } while (0)
    CUB(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CUB(cudaStreamCreateWithFlags(&ctx->worker, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    for (auto& e : ctx->ev_h2d) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (nh > 0) {
        CUB(dalloc(&ctx->hist[0], (size_t)nh * g.h * g.wa));
        CUB(dalloc(&ctx->hist[1], (size_t)nh * g.h * g.wa));
    }
    CUB(dalloc(&ctx->raw_bits, (size_t)T * g.h * g.wpr_raw * 2 + 8));
    CUB(dalloc(&ctx->fbits, (size_t)T * g.h * g.wpr4));
    if (c.out_flags & SWB_OUT_MASK) CUB(dalloc(&ctx->mask, (size_t)T * g.h * g.mpitch));
    if (c.out_flags & SWB_OUT_LABELS)
        CUB(cudaMalloc(&ctx->labels, (size_t)T * g.h * g.mpitch * ctx->label_elem));
    rc = alloc_ccl(ctx, ctx->ccl, g, T, ctx->cap_rows);
    if (rc != SWB_OK) return bail(rc);
    if (c.label_mode == SWB_LABELS_U8) {
        ctx->u8.cap = (int)std::min<long long>(ctx->cap_rows, 255ll * T);
        CUB(dalloc(&ctx->u8.stage, (size_t)255 * T));
        CUB(dalloc(&ctx->u8.rows, (size_t)ctx->u8.cap));
        CUB(dalloc(&ctx->u8.nseg, (size_t)T));
        CUB(dalloc(&ctx->u8.segoff, (size_t)T + 1));
    }
    if (c.bg_model == SWB_BG_RPCA) {
        const long long P = (long long)g.h * g.w;
        CUB(rpca_alloc(ctx->rpca, P, T));
        CUB(dalloc(&ctx->rp_gray, (size_t)P * T));
        CUB(dalloc(&ctx->rp_sparse, (size_t)P * T));
        CUB(cudaMalloc(reinterpret_cast<void**>(&ctx->d_lut), sizeof(BilateralLut)));
        BilateralLut lut;
        bilateral_lut(7, 15.0, 1.0, lut);            // data_structures.py:194: bilateral_blur(frame, 7, 15, 1)
        CUB(cudaMemcpy(ctx->d_lut, &lut, sizeof(lut), cudaMemcpyHostToDevice));
    }
    CUB(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_segoff), ((size_t)T + 1) * sizeof(int32_t)));
    CUB(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_overflow), sizeof(int32_t)));
    for (auto& e : ctx->ev) CUB(cudaEventCreate(&e));
#undef CUB
    *out = ctx;
    return SWB_OK;
}

int swb_destroy(swb_ctx* ctx) {
    if (!ctx) return SWB_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    free_ctx_buffers(ctx);
    delete ctx;
    return SWB_OK;
}

int swb_reset(swb_ctx* ctx) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    ctx->hist_has = false;
    return SWB_OK;
}

int swb_set_stream(swb_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return SWB_OK;
}

int swb_submit(swb_ctx* ctx, const uint8_t* frames, int32_t n_frames, int32_t n_halo, int32_t mem_kind) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    const swb_config& c = ctx->cfg;
    const Geom& g = ctx->g;
    const int N = c.median_n;
    if (!frames) return fail(ctx, SWB_ERR_INVALID, "null frames");
    if (n_frames <= 0 || n_frames > c.max_frames)
        return fail(ctx, SWB_ERR_CAPACITY, "n_frames %d outside 1..max_frames (%d)", n_frames, c.max_frames);
    if (n_halo != SWB_HALO_CARRY && (n_halo < 0 || n_halo > N - 1))
        return fail(ctx, SWB_ERR_INVALID, "n_halo %d outside 0..%d", n_halo, N - 1);
    if (mem_kind != SWB_MEM_HOST && mem_kind != SWB_MEM_DEVICE) return fail(ctx, SWB_ERR_INVALID, "bad mem_kind");
    if (ctx->fetching) return fail(ctx, SWB_ERR_STATE, "swb_submit between swb_collect_begin and swb_collect_end");
    CU(ctx, cudaSetDevice(c.device));
    cudaStream_t s = ctx->stream;
    const int inline_halo = n_halo == SWB_HALO_CARRY ? 0 : n_halo;
    const int n_total = inline_halo + n_frames;
    const int C = c.channels;

    // ---- sub-batches (host frames only).  The PCIe copy of a large submit takes ~30x longer than
    // filtering it, so the submit is cut into up to MAX_SUB sub-batches of frames: the copies stay
    // in order on the caller's stream and sub-batch b is filtered on the worker stream as soon as
    // its frames have landed, while b+1 is still in flight.  Frames are independent given their N-1
    // predecessors (read in place), so the only coupling is the running offset into the segment
    // table (CclChain).  Results are identical to one batch.  (Device-resident submits stay one
    // batch: running two sub-batches side by side on two streams measured 4-8 % slower, the
    // bandwidth-bound kernels only get in each other's way.)
    const bool from_host = (mem_kind == SWB_MEM_HOST);
    int tsub = n_frames, nsub = 1;
    if (c.bg_model == SWB_BG_RPCA && n_frames > ctx->rpca.nmax)
        return fail(ctx, SWB_ERR_CAPACITY, "RPCA batch of %d frames exceeds max_frames", n_frames);
    if (from_host && ctx->pipeline && !ctx->timing && c.bg_model == SWB_BG_MEDIAN) {
        const long long px = (long long)g.h * g.wa;
        const long long min_frames = (ctx->sub_min_px + px - 1) / px;   // >= 64 Mpx of work per sub-batch
        long long ts = std::max<long long>({(n_frames + MAX_SUB - 1) / MAX_SUB, min_frames, N - 1, 6});
        ts = (ts + 5) / 6 * 6;
        if (ts < n_frames) {
            tsub = (int)ts;
            nsub = (n_frames + tsub - 1) / tsub;
        }
    }

    FrameSrc src{};
    bool aligned;
    if (from_host) {
        // stage only the (32-pixel aligned) ROI columns / rows of every frame
        const long long in_pitch = (long long)g.wa * C;
        const long long in_stride = in_pitch * g.h;
        const size_t need = (size_t)in_stride * (c.max_frames + N - 1) + 256;
        if (!ctx->in_buf) {
            CU(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->in_buf), need));
            ctx->in_buf_bytes = need;
            ctx->in_pitch = in_pitch;
            ctx->in_stride = in_stride;
        }
        src.cur = ctx->in_buf + (long long)inline_halo * in_stride;
        src.frame_stride = in_stride;
        src.pitch = in_pitch;
        src.avail_w = g.wa;
        aligned = true;
        const bool full = (ctx->X0a == 0 && c.roi_y0 == 0 && g.h == c.frame_h && g.wa >= c.frame_w);
        ctx->last_frames_dev = full ? src.cur : nullptr;
        ctx->last_stride = in_stride;
        ctx->last_pitch = in_pitch;
    } else {
        const uint8_t* f0 = frames + (long long)inline_halo * c.frame_stride;
        src.cur = f0 + (long long)c.roi_y0 * c.frame_pitch + (long long)ctx->X0a * C;
        src.frame_stride = c.frame_stride;
        src.pitch = c.frame_pitch;
        src.avail_w = c.frame_w - ctx->X0a;
        aligned = (reinterpret_cast<uintptr_t>(src.cur) % 16 == 0) && (c.frame_pitch % 16 == 0) &&
                  (c.frame_stride % 16 == 0) && (ctx->X0a + g.wa <= c.frame_w);
        ctx->last_frames_dev = f0;
        ctx->last_stride = c.frame_stride;
        ctx->last_pitch = c.frame_pitch;
    }
    src.n_inline_halo = inline_halo;
    src.hist_valid = (n_halo == SWB_HALO_CARRY && ctx->hist_has && N > 1) ? 1 : 0;
    src.hist = ctx->hist[ctx->hist_cur];
    src.hist_out = (N > 1) ? ctx->hist[ctx->hist_cur ^ 1] : nullptr;

    // host frames [a, a + n) of this call (halo frames included) -> staging buffer
    auto stage_in = [&](int a, int n, cudaStream_t st) -> cudaError_t {
        const int copy_px = std::min(g.wa, c.frame_w - ctx->X0a);
        const uint8_t* sbase = frames + (long long)c.roi_y0 * c.frame_pitch + (long long)ctx->X0a * C +
                               (long long)a * c.frame_stride;
        uint8_t* dbase = ctx->in_buf + (long long)a * ctx->in_stride;
        if (copy_px == c.frame_w && g.h == c.frame_h && c.frame_pitch == ctx->in_pitch && c.frame_stride == ctx->in_stride)
            return cudaMemcpyAsync(dbase, frames + (long long)a * c.frame_stride, (size_t)ctx->in_stride * n,
                                   cudaMemcpyHostToDevice, st);
        if (c.frame_stride % c.frame_pitch == 0) {
            cudaMemcpy3DParms p;
            memset(&p, 0, sizeof(p));
            p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(sbase), (size_t)c.frame_pitch,
                                           (size_t)c.frame_pitch, (size_t)(c.frame_stride / c.frame_pitch));
            p.dstPtr = make_cudaPitchedPtr(dbase, (size_t)ctx->in_pitch, (size_t)ctx->in_pitch, (size_t)g.h);
            p.extent = make_cudaExtent((size_t)copy_px * C, (size_t)g.h, (size_t)n);
            p.kind = cudaMemcpyHostToDevice;
            return cudaMemcpy3DAsync(&p, st);
        }
        for (int f = 0; f < n; ++f) {
            cudaError_t e = cudaMemcpy2DAsync(dbase + (long long)f * ctx->in_stride, (size_t)ctx->in_pitch,
                                              sbase + (long long)f * c.frame_stride, (size_t)c.frame_pitch,
                                              (size_t)copy_px * C, (size_t)g.h, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };

    int launches = 0;
    if (nsub == 1) {
        if (from_host) CU(ctx, stage_in(0, n_total, s));
        ccl_prepare(s, n_frames, g, ctx->ccl, false);
        if (ctx->timing) CU(ctx, cudaEventRecord(ctx->ev[0], s));
        Geom gm = g;                                  // geometry of the morphology input bits
        if (c.bg_model == SWB_BG_RPCA) {
            // crop + gray (newest frame first, the reference's column order, data_structures.py:134,160) ->
            // IALM -> bilateral + threshold -> one bit per pixel, ROI-local columns
            const uint8_t* f0 = src.cur + (long long)g.dx * C;      // src.cur is 32-px aligned: step to the ROI column
            CU(ctx, launch_crop_gray(s, f0, src.frame_stride, src.pitch, C, 0, 0, g.h, g.w, n_frames, 1, ctx->rp_gray));
            CU(ctx, rpca_run(s, ctx->rp_gray, n_frames, (long long)g.h * g.w, ctx->rpca, ctx->rp_sparse, &ctx->rp_iters,
                             &launches));
            gm.dx = 0;
            gm.wpr_raw = g.wpr;
            gm.wa = g.wpr * 32;
            CU(ctx, launch_bilateral(s, ctx->rp_sparse, n_frames, g.h, g.w, ctx->d_lut, 1, nullptr, c.threshold,
                                     reinterpret_cast<uint32_t*>(ctx->raw_bits), gm.wpr_raw, 3));
            launches += 2;
        } else {
            CU(ctx, launch_fg_bits(s, src, C, N, n_frames, g, c.threshold, ctx->raw_bits, aligned, &launches, c.gpu_share, ctx->forced_ts));
        }
        if (ctx->timing) CU(ctx, cudaEventRecord(ctx->ev[1], s));
        CU(ctx, launch_morph_mask(s, reinterpret_cast<const uint32_t*>(ctx->raw_bits), n_frames, gm, ctx->morph,
                                  ctx->fbits, ctx->mask, &launches));
        if (ctx->timing) CU(ctx, cudaEventRecord(ctx->ev[2], s));
        CU(ctx, launch_ccl(s, ctx->fbits, n_frames, g, ctx->ccl, ctx->labels, ctx->label_elem, &launches,
                           ctx->timing ? &ctx->ev[3] : nullptr, 4, nullptr, true));
        if (c.label_mode == SWB_LABELS_U8) CU(ctx, launch_u8_merge(s, n_frames, ctx->ccl, ctx->u8, nullptr, &launches));
        ctx->last_ts = (c.bg_model == SWB_BG_MEDIAN) ? last_temporal_subchunk() : n_frames;
        ctx->ev_valid = ctx->timing;
    } else {
        cudaStream_t sw = ctx->worker;
        CU(ctx, cudaMemsetAsync(ctx->ccl.overflow, 0, sizeof(int32_t), s));
        CU(ctx, cudaEventRecord(ctx->ev_fork, s));
        CU(ctx, cudaStreamWaitEvent(sw, ctx->ev_fork, 0));
        const int cap_parts_sub = ctx->ccl.cap_parts / nsub;
        for (int b = 0; b < nsub; ++b) {
            const int f0 = b * tsub;
            const int nb = std::min(tsub, n_frames - f0);
            const int a = (b == 0) ? 0 : inline_halo + f0;
            CU(ctx, stage_in(a, inline_halo + f0 + nb - a, s));
            CU(ctx, cudaEventRecord(ctx->ev_h2d[b], s));
            CU(ctx, cudaStreamWaitEvent(sw, ctx->ev_h2d[b], 0));
            FrameSrc sb = src;
            if (b > 0) {                    // the N-1 frames before f0 are in place (tsub >= N-1)
                sb.cur = src.cur + (long long)f0 * src.frame_stride;
                sb.n_inline_halo = N - 1;
                sb.hist_valid = 0;
            }
            if (b < nsub - 1) sb.hist_out = nullptr;
            uint16_t* raw_b = ctx->raw_bits + (size_t)f0 * g.h * g.wpr_raw * 2;
            uint32_t* fbits_b = ctx->fbits + (size_t)f0 * g.h * g.wpr4;
            uint8_t* mask_b = ctx->mask ? ctx->mask + (size_t)f0 * g.h * g.mpitch : nullptr;
            void* labels_b = ctx->labels
                                 ? static_cast<uint8_t*>(ctx->labels) + (size_t)f0 * g.h * g.mpitch * ctx->label_elem
                                 : nullptr;
            CclBuffers cb = ctx->ccl;
            cb.parent += (size_t)f0 * g.BH * g.BW;
            cb.rowcount += (size_t)2 * f0 * g.BH;
            cb.nseg += f0;
            cb.segoff += f0;
            cb.parts += (size_t)b * cap_parts_sub;
            cb.cap_parts = cap_parts_sub;
            cb.pcount += 2 * b;
            cb.big_tiles += (size_t)f0 * ((g.BH + 7) / 8);
            CclChain chain;
            chain.frame_base = f0;
            chain.segoff_base = (b > 0) ? ctx->ccl.segoff + f0 : nullptr;   // left there by sub-batch b-1
            ccl_prepare(sw, nb, g, cb, true);
            CU(ctx, launch_fg_bits(sw, sb, C, N, nb, g, c.threshold, raw_b, aligned, &launches, c.gpu_share, ctx->forced_ts));
            CU(ctx, launch_morph_mask(sw, reinterpret_cast<const uint32_t*>(raw_b), nb, g, ctx->morph, fbits_b, mask_b,
                                      &launches));
            CU(ctx, launch_ccl(sw, fbits_b, nb, g, cb, labels_b, ctx->label_elem, &launches, nullptr, 0, &chain, true));
            if (c.label_mode == SWB_LABELS_U8) {
                U8Table ub = ctx->u8;
                ub.stage += (size_t)255 * f0;
                ub.nseg += f0;
                ub.segoff += f0;
                CU(ctx, launch_u8_merge(sw, nb, cb, ub, b > 0 ? ctx->u8.segoff + f0 : nullptr, &launches));
            }
            ctx->last_ts = last_temporal_subchunk();
        }
        CU(ctx, cudaEventRecord(ctx->ev_join, sw));
        CU(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
        ctx->ev_valid = false;
    }
    ctx->launches += launches;
    if (N > 1 && c.bg_model == SWB_BG_MEDIAN) {
        ctx->hist_cur ^= 1;
        ctx->hist_has = true;
    }
    ctx->last_T = n_frames;
    ctx->pending = true;
    return SWB_OK;
}

int swb_sync(swb_ctx* ctx) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SWB_OK;
}

// swb_collect / swb_collect_all / swb_collect_begin + swb_collect_end: the table (and optionally the dense
// outputs) of the last submit with ONE stream synchronisation in the common case.  The number of rows is
// only known on the device, so a guessed number of rows (twice the last submit's) is copied along with
// the offsets; a second copy follows only when the guess was too small.
static int collect_begin_impl(swb_ctx* ctx, swb_segment* rows, int64_t cap, uint8_t* masks, void* labels) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "swb_collect without a preceding swb_submit");
    if (masks && !ctx->mask) return fail(ctx, SWB_ERR_STATE, "masks were not enabled in swb_config.out_flags");
    if (labels && !ctx->labels) return fail(ctx, SWB_ERR_STATE, "labels were not enabled in swb_config.out_flags");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t s = ctx->stream;
    const Geom& g = ctx->g;
    const int T = ctx->last_T;
    const bool u8 = ctx->cfg.label_mode == SWB_LABELS_U8;
    const int32_t* d_segoff = u8 ? ctx->u8.segoff : ctx->ccl.segoff;
    const swb_segment* d_rows = u8 ? ctx->u8.rows : ctx->ccl.rows;
    const int64_t dev_cap = u8 ? ctx->u8.cap : ctx->cap_rows;
    CU(ctx, cudaMemcpyAsync(ctx->h_segoff, d_segoff, ((size_t)T + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaMemcpyAsync(ctx->h_overflow, ctx->ccl.overflow, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    const int64_t guess = rows ? std::min<int64_t>({cap, dev_cap, ctx->rows_guess}) : 0;
    if (guess > 0)
        CU(ctx, cudaMemcpyAsync(rows, d_rows, (size_t)guess * sizeof(swb_segment), cudaMemcpyDeviceToHost, s));
    if (masks)
        CU(ctx, cudaMemcpy2DAsync(masks, (size_t)g.w, ctx->mask, (size_t)g.mpitch, (size_t)g.w, (size_t)T * g.h,
                                  cudaMemcpyDeviceToHost, s));
    if (labels)
        CU(ctx, cudaMemcpy2DAsync(labels, (size_t)g.w * ctx->label_elem, ctx->labels, (size_t)g.mpitch * ctx->label_elem,
                                  (size_t)g.w * ctx->label_elem, (size_t)T * g.h, cudaMemcpyDeviceToHost, s));
    ctx->fetch_rows = rows;
    ctx->fetch_cap = cap;
    ctx->fetch_guess = guess;
    ctx->fetching = true;
    return SWB_OK;
}

static int collect_end_impl(swb_ctx* ctx, int64_t* n_rows, int32_t* per_frame_counts) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    if (!ctx->fetching) return fail(ctx, SWB_ERR_STATE, "swb_collect_end without swb_collect_begin");
    ctx->fetching = false;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t s = ctx->stream;
    const int T = ctx->last_T;
    const bool u8 = ctx->cfg.label_mode == SWB_LABELS_U8;
    const swb_segment* d_rows = u8 ? ctx->u8.rows : ctx->ccl.rows;
    const int64_t dev_cap = u8 ? ctx->u8.cap : ctx->cap_rows;
    swb_segment* rows = ctx->fetch_rows;
    const int64_t cap = ctx->fetch_cap, guess = ctx->fetch_guess;
    CU(ctx, cudaStreamSynchronize(s));
    const int64_t total = ctx->h_segoff[T];
    if (*ctx->h_overflow || total > dev_cap)
        return fail(ctx, SWB_ERR_CAPACITY, "%lld segments in this submit exceed max_segments (%d)", (long long)total,
                    ctx->cap_rows);
    ctx->rows_guess = std::max<int64_t>(256, 2 * total);
    if (n_rows) *n_rows = total;
    if (per_frame_counts)
        for (int f = 0; f < T; ++f) per_frame_counts[f] = ctx->h_segoff[f + 1] - ctx->h_segoff[f];
    if (rows && total > cap)
        return fail(ctx, SWB_ERR_CAPACITY, "%lld rows do not fit the caller's %lld", (long long)total, (long long)cap);
    if (rows && total > guess) {
        CU(ctx, cudaMemcpyAsync(rows + guess, d_rows + guess, (size_t)(total - guess) * sizeof(swb_segment),
                                cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
    }
    return SWB_OK;
}

int swb_collect(swb_ctx* ctx, swb_segment* rows, int64_t cap, int64_t* n_rows, int32_t* per_frame_counts) {
    const int rc = collect_begin_impl(ctx, rows, cap, nullptr, nullptr);
    return rc != SWB_OK ? rc : collect_end_impl(ctx, n_rows, per_frame_counts);
}

int swb_collect_all(swb_ctx* ctx, swb_segment* rows, int64_t cap, int64_t* n_rows, int32_t* per_frame_counts,
                    uint8_t* masks, void* labels) {
    const int rc = collect_begin_impl(ctx, rows, cap, masks, labels);
    return rc != SWB_OK ? rc : collect_end_impl(ctx, n_rows, per_frame_counts);
}

int swb_collect_begin(swb_ctx* ctx, swb_segment* rows, int64_t cap, uint8_t* masks, void* labels) {
    return collect_begin_impl(ctx, rows, cap, masks, labels);
}

int swb_collect_end(swb_ctx* ctx, int64_t* n_rows, int32_t* per_frame_counts) {
    return collect_end_impl(ctx, n_rows, per_frame_counts);
}

int swb_last_subchunk(swb_ctx* ctx, int32_t* frames) {
    if (!ctx || !frames) return fail(ctx, SWB_ERR_INVALID, "null argument");
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "no submit yet");
    *frames = ctx->last_ts;
    return SWB_OK;
}

int swb_set_option(swb_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return fail(ctx, SWB_ERR_INVALID, "null argument");
    if (!strcmp(name, "host_pipeline")) ctx->pipeline = value != 0;
    else if (!strcmp(name, "sub_batch_min_px")) {
        if (value <= 0) return fail(ctx, SWB_ERR_INVALID, "sub_batch_min_px must be positive");
        ctx->sub_min_px = value;
    } else if (!strcmp(name, "rpca_device_loop")) {
        if (value < -1 || value > 1) return fail(ctx, SWB_ERR_INVALID, "rpca_device_loop must be -1 (automatic), 0 or 1");
        ctx->rpca.device_loop = (int)value;
    } else if (!strcmp(name, "temporal_subchunk")) {
        if (value < 0 || value > 32768) return fail(ctx, SWB_ERR_INVALID, "temporal_subchunk must be in 0..32768");
        ctx->forced_ts = (int)value;
    } else return fail(ctx, SWB_ERR_INVALID, "unknown option '%s'", name);
    return SWB_OK;
}

static int copy_out(swb_ctx* ctx, const void* src, size_t elem, int32_t t0, int32_t n, void* dst, int32_t mem_kind) {
    const Geom& g = ctx->g;
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "no submit to read from");
    if (!src) return fail(ctx, SWB_ERR_STATE, "this output was not enabled in swb_config.out_flags");
    if (t0 < 0 || n < 0 || t0 + n > ctx->last_T) return fail(ctx, SWB_ERR_INVALID, "frame range [%d,%d) outside 0..%d", t0, t0 + n, ctx->last_T);
    if (n == 0) return SWB_OK;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const uint8_t* sp = static_cast<const uint8_t*>(src) + (size_t)t0 * g.h * g.mpitch * elem;
    CU(ctx, cudaMemcpy2DAsync(dst, (size_t)g.w * elem, sp, (size_t)g.mpitch * elem, (size_t)g.w * elem,
                              (size_t)n * g.h, mem_kind == SWB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                              ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SWB_OK;
}

int swb_get_masks(swb_ctx* ctx, int32_t t0, int32_t n, uint8_t* dst, int32_t mem_kind) {
    if (!ctx || !dst) return fail(ctx, SWB_ERR_INVALID, "null argument");
    return copy_out(ctx, ctx->mask, 1, t0, n, dst, mem_kind);
}

int swb_get_labels(swb_ctx* ctx, int32_t t0, int32_t n, void* dst, int32_t mem_kind) {
    if (!ctx || !dst) return fail(ctx, SWB_ERR_INVALID, "null argument");
    return copy_out(ctx, ctx->labels, (size_t)ctx->label_elem, t0, n, dst, mem_kind);
}

int swb_get_mask_bits(swb_ctx* ctx, int32_t t0, int32_t n, uint32_t* dst, int32_t mem_kind) {
    if (!ctx || !dst) return fail(ctx, SWB_ERR_INVALID, "null argument");
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "no submit to read from");
    if (t0 < 0 || n < 0 || t0 + n > ctx->last_T) return fail(ctx, SWB_ERR_INVALID, "bad frame range");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const Geom& g = ctx->g;
    CU(ctx, cudaMemcpy2DAsync(dst, (size_t)g.wpr * 4, ctx->fbits + (size_t)t0 * g.h * g.wpr4, (size_t)g.wpr4 * 4,
                              (size_t)g.wpr * 4, (size_t)n * g.h,
                              mem_kind == SWB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SWB_OK;
}

int swb_device_views(swb_ctx* ctx, uint8_t** mask, int64_t* mask_pitch, void** labels, int64_t* labels_pitch_elems,
                     swb_segment** rows, int32_t** per_frame_counts) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    if (mask) *mask = ctx->mask;
    if (mask_pitch) *mask_pitch = ctx->g.mpitch;
    if (labels) *labels = ctx->labels;
    if (labels_pitch_elems) *labels_pitch_elems = ctx->g.mpitch;
    if (rows) *rows = ctx->ccl.rows;
    if (per_frame_counts) *per_frame_counts = ctx->ccl.nseg;
    return SWB_OK;
}

int swb_enable_timing(swb_ctx* ctx, int32_t on) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    ctx->timing = on != 0;
    ctx->ev_valid = false;
    return SWB_OK;
}

int swb_get_timing(swb_ctx* ctx, const char** names, float* ms, int32_t cap, int32_t* n) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    if (!ctx->ev_valid) return fail(ctx, SWB_ERR_STATE, "timing not enabled for the last submit");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    CU(ctx, cudaEventSynchronize(ctx->ev[N_TIMERS]));
    const int k = std::min<int>(cap, N_TIMERS);
    for (int i = 0; i < k; ++i) {
        if (names) names[i] = TIMER_NAMES[i];
        if (ms) CU(ctx, cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
    }
    if (n) *n = k;
    return SWB_OK;
}

int64_t swb_launch_count(const swb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int swb_gather_crops(swb_ctx* ctx, int32_t crop, uint8_t* dst, int32_t* rects, int32_t mem_kind) {
    if (!ctx || !dst) return fail(ctx, SWB_ERR_INVALID, "null argument");
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "no submit to read from");
    if (ctx->fetching) return fail(ctx, SWB_ERR_STATE, "swb_gather_crops between swb_collect_begin and swb_collect_end");
    if (!ctx->last_frames_dev)
        return fail(ctx, SWB_ERR_STATE, "full frames are not resident on the device (host submit with a partial ROI); "
                                        "crop on the host instead");
    if (crop <= 0 || crop > 64) return fail(ctx, SWB_ERR_INVALID, "crop must be in 1..64");
    const swb_config& c = ctx->cfg;
    CU(ctx, cudaSetDevice(c.device));
    cudaStream_t s = ctx->stream;
    const int T = ctx->last_T;
    const bool u8 = c.label_mode == SWB_LABELS_U8;       // the merged table: regionprops of the uint8 image
    const int32_t* d_segoff = u8 ? ctx->u8.segoff : ctx->ccl.segoff;
    const swb_segment* d_rows = u8 ? ctx->u8.rows : ctx->ccl.rows;
    CU(ctx, cudaMemcpyAsync(ctx->h_segoff, d_segoff, ((size_t)T + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    const int total = std::min<int>(ctx->h_segoff[T], u8 ? ctx->u8.cap : ctx->cap_rows);
    if (total == 0) return SWB_OK;
    const size_t bytes = (size_t)total * crop * crop * c.channels;
    const size_t rbytes = (size_t)total * 4 * sizeof(int32_t);
    uint8_t* d = dst;
    int32_t* dr = rects;
    DevTmp tmp;
    if (mem_kind == SWB_MEM_HOST) {
        CU(ctx, tmp.alloc(&d, bytes));
        if (rects) CU(ctx, tmp.alloc(&dr, (size_t)total * 4));
    }
    cudaError_t e = launch_gather_crops_n(s, ctx->last_frames_dev, ctx->last_stride, ctx->last_pitch, c.channels,
                                          c.frame_h, c.frame_w, c.roi_x0, c.roi_y0, d_rows, total, crop, d, dr);
    ctx->launches += 1;
    if (e == cudaSuccess && mem_kind == SWB_MEM_HOST) {
        e = cudaMemcpyAsync(dst, d, bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && rects) e = cudaMemcpyAsync(rects, dr, rbytes, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(ctx, SWB_ERR_CUDA, "gather_crops: %s", cudaGetErrorString(e));
    return SWB_OK;
}

// ---- single-stage entry points -------------------------------------------------------
#define STAGE_PROLOGUE()                                   \
    do {                                                   \
        int rc__ = select_device(nullptr, device);         \
        if (rc__ != SWB_OK) return rc__;                   \
    } while (0)

int swb_stage_gray(int32_t device, const uint8_t* bgr, int32_t h, int32_t w, uint8_t* out) {
    if (!bgr || !out || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    STAGE_PROLOGUE();
    DevTmp t;
    uint8_t *d_in, *d_out;
    const size_t n = (size_t)h * w;
    CU(nullptr, t.alloc(&d_in, n * 3));
    CU(nullptr, t.alloc(&d_out, n));
    CU(nullptr, cudaMemcpy(d_in, bgr, n * 3, cudaMemcpyHostToDevice));
    CU(nullptr, launch_stage_gray(0, d_in, h, w, d_out));
    CU(nullptr, cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

int swb_stage_median(int32_t device, const uint8_t* stack, int32_t n, int32_t h, int32_t w, uint8_t* out) {
    if (!stack || !out || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    if (n < 1 || n > 9 || !(n & 1)) return fail(nullptr, SWB_ERR_INVALID, "n must be odd and in 1..9");
    STAGE_PROLOGUE();
    DevTmp t;
    uint8_t *d_in, *d_out;
    const size_t npx = (size_t)h * w;
    CU(nullptr, t.alloc(&d_in, npx * n));
    CU(nullptr, t.alloc(&d_out, npx));
    CU(nullptr, cudaMemcpy(d_in, stack, npx * n, cudaMemcpyHostToDevice));
    CU(nullptr, launch_stage_median(0, d_in, n, h, w, d_out));
    CU(nullptr, cudaMemcpy(out, d_out, npx, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

int swb_stage_absdiff(int32_t device, const uint8_t* a, const uint8_t* b, int32_t h, int32_t w, uint8_t* out) {
    if (!a || !b || !out || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    STAGE_PROLOGUE();
    DevTmp t;
    uint8_t *d_a, *d_b, *d_out;
    const size_t n = (size_t)h * w;
    CU(nullptr, t.alloc(&d_a, n));
    CU(nullptr, t.alloc(&d_b, n));
    CU(nullptr, t.alloc(&d_out, n));
    CU(nullptr, cudaMemcpy(d_a, a, n, cudaMemcpyHostToDevice));
    CU(nullptr, cudaMemcpy(d_b, b, n, cudaMemcpyHostToDevice));
    CU(nullptr, launch_stage_absdiff(0, d_a, d_b, (long long)n, d_out));
    CU(nullptr, cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

int swb_stage_thresh_to_zero(int32_t device, const uint8_t* in, int32_t h, int32_t w, int32_t thresh, uint8_t* out) {
    if (!in || !out || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    STAGE_PROLOGUE();
    DevTmp t;
    uint8_t *d_in, *d_out;
    const size_t n = (size_t)h * w;
    CU(nullptr, t.alloc(&d_in, n));
    CU(nullptr, t.alloc(&d_out, n));
    CU(nullptr, cudaMemcpy(d_in, in, n, cudaMemcpyHostToDevice));
    CU(nullptr, launch_stage_thresh(0, d_in, (long long)n, thresh, d_out));
    CU(nullptr, cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

int swb_stage_grey_morph(int32_t device, const uint8_t* in, int32_t h, int32_t w, int32_t se_h, int32_t se_w,
                         int32_t closing, uint8_t* out) {
    if (!in || !out || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    if (se_h < 1 || se_w < 1 || !(se_h & 1) || !(se_w & 1) || se_h > 31 || se_w > 31)
        return fail(nullptr, SWB_ERR_INVALID, "structuring element must be odd-sized, 1..31 (got %dx%d)", se_h, se_w);
    STAGE_PROLOGUE();
    DevTmp t;
    uint8_t *d_a, *d_b;
    const size_t n = (size_t)h * w;
    CU(nullptr, t.alloc(&d_a, n));
    CU(nullptr, t.alloc(&d_b, n));
    CU(nullptr, cudaMemcpy(d_a, in, n, cudaMemcpyHostToDevice));
    // opening = erosion (min) then dilation (max); closing is the dual
    CU(nullptr, launch_stage_minmax(0, d_a, h, w, se_h, se_w, closing ? 1 : 0, d_b));
    CU(nullptr, launch_stage_minmax(0, d_b, h, w, se_h, se_w, closing ? 0 : 1, d_a));
    CU(nullptr, cudaMemcpy(out, d_a, n, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

int swb_stage_cc_label(int32_t device, const uint8_t* in, int32_t h, int32_t w, int32_t* out_i32, uint8_t* out_u8,
                       int32_t* n_labels) {
    if (!in || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    STAGE_PROLOGUE();
    Geom g;
    int X0a;
    make_geom(0, 0, w, h, g, X0a);
    DevTmp t;
    uint8_t* d_in;
    uint32_t* d_bits;
    int32_t* d_lab;
    const size_t n = (size_t)h * w;
    CU(nullptr, t.alloc(&d_in, n));
    CU(nullptr, t.alloc(&d_bits, (size_t)g.h * g.wpr4));
    CU(nullptr, t.alloc(&d_lab, (size_t)g.h * g.mpitch));
    CclBuffers b{};
    const int cap = g.BH * g.BW;   // every 2x2 block its own component at most
    {
        const size_t nblk = (size_t)g.BH * g.BW;
        CU(nullptr, t.alloc(&b.parent, nblk));
        CU(nullptr, t.alloc(&b.rowcount, (size_t)2 * g.BH));
        CU(nullptr, t.alloc(&b.nseg, 1));
        CU(nullptr, t.alloc(&b.segoff, 2));
        CU(nullptr, t.alloc(&b.rows, (size_t)cap));
        b.cap_rows = cap;
        b.cap_parts = 2 * cap + 4096;
        CU(nullptr, t.alloc(&b.parts, (size_t)b.cap_parts));
        CU(nullptr, t.alloc(&b.pcount, 3));
        b.overflow = b.pcount + 2;
        CU(nullptr, t.alloc(&b.big_tiles, (size_t)(g.BH + 7) / 8));
        CU(nullptr, t.alloc(&b.rootlist, (size_t)cap));
    }
    CU(nullptr, cudaMemcpy(d_in, in, n, cudaMemcpyHostToDevice));
    CU(nullptr, launch_pack_bits(0, d_in, h, w, d_bits, g.wpr4));
    CU(nullptr, launch_ccl(0, d_bits, 1, g, b, d_lab, 4, nullptr, nullptr, 0));
    std::vector<int32_t> lab((size_t)h * w);
    CU(nullptr, cudaMemcpy2D(lab.data(), (size_t)w * 4, d_lab, (size_t)g.mpitch * 4, (size_t)w * 4, (size_t)h,
                             cudaMemcpyDeviceToHost));
    int32_t nseg = 0;
    CU(nullptr, cudaMemcpy(&nseg, b.nseg, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (n_labels) *n_labels = nseg;
    if (out_i32) memcpy(out_i32, lab.data(), lab.size() * sizeof(int32_t));
    if (out_u8)
        for (size_t i = 0; i < lab.size(); ++i) out_u8[i] = (uint8_t)lab[i];   // labels.astype(np.uint8)
    return SWB_OK;
}

int swb_stage_regionprops(int32_t device, const void* labels, int32_t elem_size, int32_t h, int32_t w,
                          swb_segment* rows, int32_t cap, int32_t* n_rows) {
    if (!labels || !rows || h <= 0 || w <= 0 || (elem_size != 1 && elem_size != 4))
        return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    STAGE_PROLOGUE();
    const size_t n = (size_t)h * w;
    int64_t maxlab = 0;
    if (elem_size == 1) {
        const uint8_t* p = static_cast<const uint8_t*>(labels);
        for (size_t i = 0; i < n; ++i) maxlab = std::max<int64_t>(maxlab, p[i]);
    } else {
        const int32_t* p = static_cast<const int32_t*>(labels);
        for (size_t i = 0; i < n; ++i) maxlab = std::max<int64_t>(maxlab, p[i]);
    }
    if (n_rows) *n_rows = 0;
    if (maxlab <= 0) return SWB_OK;
    DevTmp t;
    uint8_t* d_lab;
    swb_segment* d_acc;
    CU(nullptr, t.alloc(&d_lab, n * elem_size));
    CU(nullptr, t.alloc(&d_acc, (size_t)maxlab));
    CU(nullptr, cudaMemcpy(d_lab, labels, n * elem_size, cudaMemcpyHostToDevice));
    CU(nullptr, launch_stage_props(0, d_lab, elem_size, h, w, d_acc, (int)maxlab));
    std::vector<swb_segment> acc((size_t)maxlab);
    CU(nullptr, cudaMemcpy(acc.data(), d_acc, acc.size() * sizeof(swb_segment), cudaMemcpyDeviceToHost));
    int32_t k = 0;
    for (const swb_segment& s : acc) {
        if (s.area <= 0) continue;   // label value absent: find_objects yields None (skipped)
        if (k >= cap) return fail(nullptr, SWB_ERR_CAPACITY, "more than %d regions", cap);
        rows[k++] = s;
    }
    if (n_rows) *n_rows = k;
    return SWB_OK;
}

int swb_get_rpca(swb_ctx* ctx, int32_t t0, int32_t n, uint8_t* dst, int32_t mem_kind) {
    if (!ctx || !dst) return fail(ctx, SWB_ERR_INVALID, "null argument");
    if (ctx->cfg.bg_model != SWB_BG_RPCA) return fail(ctx, SWB_ERR_STATE, "the context does not use bg_model SWB_BG_RPCA");
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "no submit to read from");
    if (t0 < 0 || n < 0 || t0 + n > ctx->last_T) return fail(ctx, SWB_ERR_INVALID, "bad frame range");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t P = (size_t)ctx->g.h * ctx->g.w;
    // the stack is newest first: frame t is image last_T - 1 - t
    for (int t = t0; t < t0 + n; ++t)
        CU(ctx, cudaMemcpyAsync(dst + (size_t)(t - t0) * P, ctx->rp_sparse + (size_t)(ctx->last_T - 1 - t) * P, P,
                                mem_kind == SWB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SWB_OK;
}

int swb_rpca_stats(swb_ctx* ctx, int32_t* iterations, int32_t* jacobi_sweeps, int32_t* device_loop) {
    if (!ctx) return fail(nullptr, SWB_ERR_INVALID, "null context");
    if (ctx->cfg.bg_model != SWB_BG_RPCA) return fail(ctx, SWB_ERR_STATE, "the context does not use bg_model SWB_BG_RPCA");
    if (!ctx->pending) return fail(ctx, SWB_ERR_STATE, "no submit yet");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    int it = 0, sw = 0, mode = 0;
    CU(ctx, rpca_read_state(ctx->stream, ctx->rpca, &it, &sw, &mode));
    if (iterations) *iterations = it;
    if (jacobi_sweeps) *jacobi_sweeps = sw;
    if (device_loop) *device_loop = mode;
    return SWB_OK;
}

int swb_stage_rpca(int32_t device, const uint8_t* frames, int32_t n, int32_t h, int32_t w, uint8_t* out, int32_t* iters) {
    if (!frames || !out || n < 1 || n > 32 || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument (1 <= n <= 32)");
    STAGE_PROLOGUE();
    const long long P = (long long)h * w;
    DevTmp t;
    uint8_t *d_in, *d_out;
    CU(nullptr, t.alloc(&d_in, (size_t)P * n));
    CU(nullptr, t.alloc(&d_out, (size_t)P * n));
    RpcaWork wk;
    cudaError_t e = rpca_alloc(wk, P, n);
    if (e != cudaSuccess) { rpca_free(wk); return fail(nullptr, SWB_ERR_CUDA, "rpca_alloc: %s", cudaGetErrorString(e)); }
    int it = 0;
    e = cudaMemcpy(d_in, frames, (size_t)P * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = rpca_run(0, d_in, n, P, wk, d_out, &it, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)P * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = rpca_read_state(0, wk, &it, nullptr, nullptr);
    rpca_free(wk);
    if (e != cudaSuccess) return fail(nullptr, SWB_ERR_CUDA, "swb_stage_rpca: %s", cudaGetErrorString(e));
    if (iters) *iters = it;
    return SWB_OK;
}

int swb_stage_bilateral(int32_t device, const uint8_t* in, int32_t h, int32_t w, int32_t d, double sigma_color,
                        double sigma_space, uint8_t* out) {
    if (!in || !out || h <= 0 || w <= 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    if (d < 1 || d > 7 || !(sigma_color > 0) || !(sigma_space > 0))
        return fail(nullptr, SWB_ERR_INVALID, "bilateral: 1 <= d <= 7 and positive sigmas (the reference uses 7, 15, 1)");
    STAGE_PROLOGUE();
    DevTmp t;
    uint8_t *d_in, *d_out;
    BilateralLut* d_lut;
    const size_t n = (size_t)h * w;
    CU(nullptr, t.alloc(&d_in, n));
    CU(nullptr, t.alloc(&d_out, n));
    CU(nullptr, t.alloc(&d_lut, 1));
    BilateralLut lut;
    bilateral_lut(d, sigma_color, sigma_space, lut);
    CU(nullptr, cudaMemcpy(d_lut, &lut, sizeof(lut), cudaMemcpyHostToDevice));
    CU(nullptr, cudaMemcpy(d_in, in, n, cudaMemcpyHostToDevice));
    CU(nullptr, launch_bilateral(0, d_in, 1, h, w, d_lut, 0, d_out, 0, nullptr, 0, lut.radius));
    CU(nullptr, cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

int swb_host_gather_tiles(const uint64_t* src, int64_t pitch, int32_t rows, int32_t row_bytes, int64_t n, uint8_t* dst) {
    if ((n > 0 && (!src || !dst)) || rows <= 0 || row_bytes <= 0 || n < 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    const size_t tile = (size_t)rows * row_bytes;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(static_cast<uintptr_t>(src[i]));
        uint8_t* d = dst + (size_t)i * tile;
        for (int r = 0; r < rows; ++r) memcpy(d + (size_t)r * row_bytes, s + (size_t)r * pitch, (size_t)row_bytes);
    }
    return SWB_OK;
}

int swb_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    *ptr = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        return fail(nullptr, SWB_ERR_CUDA, "no CUDA device available; libswb200 has no CPU path");
    cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, SWB_ERR_CUDA, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return SWB_OK;
}

int swb_host_free(void* ptr) {
    if (!ptr) return SWB_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return fail(nullptr, SWB_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
    return SWB_OK;
}

int swb_synth_frames(int32_t device, uint8_t* dst, int32_t mem_kind, uint32_t seed, uint32_t video, int32_t t0,
                     int32_t n, int32_t h, int32_t w, int32_t n_birds) {
    if (!dst || n <= 0 || h <= 0 || w <= 0 || n_birds < 0 || n > 65535) return fail(nullptr, SWB_ERR_INVALID, "bad argument");
    STAGE_PROLOGUE();
    const size_t bytes = (size_t)n * h * w * 3;
    if (mem_kind == SWB_MEM_DEVICE) {
        CU(nullptr, launch_synth(0, dst, seed, video, t0, n, h, w, n_birds));
        CU(nullptr, cudaDeviceSynchronize());
    } else {
        DevTmp t;
        uint8_t* d;
        CU(nullptr, t.alloc(&d, bytes));
        CU(nullptr, launch_synth(0, d, seed, video, t0, n, h, w, n_birds));
        CU(nullptr, cudaMemcpy(dst, d, bytes, cudaMemcpyDeviceToHost));
    }
    return SWB_OK;
}

}  // extern "C"
