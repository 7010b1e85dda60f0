// K2 — binary morphology on the bit-packed mask + uint8 mask materialisation.
//
// Replaces grayscale_opening (image_filtering.py:319-322; scipy grey_opening,
// flat 3x3 at data_structures.py:202) and its dual closing, applied to
// thresh_to_zero output.  Min/max filters commute with the monotone map
// v -> [v > 0] and THRESH_TOZERO output is 0 or > thresh, so binary
// erosion/dilation of the 1-bit mask equals (grey_opening(x) > 0) bit for bit.
// Border rule: scipy's mode='reflect' only duplicates pixels that are already
// inside the window, i.e. out-of-image pixels are ignored: they read as 1 for
// an erosion and as 0 for a dilation.
//
// Streaming design: a warp owns a slab of 32 consecutive bit words (lane = word;
// lanes 1..30 produce output, lanes 0 and 31 are the horizontal halo) and marches
// down a strip of 32 (64 for the deepest chain) rows.  Every erosion / dilation stage keeps the last 2R+1
// horizontally processed rows of its input in registers (a delay line), so a row
// of 1024 pixels goes through the whole open/close chain with a handful of
// funnel shifts, LOP3s and two shuffles per stage, and nothing but the raw bit
// words is ever re-read.  The kernel also realigns K1's "raw" bit columns
// (origin X0a) to ROI-local columns (origin roi_x0), writes the final bit words
// (rows padded to wpr4 words with zeros) and expands them to the {0,255} uint8
// mask with 16-byte stores that are contiguous across the warp.
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through the runtime)

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "swb_internal.cuh"

namespace swb {

namespace {

// ---- TMA staging (STAGED variant): one tensor-map box load brings every raw bit word a warp's strip needs.
// The box starts at a word column that is a multiple of four (the innermost start of a tiled tensor load has to be
// 16-byte aligned: an unaligned start is an illegal instruction, profiles/probes/tma_probe.cu) and at a row >= 0.
__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// wait for phase `parity` of an mbarrier (bounded: a pipeline bug traps instead of hanging the GPU)
__device__ __forceinline__ void k2_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = k2_smem_u32(bar);
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1 << 20)) __trap();
    }
}
constexpr int BOXW = 36;   // words per staged row: up to 3 words of alignment slack + the slab's 32 words + the edge word

typedef CUresult (*K2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// raw bits as a tensor (words per row, rows, frames) of 32-bit elements; box = (BOXW, box_rows, 1)
bool encode_raw_bits_tensor(CUtensorMap* tmap, const uint32_t* base, int wpr_raw, int h, int T, int box_rows) {
    static K2EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<K2EncodeFn>(fn);
    }();
    if (!encode || box_rows > 256 || (wpr_raw & 3) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)wpr_raw, (cuuint64_t)h, (cuuint64_t)T};
    const cuuint64_t strides[2] = {(cuuint64_t)wpr_raw * 4, (cuuint64_t)wpr_raw * 4 * (cuuint64_t)h};
    const cuuint32_t box[3] = {(cuuint32_t)BOXW, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int SLAB = 30;   // output words per warp
constexpr int WPB = 4;     // warps per CTA

// operation chains: 0 none, 1 open (E,D), 2 close (D,E), 3 open + close (E,D,D,E)
__host__ __device__ constexpr int n_ops(int pat) { return pat == 0 ? 0 : (pat == 3 ? 4 : 2); }
// output rows per warp: longer strips when the vertical halo is deep
__host__ __device__ constexpr int strip_rows(int r, int pat) { return n_ops(pat) * r >= 8 ? 64 : 32; }
__host__ __device__ constexpr bool op_is_erode(int pat, int s) {
    return pat == 1 ? (s == 0) : (pat == 2 ? (s == 1) : (s == 0 || s == 3));
}

__device__ __forceinline__ uint32_t colmask(int j, int w, int wpr) {
    if (j < 0 || j >= wpr) return 0u;
    const int rem = w - 32 * j;
    return rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
}

template <int R>
__device__ __forceinline__ uint32_t hop(uint32_t l, uint32_t c, uint32_t r, bool erode) {
    uint32_t acc = c;   // combine pixel x with x-R..x+R
#pragma unroll
    for (int s = 1; s <= R; ++s) {
        const uint32_t right = __funnelshift_r(c, r, s);   // bit i = pixel i+s
        const uint32_t left = __funnelshift_l(l, c, s);    // bit i = pixel i-s
        acc = erode ? (acc & right & left) : (acc | right | left);
    }
    return acc;
}

// One warp, one strip of rows.  INTERIOR: every row the strip touches (its own rows and the
// vertical halo of every stage) lies inside the image, so no row needs the border fill and
// the per-row validity tests disappear; DX0: the raw bit columns are already ROI-aligned
// (full-frame ROI), so there is no realignment shift and no extra edge word.
// The row loop is unrolled by the window height 2R+1: a stage's window is "the last 2R+1
// horizontally filtered rows", AND / OR do not care about their order, so row r simply
// overwrites slot r mod (2R+1) — a static register after unrolling, no rotation moves.
// SEG: lanes per frame.  32: the warp works on one frame.  16 (rows of at most 14 words: narrow ROIs): the two
// half-warps work on the same rows of two consecutive frames, each with its own halo lanes, so a narrow
// ROI keeps 28 of 32 lanes busy instead of 12; `live` is false for the half-warp past the last frame.
// STAGED: the raw rows [row0, row0 + staged rows) of the strip, BOXW words each starting at word column col0, were
// brought to shared memory (`srow`) by one TMA box load (rows / columns past the image arrived as zeros).
template <int R, int PAT, bool INTERIOR, bool DX0, int SEG, bool STAGED = false>
__device__ __forceinline__ void morph_strip(const uint32_t* __restrict__ raw_f, const Geom& g, uint32_t* __restrict__ fbits,
                                            uint8_t* __restrict__ mask, int f, bool live, int slab, int lane, int y0,
                                            const uint32_t* srow = nullptr, int row0 = 0, int col0 = 0, int sr_rt = 0,
                                            const CUtensorMap* tmap = nullptr, uint64_t* bar = nullptr) {
    constexpr int SLABW = SEG - 2;                           // output words per segment
    const int hl = lane & (SEG - 1);                         // lane within the segment
    const int seg0 = lane & ~(SEG - 1);
    constexpr int NOPS = n_ops(PAT);
    constexpr int NW = NOPS > 0 ? NOPS : 1;
    constexpr int HR = NOPS * R;   // rows of vertical halo
    constexpr int SR = strip_rows(R, PAT);
    constexpr int W = 2 * R + 1;
    const int j = slab * SLABW + hl - 1;                     // this lane's word
    const uint32_t Vcol = colmask(j, g.w, g.wpr);
    const bool in_raw = (j >= 0 && j < g.wpr_raw);
    const bool in_raw_next = (j + 1 >= 0 && j + 1 < g.wpr_raw);
    constexpr uint32_t fill0 = (NOPS > 0 && op_is_erode(PAT, 0)) ? 0xFFFFFFFFu : 0u;
    // what a stage's output reads as outside the image columns: the identity of the next stage
    uint32_t fillc[NW];
#pragma unroll
    for (int s = 0; s < NW; ++s)
        fillc[s] = (s + 1 < NOPS && op_is_erode(PAT, s + 1)) ? ~Vcol : 0u;

    uint32_t win[NW][W];
#pragma unroll
    for (int s = 0; s < NW; ++s)
#pragma unroll
        for (int i = 0; i < W; ++i) win[s][i] = 0u;

    const int y_end = min(y0 + (sr_rt > 0 ? sr_rt : SR), g.h);   // sr_rt: strip height chosen at launch (narrow ROIs)
    const int y_stop = y_end + HR;
    // mask output: lane -> two 16-byte chunks of the slab's row
    const bool st_bits = live && hl >= 1 && hl <= SLABW && j < g.wpr4;
    bool st_mask[2];
    int src_lane[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int wi = (half * SEG + hl) >> 1;               // word within the slab
        st_mask[half] = mask != nullptr && live && wi < SLABW && slab * SLABW + wi < g.wpr;
        src_lane[half] = seg0 + (wi < SLABW ? wi + 1 : SEG - 1);
    }
    // raw words of row y (this lane's word; lane 31 also fetches the word after it), 0 outside the image
    auto fetch = [&](int y, uint32_t& lo, uint32_t& edge) {
        const uint32_t lo_prev = lo;                         // the row fetched by the previous call
        (void)lo_prev;
        lo = 0u;
        edge = 0u;
        if constexpr (STAGED) {
            // The strip's SR + 2 HR raw rows pass through a staging buffer of HALF as many rows in two fills: when
            // the (warp-uniform) row counter reaches the second half, every row of the first half has been read into
            // registers, so the same buffer takes rows [row0 + HALF, row0 + 2 HALF).  Half the shared memory per
            // warp = eight instead of four resident CTAs per SM; the second fill's latency is paid once per strip.
            constexpr int HALF = (SR + 2 * HR) / 2;
            const int yrel = y - row0;                               // rows above / below the image (and the one-row
            if (yrel == HALF && y < g.h) {                           // overshoot of the prefetch) read as zeros
                asm volatile("" ::"r"(__shfl_sync(0xFFFFFFFFu, lo_prev, 0)) : "memory");   // the first half's loads have landed
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2_smem_u32(bar)),
                                 "r"((uint32_t)(HALF * BOXW * 4))
                                 : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                            k2_smem_u32(srow)),
                        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col0), "r"(row0 + HALF), "r"(f), "r"(k2_smem_u32(bar))
                        : "memory");
                }
                k2_mbar_wait(bar, 1u);
            }
            if ((unsigned)yrel < (unsigned)(2 * HALF) && j >= 0 && y < g.h) {
                const uint32_t* rowp = srow + (yrel >= HALF ? yrel - HALF : yrel) * BOXW + (j - col0);
                lo = rowp[0];
                if (!DX0 && hl == SEG - 1) edge = rowp[1];
            }
        } else if ((INTERIOR || (unsigned)y < (unsigned)g.h) && live) {
            const uint32_t* rowp = raw_f + (long long)y * g.wpr_raw;
            if (in_raw) lo = __ldg(rowp + j);
            if (!DX0 && hl == SEG - 1 && in_raw_next) edge = __ldg(rowp + j + 1);
        }
    };
    uint32_t lo_n, edge_n;
    fetch(y0 - HR, lo_n, edge_n);
    // output pointers advance by one row per output row (the strip's rows come out in order, starting at y0):
    // recomputing (f * h + y) * pitch + j per row was a fifth of the kernel's instructions (ncu source view)
    const long long orow0 = (long long)f * g.h + y0;
    uint32_t* fb_out = fbits + orow0 * g.wpr4 + j;
    uint8_t* m_out = mask + orow0 * g.mpitch + (long long)slab * (SLABW * 32);
    for (int base = y0 - HR; base < y_stop; base += W) {
#pragma unroll
        for (int u = 0; u < W; ++u) {
            const int yin = base + u;
            if (yin >= y_stop) break;                        // warp-uniform
            // ---- raw row (prefetched one row ahead), realigned to ROI columns, border filled for the first op
            const uint32_t lo = lo_n, edge = edge_n;
            if (INTERIOR || yin + 1 < y_stop) fetch(yin + 1, lo_n, edge_n);
            uint32_t v;
            if constexpr (DX0) {
                v = lo;
            } else {
                uint32_t hi = __shfl_down_sync(0xFFFFFFFFu, lo, 1);
                if (hl == SEG - 1) hi = edge;
                v = __funnelshift_r(lo, hi, g.dx);
            }
            uint32_t cur;
            if constexpr (INTERIOR) {
                cur = (v & Vcol) | (~Vcol & fill0);
            } else {
                const uint32_t V0 = ((unsigned)yin < (unsigned)g.h) ? Vcol : 0u;
                cur = (v & V0) | (~V0 & fill0);
            }
            // ---- erosion / dilation chain: stage s emits row yin - (s + 1) * R
#pragma unroll
            for (int s = 0; s < NOPS; ++s) {
                const bool erode = op_is_erode(PAT, s);
                // lanes 0 / 31 have no neighbour on one side: whatever they read only spoils the far R
                // bits of the halo word per stage, which no output word ever looks at
                const uint32_t l = __shfl_up_sync(0xFFFFFFFFu, cur, 1);
                const uint32_t r = __shfl_down_sync(0xFFFFFFFFu, cur, 1);
                win[s][u] = hop<R>(l, cur, r, erode);
                uint32_t acc = win[s][0];
#pragma unroll
                for (int i = 1; i < W; ++i) acc = erode ? (acc & win[s][i]) : (acc | win[s][i]);
                if constexpr (INTERIOR) {
                    cur = (acc & Vcol) | fillc[s];
                } else {
                    const int ys = yin - (s + 1) * R;
                    const bool rv = (unsigned)ys < (unsigned)g.h;
                    const uint32_t fill_next = (s + 1 < NOPS && op_is_erode(PAT, s + 1)) ? 0xFFFFFFFFu : 0u;
                    cur = rv ? ((acc & Vcol) | fillc[s]) : fill_next;
                }
            }
            // ---- outputs for row yin - HR
            const int yout = yin - HR;
            if (yout < y0) continue;                             // warp-uniform (pipeline warm-up)
            if (st_bits) *fb_out = cur;
            fb_out += g.wpr4;
            if (mask != nullptr) {
                uint8_t* mrow = m_out;
                m_out += g.mpitch;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int c = half * SEG + hl;               // 16-byte chunk of the slab's row
                    const uint32_t word = __shfl_sync(0xFFFFFFFFu, cur, src_lane[half]);
                    if (st_mask[half]) {
                        const uint32_t b16 = (word >> ((c & 1) * 16)) & 0xFFFFu;
                        uint4 o;
                        o.x = expand4(b16 & 0xFu);
                        o.y = expand4((b16 >> 4) & 0xFu);
                        o.z = expand4((b16 >> 8) & 0xFu);
                        o.w = expand4(b16 >> 12);
                        __stcs(reinterpret_cast<uint4*>(mrow + c * 16), o);
                    }
                }
            }
        }
    }
}

template <int R, int PAT, int SEG>
__global__ void __launch_bounds__(32 * WPB)
k_morph_mask(const uint32_t* __restrict__ raw_bits, Geom g, int T, uint32_t* __restrict__ fbits,
             uint8_t* __restrict__ mask) {
    constexpr int HR = n_ops(PAT) * R;
    constexpr int SR = strip_rows(R, PAT);
    wait_for_previous_kernel();                              // launched as a dependent of K1 (raw bits)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slab = blockIdx.x;
    const int f = blockIdx.z * (32 / SEG) + lane / SEG;      // SEG = 16: one frame per half-warp
    const bool live = f < T;
    const int y0 = (blockIdx.y * WPB + warp) * SR;
    if (y0 >= g.h) return;                                   // warp-uniform
    const uint32_t* raw_f = raw_bits + (long long)f * g.h * g.wpr_raw;
    const bool interior = (y0 - HR >= 0) && (y0 + SR + HR <= g.h);   // warp-uniform
    if (g.dx == 0) {
        if (interior) morph_strip<R, PAT, true, true, SEG>(raw_f, g, fbits, mask, f, live, slab, lane, y0);
        else morph_strip<R, PAT, false, true, SEG>(raw_f, g, fbits, mask, f, live, slab, lane, y0);
    } else {
        if (interior) morph_strip<R, PAT, true, false, SEG>(raw_f, g, fbits, mask, f, live, slab, lane, y0);
        else morph_strip<R, PAT, false, false, SEG>(raw_f, g, fbits, mask, f, live, slab, lane, y0);
    }
}

// Narrow ROIs (rows of at most 14 words): a warp works on one strip of TWO consecutive frames (SEG = 16), warps are
// numbered linearly over (frame pair, strip) so that no CTA carries idle warps, and the strip height `sr` is chosen
// at launch (narrow_strip_rows).
template <int R, int PAT>
__global__ void __launch_bounds__(32 * WPB)
k_morph_mask_narrow(const uint32_t* __restrict__ raw_bits, Geom g, int T, uint32_t* __restrict__ fbits,
                    uint8_t* __restrict__ mask, int sr, int nstrips) {
    constexpr int HR = n_ops(PAT) * R;
    wait_for_previous_kernel();                              // launched as a dependent of K1 (raw bits)
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * WPB + (threadIdx.x >> 5);
    const int pair = wid / nstrips;
    const int strip = wid - pair * nstrips;
    if (2 * pair >= T) return;                               // warp-uniform
    const int f = 2 * pair + (lane >> 4);                    // one frame per half-warp
    const bool live = f < T;
    const int y0 = strip * sr;
    const uint32_t* raw_f = raw_bits + (long long)f * g.h * g.wpr_raw;
    const bool interior = (y0 - HR >= 0) && (y0 + sr + HR <= g.h);   // warp-uniform
    if (g.dx == 0) {
        if (interior) morph_strip<R, PAT, true, true, 16>(raw_f, g, fbits, mask, f, live, 0, lane, y0, nullptr, 0, 0, sr);
        else morph_strip<R, PAT, false, true, 16>(raw_f, g, fbits, mask, f, live, 0, lane, y0, nullptr, 0, 0, sr);
    } else {
        if (interior) morph_strip<R, PAT, true, false, 16>(raw_f, g, fbits, mask, f, live, 0, lane, y0, nullptr, 0, 0, sr);
        else morph_strip<R, PAT, false, false, 16>(raw_f, g, fbits, mask, f, live, 0, lane, y0, nullptr, 0, 0, sr);
    }
}

// Strip height for the narrow kernel: balanced strips of at most 24 rows.  Measured on a B200 at 320x240 /
// T = 2048 (3x3 open): 12, 16 or 24 rows 0.0491 ms, 32 rows 0.0509, 48 rows (one exact wave of warps) 0.0508,
// 60 rows 0.0525 — the kernel is bound by the latency of its stores, not by how its warps fill the machine,
// so the shorter strips (more warps in flight per frame) win slightly and help the 21-frame queue batches most.
int narrow_strip_rows(int h) {
    static const int forced = [] { const char* e = getenv("SWB_K2_NARROW_SR"); return e ? atoi(e) : 0; }();
    if (forced > 0) return std::min(forced, std::max(h, 1));
    const int k = (h + 23) / 24;
    return std::max(1, (h + k - 1) / std::max(k, 1));
}

// The same strips with the raw rows staged by TMA (frames wider than one slab whose raw rows are 16-byte multiples):
// the loads of a strip are all in flight at once instead of one row ahead of the arithmetic (the streaming kernel's
// top stall was that load: 6.5 cycles of long-scoreboard stall per issued instruction at 4K).
template <int R, int PAT>
__global__ void __launch_bounds__(32 * WPB)
k_morph_mask_staged(const __grid_constant__ CUtensorMap tmap, Geom g, int T, uint32_t* __restrict__ fbits,
                    uint8_t* __restrict__ mask) {
    constexpr int HR = n_ops(PAT) * R;
    constexpr int SR = strip_rows(R, PAT);
    constexpr int HALF = (SR + 2 * HR) / 2;                       // rows per fill of the staging buffer (two fills per strip)
    static_assert((SR + 2 * HR) % 2 == 0, "two equal fills");
    constexpr int REGION = (HALF * BOXW * 4 + 127) / 128 * 128;   // per-warp staging area, 128-byte aligned for TMA
    extern __shared__ __align__(128) uint8_t k2_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t* srow = reinterpret_cast<uint32_t*>(k2_smem + warp * REGION);
    uint64_t* bar = reinterpret_cast<uint64_t*>(k2_smem + WPB * REGION) + warp;
    const int slab = blockIdx.x;
    const int f = blockIdx.z;
    const int y0 = (blockIdx.y * WPB + warp) * SR;
    if (y0 >= g.h) return;                                   // warp-uniform
    const int row0 = max(y0 - HR, 0);                        // box start: a row inside the image ...
    const int col0 = max(slab * SLAB - 1, 0) & ~3;           // ... and a 16-byte aligned word column
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    wait_for_previous_kernel();                              // launched as a dependent of K1 (raw bits)
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2_smem_u32(bar)),
                     "r"((uint32_t)(HALF * BOXW * 4))
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                k2_smem_u32(srow)),
            "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(col0), "r"(row0), "r"(f), "r"(k2_smem_u32(bar))
            : "memory");
    }
    k2_mbar_wait(bar, 0u);
    const bool interior = (y0 - HR >= 0) && (y0 + SR + HR <= g.h);   // warp-uniform
    if (g.dx == 0) {
        if (interior) morph_strip<R, PAT, true, true, 32, true>(nullptr, g, fbits, mask, f, true, slab, lane, y0, srow, row0, col0, 0, &tmap, bar);
        else morph_strip<R, PAT, false, true, 32, true>(nullptr, g, fbits, mask, f, true, slab, lane, y0, srow, row0, col0, 0, &tmap, bar);
    } else {
        if (interior) morph_strip<R, PAT, true, false, 32, true>(nullptr, g, fbits, mask, f, true, slab, lane, y0, srow, row0, col0, 0, &tmap, bar);
        else morph_strip<R, PAT, false, false, 32, true>(nullptr, g, fbits, mask, f, true, slab, lane, y0, srow, row0, col0, 0, &tmap, bar);
    }
}

bool k2_tma_enabled() {
    static const bool on = [] { const char* e = getenv("SWB_K2_TMA"); return !(e && e[0] == '0'); }();
    return on;
}

template <int R, int PAT>
cudaError_t launch_pat(cudaStream_t s, const uint32_t* raw_bits, int T, const Geom& g, uint32_t* fbits,
                       uint8_t* mask) {
    constexpr int SR = strip_rows(R, PAT);
    const int nstrips = (g.h + SR - 1) / SR;
    if (g.wpr4 <= 14 && g.wpr_raw <= 15) {                   // narrow ROI: two frames per warp
        const int sr = narrow_strip_rows(g.h);
        const int strips = (g.h + sr - 1) / sr;
        const long long warps = (long long)strips * ((T + 1) / 2);
        dim3 grid((unsigned)((warps + WPB - 1) / WPB));
        launch_dependent(k_morph_mask_narrow<R, PAT>, grid, dim3(32 * WPB), 0, s, raw_bits, g, T, fbits, mask, sr, strips);
        return cudaGetLastError();
    }
    const int nslabs = (g.wpr4 + SLAB - 1) / SLAB;
    dim3 grid(nslabs, (nstrips + WPB - 1) / WPB, T);
    // Measured on a B200: 5x5 open + close at 4K 0.566 -> 0.546 ms, 3x3 open at 1080p 0.468 -> 0.478 ms (the short
    // chains are bound by the mask stores and their own arithmetic, not by the raw-row load): staged only where the
    // vertical halo is deep (64-row strips).
    if (k2_tma_enabled() && nslabs >= 2 && SR == 64) {
        constexpr int ROWS = (SR + 2 * n_ops(PAT) * R) / 2;      // rows per fill: the strip's rows pass through in two fills
        constexpr int SMEM = WPB * ((ROWS * BOXW * 4 + 127) / 128 * 128) + WPB * 8;
        CUtensorMap tmap;
        if (encode_raw_bits_tensor(&tmap, raw_bits, g.wpr_raw, g.h, T, ROWS)) {
            static PerDeviceOnce once;
            if (once.need())
                cudaFuncSetAttribute(k_morph_mask_staged<R, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
            launch_dependent(k_morph_mask_staged<R, PAT>, grid, dim3(32 * WPB), (size_t)SMEM, s, tmap, g, T, fbits, mask);
            return cudaGetLastError();
        }
    }
    launch_dependent(k_morph_mask<R, PAT, 32>, grid, dim3(32 * WPB), 0, s, raw_bits, g, T, fbits, mask);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_morph_mask(cudaStream_t s, const uint32_t* raw_bits, int T, const Geom& g,
                              const MorphCfg& m, uint32_t* fbits, uint8_t* mask, int* n_launches) {
    if (n_launches) *n_launches += 1;
    int pat = 0;
    if (m.n_ops == 4) pat = 3;
    else if (m.n_ops == 2) pat = m.is_erode[0] ? 1 : 2;
    if (pat == 0) return launch_pat<1, 0>(s, raw_bits, T, g, fbits, mask);
    if (m.radius == 2) {
        if (pat == 1) return launch_pat<2, 1>(s, raw_bits, T, g, fbits, mask);
        if (pat == 2) return launch_pat<2, 2>(s, raw_bits, T, g, fbits, mask);
        return launch_pat<2, 3>(s, raw_bits, T, g, fbits, mask);
    }
    if (pat == 1) return launch_pat<1, 1>(s, raw_bits, T, g, fbits, mask);
    if (pat == 2) return launch_pat<1, 2>(s, raw_bits, T, g, fbits, mask);
    return launch_pat<1, 3>(s, raw_bits, T, g, fbits, mask);
}

}  // namespace swb
