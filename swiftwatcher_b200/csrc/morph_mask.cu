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
// One CTA owns a 32-word x 64-row tile of one frame (+ halo), keeps it in
// shared memory, runs each erosion/dilation as a horizontal pass (funnel
// shifts across word boundaries) and a vertical pass, then writes the final
// bit words and expands them to the {0,255} uint8 mask with 16-byte stores
// that are contiguous across the warp.  It also realigns the "raw" bit columns
// produced by K1 (origin X0a) to ROI-local columns (origin roi_x0).
#include "swb_internal.cuh"

namespace swb {

namespace {

constexpr int TW = 32;         // tile width in 32-bit words (1024 pixels)
constexpr int TH = 64;         // tile height in rows
constexpr int HR_MAX = 8;      // max vertical halo: 4 ops x radius 2
constexpr int SW = TW + 2;     // smem row: one halo word each side
constexpr int SROWS = TH + 2 * HR_MAX;

__device__ __forceinline__ uint32_t colmask(int j, int w, int wpr) {
    if (j < 0 || j >= wpr) return 0u;
    int rem = w - 32 * j;
    return rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
}

template <int R>
__device__ __forceinline__ uint32_t hop(uint32_t l, uint32_t c, uint32_t r, bool erode) {
    // combine pixel x with x-R..x+R
    uint32_t acc = c;
#pragma unroll
    for (int s = 1; s <= R; ++s) {
        uint32_t right = __funnelshift_r(c, r, s);   // bit i = pixel i+s
        uint32_t left = __funnelshift_l(l, c, s);    // bit i = pixel i-s
        acc = erode ? (acc & right & left) : (acc | right | left);
    }
    return acc;
}

template <int R>
__global__ void __launch_bounds__(256)
k_morph_mask(const uint32_t* __restrict__ raw_bits, Geom g, MorphCfg m, uint32_t* __restrict__ fbits,
             uint8_t* __restrict__ mask) {
    __shared__ uint32_t bufA[SROWS][SW];
    __shared__ uint32_t bufB[SROWS][SW];

    const int f = blockIdx.z;
    const int tx0 = blockIdx.x * TW;
    const int ty0 = blockIdx.y * TH;
    const int HR = m.n_ops * R;                 // vertical halo actually needed
    const int rows_ext = TH + 2 * HR;
    const int tid = threadIdx.x;

    const uint32_t* raw_f = raw_bits + (long long)f * g.h * g.wpr_raw;

    // ---- load + realign + border fill for the first op -------------------------
    const uint32_t fill0 = (m.n_ops > 0 && m.is_erode[0]) ? 0xFFFFFFFFu : 0u;
    for (int i = tid; i < rows_ext * SW; i += blockDim.x) {
        const int rr = i / SW, cc = i - rr * SW;
        const int y = ty0 - HR + rr;
        const int j = tx0 - 1 + cc;
        uint32_t v = 0, V = 0;
        if ((unsigned)y < (unsigned)g.h) {
            V = colmask(j, g.w, g.wpr);
            if (V) {
                const uint32_t* rowp = raw_f + (long long)y * g.wpr_raw;
                uint32_t lo = rowp[j];
                uint32_t hi = (j + 1 < g.wpr_raw) ? rowp[j + 1] : 0u;
                v = __funnelshift_r(lo, hi, g.dx) & V;
            }
        }
        bufA[rr][cc] = v | (~V & fill0);
    }
    __syncthreads();

    // ---- erosion / dilation chain -----------------------------------------------
    for (int op = 0; op < m.n_ops; ++op) {
        const bool erode = m.is_erode[op] != 0;
        const uint32_t ident = erode ? 0xFFFFFFFFu : 0u;
        const uint32_t fill_next = (op + 1 < m.n_ops && m.is_erode[op + 1]) ? 0xFFFFFFFFu : 0u;
        // horizontal: A -> B
        for (int i = tid; i < rows_ext * SW; i += blockDim.x) {
            const int rr = i / SW, cc = i - rr * SW;
            uint32_t l = cc > 0 ? bufA[rr][cc - 1] : ident;
            uint32_t c = bufA[rr][cc];
            uint32_t r = cc < SW - 1 ? bufA[rr][cc + 1] : ident;
            bufB[rr][cc] = hop<R>(l, c, r, erode);
        }
        __syncthreads();
        // vertical: B -> A, then re-apply the image border for the next op
        for (int i = tid; i < rows_ext * SW; i += blockDim.x) {
            const int rr = i / SW, cc = i - rr * SW;
            uint32_t acc = bufB[rr][cc];
#pragma unroll
            for (int s = 1; s <= R; ++s) {
                uint32_t up = rr - s >= 0 ? bufB[rr - s][cc] : ident;
                uint32_t dn = rr + s < rows_ext ? bufB[rr + s][cc] : ident;
                acc = erode ? (acc & up & dn) : (acc | up | dn);
            }
            const int y = ty0 - HR + rr;
            const int j = tx0 - 1 + cc;
            const uint32_t V = ((unsigned)y < (unsigned)g.h) ? colmask(j, g.w, g.wpr) : 0u;
            bufA[rr][cc] = (acc & V) | (~V & fill_next);
        }
        __syncthreads();
    }

    // ---- outputs -------------------------------------------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < TH; r += (blockDim.x >> 5)) {
        const int y = ty0 + r;
        if (y >= g.h) break;
        const uint32_t* srow = &bufA[HR + r][1];
        // final bit words: lane = word
        {
            const int j = tx0 + lane;
            if (j < g.wpr4) fbits[((long long)f * g.h + y) * g.wpr4 + j] = srow[lane];
        }
        if (mask != nullptr) {
            uint8_t* mrow = mask + ((long long)f * g.h + y) * g.mpitch + (long long)tx0 * 32;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int wj = half * 16 + (lane >> 1);      // word within tile
                if (tx0 + wj < g.wpr) {
                    const uint32_t bits16 = (srow[wj] >> ((lane & 1) * 16)) & 0xFFFFu;
                    uint4 o;
                    // nibble -> 4 bytes of 0x00 / 0xFF
                    o.x = (((bits16 >> 0) & 0xFu) * 0x00204081u & 0x01010101u) * 0xFFu;
                    o.y = (((bits16 >> 4) & 0xFu) * 0x00204081u & 0x01010101u) * 0xFFu;
                    o.z = (((bits16 >> 8) & 0xFu) * 0x00204081u & 0x01010101u) * 0xFFu;
                    o.w = (((bits16 >> 12) & 0xFu) * 0x00204081u & 0x01010101u) * 0xFFu;
                    __stcs(reinterpret_cast<uint4*>(mrow + half * 512 + lane * 16), o);
                }
            }
        }
    }
}

}  // namespace

cudaError_t launch_morph_mask(cudaStream_t s, const uint32_t* raw_bits, int T, const Geom& g,
                              const MorphCfg& m, uint32_t* fbits, uint8_t* mask, int* n_launches) {
    dim3 grid((g.wpr + TW - 1) / TW, (g.h + TH - 1) / TH, T);
    dim3 block(256);
    if (n_launches) *n_launches += 1;
    if (m.radius == 2) k_morph_mask<2><<<grid, block, 0, s>>>(raw_bits, g, m, fbits, mask);
    else k_morph_mask<1><<<grid, block, 0, s>>>(raw_bits, g, m, fbits, mask);
    return cudaGetLastError();
}

}  // namespace swb
