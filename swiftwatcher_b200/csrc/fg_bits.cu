// K1 — temporal foreground kernel.
//
// crop_frame + convert_grayscale (image_filtering.py:199-203, :188-196), then the
// rolling temporal median / absdiff / threshold that BASELINE.json puts in place
// of rpca + bilateral_blur (data_structures.py:191-200; thresh_to_zero is
// image_filtering.py:310-316), fused into one pass that reads every BGR byte
// once and emits one bit per pixel.
//
// Design: march through time, not space.  A thread owns 16 horizontally
// adjacent pixels and walks a run of consecutive frames; the last N gray
// values of its pixels live in registers as packed u16x2 lanes (pixel k in
// the low half, pixel k+8 in the high half of register k), so the gray ring
// never touches HBM.  The median is a min/max network on VIMNMX.U16x2 (native
// on sm_100a; the u8x4 video min/max are emulated).  The time
// loop is unrolled by N so ring slots are static register names.
//
// Work unit = (256-thread column group, temporal sub-chunk of Ts frames); each
// sub-chunk re-reads its N-1 preceding frames as warm-up.
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through the runtime)

#include <cstdlib>
#include <cstring>

#include "swb_internal.cuh"

namespace swb {

namespace {

__device__ __forceinline__ uint32_t vmin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint32_t med3(uint32_t x, uint32_t y, uint32_t z) {
    return vmax2(vmin2(x, y), vmin2(vmax2(x, y), z));
}
__device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) {
    uint32_t lo = vmin2(a, b);
    b = vmax2(a, b);
    a = lo;
}

// Median of N packed u16x2 values (order statistic (N-1)/2), any odd N <= 9.
template <int N>
__device__ __forceinline__ uint32_t median_lanes(const uint32_t (&v)[N]) {
    if constexpr (N == 1) {
        return v[0];
    } else if constexpr (N == 3) {
        return med3(v[0], v[1], v[2]);
    } else if constexpr (N == 5) {
        // med5 = med3(e, max(min(a,b), min(c,d)), min(max(a,b), max(c,d)))
        uint32_t lo = vmax2(vmin2(v[0], v[1]), vmin2(v[2], v[3]));
        uint32_t hi = vmin2(vmax2(v[0], v[1]), vmax2(v[2], v[3]));
        return med3(v[4], lo, hi);
    } else if constexpr (N == 9) {
        // sort three triples, then med3(max of mins, med of meds, min of maxes)
        uint32_t a[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) a[i] = v[i];
#pragma unroll
        for (int c = 0; c < 9; c += 3) {
            cswap(a[c], a[c + 1]);
            cswap(a[c + 1], a[c + 2]);
            cswap(a[c], a[c + 1]);
        }
        uint32_t lo = vmax2(vmax2(a[0], a[3]), a[6]);
        uint32_t hi = vmin2(vmin2(a[2], a[5]), a[8]);
        uint32_t mid = med3(a[1], a[4], a[7]);
        return med3(lo, mid, hi);
    } else {
        // odd-even transposition sort; the compiler prunes the unused outputs
        uint32_t a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = v[i];
#pragma unroll
        for (int r = 0; r < N; ++r) {
#pragma unroll
            for (int i = (r & 1); i + 1 < N; i += 2) cswap(a[i], a[i + 1]);
        }
        return a[N / 2];
    }
}

// ---- pipe-balanced medians for the v2 kernel -----------------------------------------
// K1 is bound by the ALU pipe (VIMNMX, LOP3, PRMT ...), while the FMA pipe idles.  The
// median of three is x + y + z - min3 - max3: two VIMNMX3 on the ALU pipe and four adds
// that are forced onto the FMA pipe as IMAD (multiplier `one` is a kernel argument equal
// to 1, so ptxas cannot turn the IMAD back into an ALU-pipe IADD3).  u16x2 lanes never
// carry or borrow here (sums <= 765, and sum >= min + max per lane).
struct FmaAdd {
    uint32_t one, mone;
    __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) const {
        uint32_t d;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
        return d;
    }
    __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) const {   // a - b
        uint32_t d;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(mone), "r"(a));
        return d;
    }
    __device__ __forceinline__ uint32_t med3(uint32_t x, uint32_t y, uint32_t z) const {
        const uint32_t s = add(add(x, y), z);
        return sub(sub(s, __vimin3_u16x2(x, y, z)), __vimax3_u16x2(x, y, z));
    }
};

template <int N>
__device__ __forceinline__ uint32_t median_lanes_fma(const uint32_t (&v)[N], const FmaAdd& f) {
    if constexpr (N == 3) {
        return f.med3(v[0], v[1], v[2]);
    } else if constexpr (N == 5) {
        // pairs: min on the ALU pipe, max = a + b - min on the FMA pipe
        const uint32_t mn1 = vmin2(v[0], v[1]), mx1 = f.sub(f.add(v[0], v[1]), mn1);
        const uint32_t mn2 = vmin2(v[2], v[3]), mx2 = f.sub(f.add(v[2], v[3]), mn2);
        return f.med3(v[4], vmax2(mn1, mn2), vmin2(mx1, mx2));
    } else if constexpr (N == 9) {
        // three sorted triples (min3, max3, middle by sum), then med3(max of mins, med of
        // middles, min of maxes)
        uint32_t lo[3], hi[3], mid[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t a = v[3 * c], b = v[3 * c + 1], d = v[3 * c + 2];
            lo[c] = __vimin3_u16x2(a, b, d);
            hi[c] = __vimax3_u16x2(a, b, d);
            mid[c] = a + b + d - lo[c] - hi[c];      // two IADD3 (fewer instructions; ALU has room here)
        }
        const uint32_t L = __vimax3_u16x2(lo[0], lo[1], lo[2]);
        const uint32_t H = __vimin3_u16x2(hi[0], hi[1], hi[2]);
        const uint32_t M = f.med3(mid[0], mid[1], mid[2]);
        return f.med3(L, M, H);
    } else {
        return median_lanes<N>(v);
    }
}

// cv2 4.13 BGR2GRAY: (3735 B + 19235 G + 9798 R + 2^14) >> 15, evaluated with all
// terms doubled so that the result is byte 2 of the sum (sum < 2^24).
__device__ __forceinline__ uint32_t gray_sum(uint32_t b, uint32_t g, uint32_t r) {
    return 7470u * b + 38470u * g + 19596u * r + 32768u;
}

// 48 BGR bytes (12 words, 16 pixels) -> 8 registers of (gray[k], gray[k+8]) u16x2.
__device__ __forceinline__ void bgr48_to_lanes(const uint32_t (&w)[12], uint32_t (&out)[8]) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int o = 3 * i;
        uint32_t b = (w[o >> 2] >> (8 * (o & 3))) & 0xFFu;
        uint32_t g = (w[(o + 1) >> 2] >> (8 * ((o + 1) & 3))) & 0xFFu;
        uint32_t r = (w[(o + 2) >> 2] >> (8 * ((o + 2) & 3))) & 0xFFu;
        s[i] = gray_sum(b, g, r);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = __byte_perm(s[k], s[k + 8], 0x7632);
}

// 16 gray bytes (4 words) -> lanes
__device__ __forceinline__ void gray16_to_lanes(const uint32_t (&w)[4], uint32_t (&out)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // byte k (word k>>2) into bits 0..7, byte k+8 (word (k>>2)+2) into bits 16..23
        uint32_t lo = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        uint32_t hi = (w[(k >> 2) + 2] >> (8 * (k & 3))) & 0xFFu;
        out[k] = lo | (hi << 16);
    }
}

__device__ __forceinline__ uint4 lanes_to_gray16(const uint32_t (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        // bytes 4q..4q+3 (low halves) and 8+4q..8+4q+3 (high halves)
        uint32_t lo01 = __byte_perm(v[4 * q + 0], v[4 * q + 1], 0x0040);  // b0=v0.b0 b1=v1.b0
        uint32_t lo23 = __byte_perm(v[4 * q + 2], v[4 * q + 3], 0x0040);
        uint32_t hi01 = __byte_perm(v[4 * q + 0], v[4 * q + 1], 0x0062);  // b0=v0.b2 b1=v1.b2
        uint32_t hi23 = __byte_perm(v[4 * q + 2], v[4 * q + 3], 0x0062);
        w[q] = __byte_perm(lo01, lo23, 0x5410);
        w[q + 2] = __byte_perm(hi01, hi23, 0x5410);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int C, bool ALIGNED>
struct RawPixels {
    uint32_t w[C == 3 ? 12 : 4];
};

// Load the 16 pixels of this thread from a BGR/gray source row.
template <int C, bool ALIGNED>
__device__ __forceinline__ void load_raw(RawPixels<C, ALIGNED>& p, const uint8_t* ptr, int avail_px) {
    constexpr int NW = (C == 3) ? 12 : 4;
    if constexpr (ALIGNED) {
        const uint4* q = reinterpret_cast<const uint4*>(ptr);
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            uint4 v = __ldcs(q + i);
            p.w[4 * i + 0] = v.x;
            p.w[4 * i + 1] = v.y;
            p.w[4 * i + 2] = v.z;
            p.w[4 * i + 3] = v.w;
        }
    } else {
        // guarded byte loads: pixels at or beyond avail_px read as 0
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int byte = 4 * i + b;
                int px = byte / C;
                if (px < avail_px) v |= (uint32_t)ptr[byte] << (8 * b);
            }
            p.w[i] = v;
        }
    }
}

template <int C, bool ALIGNED>
__device__ __forceinline__ void raw_to_lanes(const RawPixels<C, ALIGNED>& p, uint32_t (&out)[8]) {
    if constexpr (C == 3) bgr48_to_lanes(p.w, out);
    else gray16_to_lanes(p.w, out);
}

// |x - m| > thresh per u16 lane -> 16 result bits (bit k = pixel k)
__device__ __forceinline__ uint32_t fg_bits16(const uint32_t (&cur)[8], const uint32_t (&med)[8],
                                              uint32_t bias /* (0x7FFF - thresh) * 0x10001 */) {
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t d = vmax2(cur[k], med[k]) - vmin2(cur[k], med[k]);  // per-lane |x-m|, no borrow
        uint32_t t = d + bias;                                       // bit 15 / 31 = (d > thresh)
        acc |= (t >> (15 - k)) & (0x00010001u << k);
    }
    return (acc & 0xFFu) | ((acc >> 8) & 0xFF00u);
}

template <int N, int C, bool ALIGNED>
__global__ void __launch_bounds__(256)
k_fg_bits(FrameSrc src, int T, int Ts, int h, int wa, int thresh, uint16_t* __restrict__ raw_bits) {
    const int gpr = wa >> 4;  // 16-pixel groups per row
    const int G = h * gpr;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int row = g / gpr;
    const int col = g - row * gpr;
    const int t_start = blockIdx.y * Ts;
    const int t_end = min(T, t_start + Ts);
    const int n_out = t_end - t_start;
    if (n_out <= 0) return;

    const long long pix_off = (long long)row * src.pitch + (long long)col * 16 * C;
    const int avail_px = src.avail_w - col * 16;
    const uint32_t bias = (uint32_t)(0x7FFF - thresh) * 0x00010001u;

    uint32_t ring[N][8];

    auto load_frame = [&](int j, uint32_t (&dst)[8]) {
        // frame j relative to the first output frame of this submit
        if (j < 0 && src.hist_valid) {
            // carried history: N-1 compact gray frames [h][wa] in our own buffer
            const int slot = j + (N - 1);
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src.hist + ((long long)slot * h + row) * wa + col * 16));
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            gray16_to_lanes(w4, dst);
        } else {
            if (j < -src.n_inline_halo) j = -src.n_inline_halo;  // replicate earliest frame
            RawPixels<C, ALIGNED> p;
            load_raw<C, ALIGNED>(p, src.cur + (long long)j * src.frame_stride + pix_off, avail_px);
            raw_to_lanes<C, ALIGNED>(p, dst);
        }
    };

    // warm-up: frames t_start-(N-1) .. t_start-1 into slots 0..N-2
#pragma unroll
    for (int s = 0; s < N - 1; ++s) load_frame(t_start - (N - 1) + s, ring[s]);

    RawPixels<C, ALIGNED> nxt;
    load_raw<C, ALIGNED>(nxt, src.cur + (long long)t_start * src.frame_stride + pix_off, avail_px);

    const bool write_hist = (src.hist_out != nullptr) && (t_end == T);
    uint16_t* out = raw_bits + ((long long)t_start * h + row) * gpr + col;
    const long long out_step = (long long)h * gpr;

    for (int base = 0; base < n_out; base += N) {
#pragma unroll
        for (int p = 0; p < N; ++p) {
            const int k = base + p;
            if (k < n_out) {
                const int slot = (N - 1 + p) % N;  // static after unrolling
                raw_to_lanes<C, ALIGNED>(nxt, ring[slot]);
                if (k + 1 < n_out)
                    load_raw<C, ALIGNED>(nxt, src.cur + (long long)(t_start + k + 1) * src.frame_stride + pix_off,
                                         avail_px);
                uint32_t med[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t v[N];
#pragma unroll
                    for (int s = 0; s < N; ++s) v[s] = ring[s][q];
                    med[q] = median_lanes<N>(v);
                }
                out[(long long)k * out_step] = (uint16_t)fg_bits16(ring[slot], med, bias);
                if (write_hist && k == n_out - 1) {
                    // last N-1 gray frames, oldest first: hist[s] = frame (T-1) - (N-2-s)
#pragma unroll
                    for (int s = 0; s < N - 1; ++s) {
                        const int m = N - 2 - s;                  // frames back from the newest
                        const int hs = ((slot - m) % N + N) % N;  // static
                        *reinterpret_cast<uint4*>(src.hist_out + ((long long)s * h + row) * wa + col * 16) =
                            lanes_to_gray16(ring[hs]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// v2: the same march through time, fed by the bulk-copy engine.
//
// One elected thread streams the CTA's 256 x 16-pixel slice of every frame into a
// ring of shared-memory stages with cp.async.bulk (UBLKCP) completing on mbarriers;
// all warps consume a stage with three conflict-free LDS.128 per thread, release it
// with one mbarrier arrive per warp, and the producer refills a stage one step after
// it was consumed.  Bytes in flight no longer cost registers, so the median/threshold
// math overlaps the HBM latency at 2 CTAs per SM.
//
// Instruction diet (the v1 kernel was ALU-pipe bound, not HBM bound):
//  * gray = two IDP.2A (16-bit weights x packed pixel bytes, no byte extraction):
//    cv2's (3735 B + 19235 G + 9798 R + 2^14) >> 15 with every term doubled, so the
//    result is byte 2 of the accumulator;
//  * |x - median| = one VABSDIFF4 on the u16x2 lanes (high bytes are zero);
//  * (d > thresh) = one VIADDMNMX.S16x2.RELU: max(min(d - thresh, 1), 0) -> 0/1 per lane,
//    shifted into place and accumulated by one IMAD.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// time a waiting thread may stay suspended inside one mbarrier.try_wait before the loop around it spins again: with
// the default (short) limit the waiting warps of the N = 9 kernel spent 5 % of all issued instructions in that loop
constexpr uint32_t MBAR_SUSPEND_NS = 1000000u;
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    int spins = 0;
    const long long t_wait0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(MBAR_SUSPEND_NS)
            : "memory");
        if (done) break;
        // never hang the GPU on a pipeline bug: give up after ~2 s of SM clock (a try_wait may suspend for up to 1 ms)
        if ((++spins & 63) == 0 && clock64() - t_wait0 > (1ll << 32)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One TMA tensor load: the box (row bytes, rows, 1 frame) at (x = 0, row y, frame z) of the ROI tensor
// -> shared memory, completing on the mbarrier (SASS UTMALDG).  Out-of-range rows are zero-filled.
__device__ __forceinline__ void tma_load_box(void* dst_smem, const CUtensorMap* tmap, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(0), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// 2L BGR pixels (6L bytes, 3L/2 words) -> L u16x2 lanes of (gray[k], gray[k+L]) with IDP.2A
template <int L>
__device__ __forceinline__ void bgr_to_lanes_dp(const uint32_t (&w)[3 * L / 2], uint32_t (&out)[L]) {
    constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u, RND = 32768u;
    constexpr int NW = 3 * L / 2;
    uint32_t s[2 * L];
#pragma unroll
    for (int i = 0; i < 2 * L; ++i) {
        const int o = 3 * i, wi = o >> 2, r = o & 3;
        uint32_t acc;
        if (r == 0) {          // B G R x
            acc = __dp2a_lo(WB | (WG << 16), w[wi], RND);
            acc = __dp2a_hi(WR, w[wi], acc);
        } else if (r == 1) {   // x B G R
            acc = __dp2a_lo(WB << 16, w[wi], RND);
            acc = __dp2a_hi(WG | (WR << 16), w[wi], acc);
        } else if (r == 2) {   // x x B G | R
            acc = __dp2a_hi(WB | (WG << 16), w[wi], RND);
            acc = __dp2a_lo(WR, w[(wi + 1) % NW], acc);
        } else {               // x x x B | G R
            acc = __dp2a_hi(WB << 16, w[wi], RND);
            acc = __dp2a_lo(WG | (WR << 16), w[(wi + 1) % NW], acc);
        }
        s[i] = acc;
    }
#pragma unroll
    for (int k = 0; k < L; ++k) out[k] = __byte_perm(s[k], s[k + L], 0x7632);
}

// 2L gray bytes (L/2 words) -> lanes
template <int L>
__device__ __forceinline__ void gray_to_lanes(const uint32_t (&w)[L / 2], uint32_t (&out)[L]) {
#pragma unroll
    for (int k = 0; k < L; ++k) {
        const uint32_t lo = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        const uint32_t hi = (w[(k + L) >> 2] >> (8 * ((k + L) & 3))) & 0xFFu;
        out[k] = lo | (hi << 16);
    }
}

// lanes -> 2L gray bytes (word t = pixels 4t..4t+3; pixel k < L is the low half of lane k,
// pixel k >= L the high half of lane k - L): three PRMT per word
template <int L>
__device__ __forceinline__ void lanes_to_gray(const uint32_t (&v)[L], uint32_t (&w)[L / 2]) {
#pragma unroll
    for (int t = 0; t < L / 4; ++t) {
        const uint32_t p01 = __byte_perm(v[4 * t + 0], v[4 * t + 1], 0x6240);   // (a.lo, b.lo, a.hi, b.hi)
        const uint32_t p23 = __byte_perm(v[4 * t + 2], v[4 * t + 3], 0x6240);
        w[t] = __byte_perm(p01, p23, 0x5410);
        w[t + L / 4] = __byte_perm(p01, p23, 0x7632);
    }
}

// (|x - m| > thresh) per u16 lane -> 2L bits (bit k = pixel k)
template <int L>
__device__ __forceinline__ uint32_t fg_bits_v2(const uint32_t (&cur)[L], const uint32_t (&med)[L],
                                               uint32_t neg_thresh_x2) {
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        const uint32_t d = __vabsdiffu4(cur[k], med[k]);                          // lanes hold 0..255
        const uint32_t f = __viaddmin_s16x2_relu(d, neg_thresh_x2, 0x00010001u);  // 1 where d > thresh
        acc += f << k;
    }
    constexpr uint32_t M = (1u << L) - 1u;
    return (acc & M) | ((acc >> (16 - L)) & (M << L));
}

// ---- shared-core sliding medians -------------------------------------------------------
// Consecutive windows share most of their frames, so K outputs are produced per loop
// iteration from one partially sorted "core" plus a few extras (all on packed u16x2 lanes):
//  * N = 5, K = 2: windows {t-4..t} and {t-3..t+1} share {t-3..t}.  With the core kept as two
//    sorted pairs, its two middle values s2 <= s3 cost 4 min/max, and the median of five is
//    clamp(extra, s2, s3).  The pair (t-1, t) is the old pair of the next iteration.
//    5 min/max per output instead of 14.
//  * N = 9, K = 3: windows of outputs t, t+1, t+2 share {t-6..t-1} = two time-aligned triples
//    that are sorted once (and reused by the next iteration).  Their merged middle four
//    m2..m5 cost 10 min/max; every window adds three extras, sorted (y1 <= y2 <= y3), and the
//    fifth smallest of the nine is the fourth smallest of {m2..m5} U {y1..y3}
//    = min(m5, max(m4, y1), max(m3, y2), max(m2, y3)).  ~15 min/max per output instead of 26.
// a + b - min(a, b) stands in for max(a, b) where that moves work from the ALU pipe to the
// FMA pipe (lanes never carry: per-lane sums stay below 2^16 and above each term).
template <int N>
__host__ __device__ constexpr int group_k() { return N == 5 ? 2 : (N == 9 ? 3 : 1); }

struct Sorted3 { uint32_t lo, mid, hi; };
__device__ __forceinline__ Sorted3 sort3_lanes(uint32_t a, uint32_t b, uint32_t c, const FmaAdd& f) {
    Sorted3 s;
    s.lo = __vimin3_u16x2(a, b, c);
    s.hi = __vimax3_u16x2(a, b, c);
    s.mid = f.sub(f.sub(f.add(f.add(a, b), c), s.lo), s.hi);
    return s;
}
// fourth smallest of sorted (m2 <= m3 <= m4 <= m5) U sorted (y.lo <= y.mid <= y.hi)
__device__ __forceinline__ uint32_t select4of7(uint32_t m2, uint32_t m3, uint32_t m4, uint32_t m5, const Sorted3& y) {
    return vmin2(m5, __vimin3_u16x2(vmax2(m4, y.lo), vmax2(m3, y.mid), vmax2(m2, y.hi)));
}

// Temporal sub-chunk length: long enough that the N-1 warm-up frames are a few
// percent of the work, short enough that the grid has several waves of CTAs; a
// multiple of 6 so that only the last sub-chunk of a submit has a partial group.
thread_local int g_last_ts = 0;   // what the last launch on this thread used (swb_last_subchunk)

int pick_ts_impl(int T, int n_col_blocks, int median_n, int gpu_share) {
    static const int forced = [] { const char* e = getenv("SWB_K1_TS"); return e ? atoi(e) : 0; }();
    if (forced > 0) return std::min(T, (forced + 5) / 6 * 6);
    // several waves of CTAs when the context has the GPU to itself; its share of that when other contexts
    // (videos) run beside it: their CTAs fill the machine together, and longer sub-chunks re-read less
    const int target_ctas = gpu_share > 1 ? std::max(148 * 2 * 2 / gpu_share, 16) : 148 * 2 * 4;
    int ts = T;
    while (ts > 32 && (long long)n_col_blocks * ((T + ts - 1) / ts) < target_ctas) ts = (ts + 1) / 2;
    const int min_ts = 8 * (median_n - 1) > 0 ? 8 * (median_n - 1) : 1;   // <= 12.5% warm-up
    if (ts < min_ts) ts = min_ts;
    ts = (ts + 5) / 6 * 6;
    if (ts > T) ts = T;
    if (ts < 1) ts = 1;
    return ts;
}
thread_local int g_forced_ts = 0;  // swb_set_option("temporal_subchunk"): set around a launch by launch_fg_bits
int pick_ts(int T, int n_col_blocks, int median_n, int gpu_share) {
    if (g_forced_ts > 0) return g_last_ts = std::min(T, (g_forced_ts + 5) / 6 * 6);
    return g_last_ts = pick_ts_impl(T, n_col_blocks, median_n, gpu_share);
}
// Cropped ROIs (tensor-map tiles): the kernel is short (~0.1 ms per 2048 frames of 320x240) and every sub-chunk
// re-reads its (N-1)-frame warm-up, so ONE wave of long sub-chunks beats four waves of short ones (measured at
// 320x240 / T = 2048: 0.1149 ms with one wave, 0.1175 with two, 0.1193 with three, 0.1211 with four).
int roi_waves() {
    static const int w = [] { const char* e = getenv("SWB_K1_ROI_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 1; }();
    return w;
}
// ... and the sub-chunk length is computed, not halved down, so that the grid is just UNDER a whole number of waves
// (20 row blocks x 14 sub-chunks of 150 frames = 280 of 296 CTA slots at 320x240 / T = 2048).
int pick_ts_roi(int T, int n_col_blocks, int median_n, int gpu_share, int occ) {
    static const bool env_ts = getenv("SWB_K1_TS") != nullptr;
    if (g_forced_ts > 0 || gpu_share > 1 || env_ts) return pick_ts(T, n_col_blocks, median_n, gpu_share);
    const int chunks = std::max(1, roi_waves() * 148 * occ / std::max(n_col_blocks, 1));
    int ts = (T + chunks - 1) / chunks;
    ts = std::max(ts, std::max(8 * (median_n - 1), 1));
    ts = std::min((ts + 5) / 6 * 6, T);
    return g_last_ts = std::max(ts, 1);
}

// ---- tensor map of the ROI inside the frame stack: (row bytes / 8, ROI rows, frames) of 8-byte elements ----
bool tma_enabled() {
    static const bool on = [] { const char* e = getenv("SWB_K1_TMA"); return !(e && e[0] == '0'); }();
    return on;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool encode_roi_tensor(CUtensorMap* tmap, const uint8_t* base, long long row_bytes, int rows_total, int frames,
                       long long pitch, long long frame_stride, int box_rows) {
    static EncodeTiledFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<EncodeTiledFn>(fn);
    }();
    if (!encode) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (frame_stride & 15) || (row_bytes & 15)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)(row_bytes / 8), (cuuint64_t)rows_total, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)(row_bytes / 8), (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (box[0] > 256 || box[1] > 256) return false;
    return encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int CONSUMERS = 256;               // 8 consumer warps: one pixel group per thread
constexpr int V2_THREADS = CONSUMERS + 32;   // + 1 producer warp

// L = u16x2 registers per ring slot: a thread owns 2L adjacent pixels (16 for N <= 5; 8 for the
// longer windows, whose ring would not fit the register file at two CTAs per SM).
template <int N, int C, int S, int L, int OCC>
__global__ void __launch_bounds__(V2_THREADS, OCC)
k_fg_bits_v2(FrameSrc src, int T, int Ts, int h, int wa, int thresh, uint32_t one,
             uint8_t* __restrict__ raw_bits, const __grid_constant__ CUtensorMap tmap, int tile_rows) {
    constexpr int PPT = 2 * L;               // pixels per thread
    constexpr int TB = PPT * C;              // bytes per thread per frame
    constexpr int STAGE_BYTES = CONSUMERS * TB;
    constexpr int NWORDS = TB / 4;
    constexpr int K = group_k<N>();          // outputs per loop iteration
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
    uint64_t* empty = full + S;

    const int tid = threadIdx.x;
    const int gpr = wa / PPT;                // pixel groups per row
    const int G = h * gpr;
    // tile_rows > 0: the CTA owns tile_rows whole ROI rows (a TMA box); else 256 consecutive groups
    const int cta_groups = tile_rows > 0 ? tile_rows * gpr : CONSUMERS;
    const int g0 = blockIdx.x * cta_groups;
    const int t_start = blockIdx.y * Ts;
    const int t_end = min(T, t_start + Ts);
    const int n_out = t_end - t_start;
    if (n_out <= 0) return;                  // block-uniform
    const int n_iter = (n_out + K - 1) / K;
    const int n_total = K * n_iter + N - 1;  // frames walked: j_first .. (a partial last group re-reads frame T-1)
    const int j_first = t_start - (N - 1);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= CONSUMERS) {
        // ===== producer warp: streams this CTA's slice of every frame =====
        // (groups per row and per CTA are multiples of 4, so every copy is 16-byte granular)
        const int lane = tid - CONSUMERS;
        const int ngroups = min(cta_groups, G - g0);
        const uint32_t bytes = (uint32_t)(ngroups * TB);
        const bool contiguous = (src.pitch == (long long)gpr * TB);
        if ((contiguous || tile_rows > 0) && lane != 0) return;   // one copy per frame: one lane is enough
        // Rows of a cropped ROI are not adjacent in memory: the slice is one copy per row segment.
        // The lanes of the warp issue them side by side (lane i takes segments i, i + 32, ...):
        // a single lane issuing 12+ small copies per frame was slower than the eight consumer warps.
        const int r0 = g0 / gpr, c0 = g0 - r0 * gpr;
        const int len0 = min(gpr - c0, ngroups);      // first segment (the rest of row r0)
        const int nseg = contiguous ? 1 : 1 + (ngroups - len0 + gpr - 1) / gpr;
        for (int p = 0; p < n_total; ++p) {
            const int st = p % S;
            if (p >= S) mbar_wait(&empty[st], (uint32_t)((p / S - 1) & 1));
            int j = min(j_first + p, T - 1);
            uint8_t* dst = smem + st * STAGE_BYTES;
            if (j < 0 && src.hist_valid) {                  // carried history: compact gray frames
                if (lane == 0) {
                    const uint8_t* hf = src.hist + (long long)(j + (N - 1)) * h * wa;
                    mbar_arrive_expect_tx(&full[st], (uint32_t)(ngroups * PPT));
                    bulk_g2s(dst, hf + (long long)g0 * PPT, (uint32_t)(ngroups * PPT), &full[st]);
                }
                continue;
            }
            if (j < -src.n_inline_halo) j = -src.n_inline_halo;   // replicate the earliest frame
            const uint8_t* fr = src.cur + (long long)j * src.frame_stride;
            if (tile_rows > 0) {
                // the rows of a cropped ROI are apart in memory: one tensor-map box load brings the tile
                // (tile_rows x row bytes; rows past the ROI are zero-filled and still counted by the barrier)
                mbar_arrive_expect_tx(&full[st], (uint32_t)(tile_rows * gpr * TB));
                tma_load_box(dst, &tmap, blockIdx.x * tile_rows, j + src.n_inline_halo, &full[st]);
            } else if (contiguous) {
                mbar_arrive_expect_tx(&full[st], bytes);
                bulk_g2s(dst, fr + (long long)g0 * TB, bytes, &full[st]);
            } else {
                if (lane == 0) mbar_arrive_expect_tx(&full[st], bytes);
                __syncwarp();                               // the expectation is posted before any copy can complete
                for (int i = lane; i < nseg; i += 32) {
                    const int off = (i == 0) ? 0 : len0 + (i - 1) * gpr;      // first group of the segment within the slice
                    const int n = (i == 0) ? len0 : min(gpr, ngroups - off);
                    bulk_g2s(dst + off * TB, fr + (long long)(r0 + i) * src.pitch + (long long)(i == 0 ? c0 : 0) * TB,
                             (uint32_t)(n * TB), &full[st]);
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int g = g0 + tid;
    const bool active = g < G && tid < cta_groups;
    const int row = active ? g / gpr : 0;
    const int col = active ? g - row * gpr : 0;
    const uint32_t neg_th = ((uint32_t)(-thresh) & 0xFFFFu) * 0x00010001u;
    FmaAdd fa;
    fa.one = one;
    fa.mone = 0u - one;
    // raw bits: one bit per pixel, PPT / 8 bytes per thread per frame
    uint8_t* out = raw_bits + (((long long)t_start * h + row) * gpr + col) * (PPT / 8);
    const long long out_step = (long long)h * gpr * (PPT / 8);
    const uint8_t* my_smem = smem + tid * TB;
    const bool lane0 = (tid & 31) == 0;
    // the last N-1 gray frames of the submit (pipeline frames hist_from .. hist_to) are left
    // for the next submit as they pass through (only the last temporal sub-chunk sees them)
    const int hist_to = (T - 1) - j_first;
    const int hist_from = (src.hist_out != nullptr && t_end == T) ? hist_to - (N - 2) : 0x7FFFFFFF;

    // take pipeline frame p (held by stage st, barrier phase parity `par`) as packed gray lanes;
    // `maybe_hist`: the stage may hold a carried gray frame (only the warm-up frames of the first
    // temporal sub-chunk can).  In the grouped loops st is a compile-time constant after unrolling.
    auto consume_at = [&](int st, uint32_t par, int p, uint32_t (&dst)[L], bool maybe_hist) {
        mbar_wait(&full[st], par);
        if (C == 3 && maybe_hist && src.hist_valid && j_first + p < 0) {   // block-uniform
            uint32_t gw[L / 2];
            const uint32_t* sg = reinterpret_cast<const uint32_t*>(smem + st * STAGE_BYTES + tid * PPT);
#pragma unroll
            for (int i = 0; i < L / 2; ++i) gw[i] = sg[i];
            gray_to_lanes<L>(gw, dst);
        } else {
            uint32_t w[NWORDS];
            if constexpr (TB % 16 == 0) {
                const uint4* sp4 = reinterpret_cast<const uint4*>(my_smem + st * STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < TB / 16; ++i) {
                    const uint4 v = sp4[i];
                    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                }
            } else {
                const uint2* sp2 = reinterpret_cast<const uint2*>(my_smem + st * STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < TB / 8; ++i) {
                    const uint2 v = sp2[i];
                    w[2 * i] = v.x; w[2 * i + 1] = v.y;
                }
            }
            if constexpr (C == 3) bgr_to_lanes_dp<L>(w, dst);
            else gray_to_lanes<L>(w, dst);
        }
        // release the stage only after the loaded words have been consumed (the converted
        // lanes depend on every LDS), so the bulk engine can never overwrite data in flight
        asm volatile("" ::"r"(dst[0]), "r"(dst[L - 1]) : "memory");
        __syncwarp();
        if (lane0) mbar_arrive(&empty[st]);
    };
    auto consume = [&](int p, uint32_t (&dst)[L], bool maybe_hist) {
        consume_at(p % S, (uint32_t)((p / S) & 1), p, dst, maybe_hist);
    };
    // leave pipeline frame p for the next submit if it is one of the last N-1 frames (rare)
    auto keep_for_next = [&](int p, const uint32_t (&v)[L]) {
        if (p >= hist_from && p <= hist_to && active) {
            uint32_t hw[L / 2];
            lanes_to_gray<L>(v, hw);
            uint32_t* hp = reinterpret_cast<uint32_t*>(
                src.hist_out + (((long long)(p - hist_from) * h + row) * gpr + col) * PPT);
#pragma unroll
            for (int i = 0; i < L / 2; ++i) hp[i] = hw[i];
        }
    };
    const bool short_tail = hist_from < N - 1;   // warm-up frames are among the last N-1 (n_out < N-1)
    // foreground flag of one lane (two pixels): 1 per u16 half where |x - median| > thresh
    auto fg_flag = [&](uint32_t x, uint32_t med) -> uint32_t {
        return __viaddmin_s16x2_relu(__vabsdiffu4(x, med), neg_th, 0x00010001u);
    };
    // store the bits of output k (frame t_start + k); acc = sum over lanes q of flag << q
    auto emit_acc = [&](int k, uint32_t acc) {
        constexpr uint32_t M = (1u << L) - 1u;
        const uint32_t bits = (acc & M) | ((acc >> (16 - L)) & (M << L));
        if (active && k < n_out) {
            uint8_t* o = out + (long long)k * out_step;
            if constexpr (PPT == 16) *reinterpret_cast<uint16_t*>(o) = (uint16_t)bits;
            else *o = (uint8_t)bits;
        }
    };

    if constexpr (N == 5) {
        // ---- two outputs per iteration (see "shared-core sliding medians") ----
        // iteration m: outputs t, t+1 (t = t_start + 2m; frame t is pipeline frame 2m + 4)
        uint32_t ev[2][L];        // raw even pipeline frames: ev[m & 1] = frame t-4, ev[(m+1) & 1] = frame t-2
        uint32_t odd[L];          // raw frame t-1
        uint32_t plo[L], phi[L];  // sorted pair (t-3, t-2)
        static_assert(S == 4, "N = 5: four frames per unrolled loop body = four stages");
        {
            uint32_t a[L];
            consume(0, ev[0], true);
            consume(1, a, true);
            consume(2, ev[1], true);
            consume(3, odd, true);
            if (short_tail) {
                keep_for_next(0, ev[0]);
                keep_for_next(1, a);
                keep_for_next(2, ev[1]);
                keep_for_next(3, odd);
            }
#pragma unroll
            for (int q = 0; q < L; ++q) {
                plo[q] = vmin2(a[q], ev[1][q]);
                phi[q] = vmax2(a[q], ev[1][q]);
            }
        }
        for (int m0 = 0; m0 < n_iter; m0 += 2) {
            const uint32_t par = (uint32_t)(((m0 >> 1) + 1) & 1);   // frames 4 + 2 m0 + i: stage i, round 1 + m0 / 2
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int m = m0 + u;
                if (m >= n_iter) break;                         // block-uniform; no state is live after the loop
                uint32_t x0[L], x1[L], s2[L], s3[L];
                consume_at(2 * u, par, 2 * m + 4, x0, false);
                uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
                for (int q = 0; q < L; ++q) {
                    const uint32_t qlo = vmin2(odd[q], x0[q]);
                    const uint32_t qhi = fa.sub(fa.add(odd[q], x0[q]), qlo);
                    const uint32_t a = vmax2(plo[q], qlo), b = vmin2(phi[q], qhi);
                    s2[q] = vmin2(a, b);
                    s3[q] = fa.sub(fa.add(a, b), s2[q]);
                    acc0 += fg_flag(x0[q], vmax2(s2[q], vmin2(ev[u][q], s3[q]))) << q;
                    plo[q] = qlo;
                    phi[q] = qhi;
                    ev[u][q] = x0[q];
                }
                emit_acc(2 * m, acc0);
                consume_at(2 * u + 1, par, 2 * m + 5, x1, false);
#pragma unroll
                for (int q = 0; q < L; ++q) {
                    acc1 += fg_flag(x1[q], vmax2(s2[q], vmin2(x1[q], s3[q]))) << q;
                    odd[q] = x1[q];
                }
                emit_acc(2 * m + 1, acc1);
                if (2 * m + 5 >= hist_from) {                   // block-uniform, last frames of the submit only
                    keep_for_next(2 * m + 4, ev[u]);
                    keep_for_next(2 * m + 5, x1);
                }
            }
        }
    } else if constexpr (N == 9) {
        // ---- three outputs per iteration ----
        // iteration m: outputs t, t+1, t+2 (t = t_start + 3m; frame t is pipeline frame 3m + 8).
        // Triples are aligned to t_start: {t-9,t-8,t-7}, A = {t-6,t-5,t-4}, B = {t-3,t-2,t-1}.
        // The raw last two frames of each triple are needed again three iterations later (as the
        // extras t-8, t-7): they wait in a thread-private shared-memory ring, packed u8x4
        // (first | second << 8), slot m % 3 -- read at the top of iteration m, rewritten at its end.
        static_assert(L == 4, "one 16-byte ring entry per thread");
        uint4* rw_ring = reinterpret_cast<uint4*>(smem + S * STAGE_BYTES + 2 * S * 8) + tid;   // [3][CONSUMERS]
        uint32_t st[2][3][L];     // sorted triples: st[m & 1] = A, st[(m + 1) & 1] = B
        static_assert(S == 6, "N = 9: six frames per unrolled loop body = six stages");
        {
            uint32_t a[L], b[L], c[L];
            consume(0, a, true);
            consume(1, b, true);
            if (short_tail) { keep_for_next(0, a); keep_for_next(1, b); }
            rw_ring[0] = make_uint4(a[0] | (b[0] << 8), a[1] | (b[1] << 8), a[2] | (b[2] << 8), a[3] | (b[3] << 8));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                consume(2 + 3 * i, a, true);
                consume(3 + 3 * i, b, true);
                consume(4 + 3 * i, c, true);
                if (short_tail) { keep_for_next(2 + 3 * i, a); keep_for_next(3 + 3 * i, b); keep_for_next(4 + 3 * i, c); }
#pragma unroll
                for (int q = 0; q < L; ++q) {
                    const Sorted3 s = sort3_lanes(a[q], b[q], c[q], fa);
                    st[i][0][q] = s.lo; st[i][1][q] = s.mid; st[i][2][q] = s.hi;
                }
                rw_ring[(1 + i) * CONSUMERS] =
                    make_uint4(b[0] | (c[0] << 8), b[1] | (c[1] << 8), b[2] | (c[2] << 8), b[3] | (c[3] << 8));
            }
        }
        int slot = 0;                                           // m % 3
        for (int m0 = 0; m0 < n_iter; m0 += 2) {
            const uint32_t jp = (uint32_t)((m0 >> 1) & 1);      // frame 8 + 3 m0 + i: stage (8 + i) % 6, round m0 / 2 + (8 + i) / 6
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int m = m0 + u;
                if (m >= n_iter) break;                         // block-uniform; no state is live after the loop
                const int ia = u, ib = u ^ 1;                   // static after unrolling
                uint32_t x0[L], x1[L], x2[L];
                consume_at((8 + 3 * u) % 6, jp ^ (uint32_t)(((8 + 3 * u) / 6) & 1), 3 * m + 8, x0, false);
                consume_at((9 + 3 * u) % 6, jp ^ (uint32_t)(((9 + 3 * u) / 6) & 1), 3 * m + 9, x1, false);
                consume_at((10 + 3 * u) % 6, jp ^ (uint32_t)(((10 + 3 * u) / 6) & 1), 3 * m + 10, x2, false);
                if (3 * m + 10 >= hist_from) {                  // block-uniform, last frames of the submit only
                    keep_for_next(3 * m + 8, x0);
                    keep_for_next(3 * m + 9, x1);
                    keep_for_next(3 * m + 10, x2);
                }
                uint4* rwp = rw_ring + slot * CONSUMERS;
                slot = (slot == 2) ? 0 : slot + 1;
                const uint4 e4 = *rwp;
                const uint32_t ep[L] = {e4.x, e4.y, e4.z, e4.w};
                uint32_t np[L];
                uint32_t acc0 = 0u, acc1 = 0u, acc2 = 0u;
#pragma unroll
                for (int q = 0; q < L; ++q) {
                    // middle four of A U B
                    const uint32_t pp = vmax2(st[ia][0][q], st[ib][0][q]);
                    const uint32_t qq = vmin2(st[ia][2][q], st[ib][2][q]);
                    const uint32_t uu = vmin2(st[ia][1][q], st[ib][1][q]);
                    const uint32_t vv = fa.sub(fa.add(st[ia][1][q], st[ib][1][q]), uu);
                    const uint32_t m2 = vmin2(pp, uu), m5 = vmax2(qq, vv);
                    const uint32_t gg = vmax2(pp, uu), hh = vmin2(qq, vv);
                    const uint32_t m3 = vmin2(gg, hh);
                    const uint32_t m4 = fa.sub(fa.add(gg, hh), m3);
                    const uint32_t e0 = ep[q] & 0x00FF00FFu, e1 = (ep[q] >> 8) & 0x00FF00FFu;
                    Sorted3 y = sort3_lanes(e0, e1, x0[q], fa);             // extras of window t
                    acc0 += fg_flag(x0[q], select4of7(m2, m3, m4, m5, y)) << q;
                    y = sort3_lanes(e1, x0[q], x1[q], fa);                  // extras of window t+1
                    acc1 += fg_flag(x1[q], select4of7(m2, m3, m4, m5, y)) << q;
                    y = sort3_lanes(x0[q], x1[q], x2[q], fa);               // extras of window t+2 = the new triple
                    acc2 += fg_flag(x2[q], select4of7(m2, m3, m4, m5, y)) << q;
                    st[ia][0][q] = y.lo; st[ia][1][q] = y.mid; st[ia][2][q] = y.hi;   // A is dead: the next B
                    np[q] = fa.add(x1[q], x2[q] << 8);
                }
                *rwp = make_uint4(np[0], np[1], np[2], np[3]);
                emit_acc(3 * m, acc0);
                emit_acc(3 * m + 1, acc1);
                emit_acc(3 * m + 2, acc2);
            }
        }
    } else {
        // ---- generic ring: one output per step, time loop unrolled by N ----
        uint32_t ring[N][L];
#pragma unroll
        for (int s = 0; s < N - 1; ++s) {
            consume(s, ring[s], true);
            if (short_tail) keep_for_next(s, ring[s]);
        }
        for (int base = 0; base < n_out; base += N) {
#pragma unroll
            for (int ph = 0; ph < N; ++ph) {
                const int k = base + ph;                        // output frame t_start + k
                if (k >= n_out) break;                          // block-uniform
                {
                    const int slot = (N - 1 + ph) % N;          // static after unrolling
                    consume(k + N - 1, ring[slot], false);
                    if (k + N - 1 >= hist_from) keep_for_next(k + N - 1, ring[slot]);
                    uint32_t med[L];
#pragma unroll
                    for (int q = 0; q < L; ++q) {
                        uint32_t v[N];
#pragma unroll
                        for (int s = 0; s < N; ++s) v[s] = ring[s][q];
                        med[q] = median_lanes_fma<N>(v, fa);
                    }
                    uint32_t acc = 0u;
#pragma unroll
                    for (int q = 0; q < L; ++q) acc += fg_flag(ring[slot][q], med[q]) << q;
                    emit_acc(k, acc);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// v3, N = 9 only: the same three-outputs-per-iteration median (see "shared-core sliding medians"),
// with the loop's non-arithmetic instructions cut down.  ncu on v2 at 4K: 19.6 instructions per pixel
// and frame, of which ~13 are arithmetic; the kernel is bound by instruction issue and the ALU pipe,
// not by HBM.  What changed:
//  * a pipeline stage holds the THREE frames of one loop iteration (one mbarrier wait, one release and
//    one producer round per iteration instead of three), three stages deep; the stage index is a
//    run-time offset, so the loop is unrolled by two (triple roles swap) instead of by the stage count;
//  * the three frames are converted back to back: the gray weights are materialised once per iteration;
//  * the steady-state loop carries no carried-history code: the iterations that see one of the
//    submit's last eight frames run a second copy of the body;
//  * output bytes come from one multiply and one shift (the byte is stored with STG.U8);
//  * column blocks are balanced over the last wave: the blocks that would run alone at the end of the
//    grid are cut into short temporal sub-chunks (see launch_n9).
// ------------------------------------------------------------------------------------
template <int V>
struct IntC { static constexpr int value = V; };

template <int C, int OCC, bool FMA_MERGE>
__global__ void __launch_bounds__(V2_THREADS, OCC)
k_fg_n9(FrameSrc src, int T, int Ts, int h, int wa, int thresh, uint32_t one, uint8_t* __restrict__ raw_bits,
        const __grid_constant__ CUtensorMap tmap, int tile_rows, int n_long_blocks, int ts_tail) {
    constexpr int N = 9, L = 4, PPT = 8;
    constexpr int TB = PPT * C;                   // bytes per thread per frame
    constexpr int FRAME_BYTES = CONSUMERS * TB;
    constexpr int STAGE_BYTES = 3 * FRAME_BYTES;  // one loop iteration
    constexpr int S = 3;
    constexpr int NWORDS = TB / 4;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
    uint64_t* empty = full + S;
    uint4* rw_ring = reinterpret_cast<uint4*>(smem + S * STAGE_BYTES + 2 * S * 8 + 16) + threadIdx.x;   // [3][CONSUMERS]

    const int tid = threadIdx.x;
    const int gpr = wa / PPT;
    const int G = h * gpr;
    const int cta_groups = tile_rows > 0 ? tile_rows * gpr : CONSUMERS;
    // block -> (column block, temporal sub-chunk).  The first n_long_blocks column blocks walk sub-chunks of
    // Ts frames; the remaining ones (the partial last wave of the grid) are cut into sub-chunks of ts_tail.
    int cb, t_start, t_len;
    {
        const int per_long = (T + Ts - 1) / Ts;
        const int long_ctas = n_long_blocks * per_long;
        if ((int)blockIdx.x < long_ctas) {
            cb = blockIdx.x / per_long;
            t_start = (blockIdx.x - cb * per_long) * Ts;
            t_len = Ts;
        } else {
            const int per_tail = (T + ts_tail - 1) / ts_tail;
            const int i = blockIdx.x - long_ctas;
            cb = n_long_blocks + i / per_tail;
            t_start = (i - (i / per_tail) * per_tail) * ts_tail;
            t_len = ts_tail;
        }
    }
    const int g0 = cb * cta_groups;
    const int t_end = min(T, t_start + t_len);
    const int n_out = t_end - t_start;
    if (n_out <= 0) return;                       // block-uniform
    const int n_iter = (n_out + 2) / 3;
    const int n_groups = n_iter + 3;              // group j = pipeline frames 3j-1, 3j, 3j+1 (group 0: frames 0, 1)
    const int j_first = t_start - (N - 1);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= CONSUMERS) {
        // ===== producer warp: one round per loop iteration =====
        const int lane = tid - CONSUMERS;
        const int ngroups = min(cta_groups, G - g0);
        const uint32_t bytes = (uint32_t)(ngroups * TB);
        const bool contiguous = (src.pitch == (long long)gpr * TB);
        if ((contiguous || tile_rows > 0) && lane != 0) return;
        const int r0 = g0 / gpr, c0 = g0 - r0 * gpr;
        const int len0 = min(gpr - c0, ngroups);
        const int nseg = contiguous ? 1 : 1 + (ngroups - len0 + gpr - 1) / gpr;
        for (int jg = 0; jg < n_groups; ++jg) {
            const int st = jg % S;
            if (jg >= S) mbar_wait(&empty[st], (uint32_t)((jg / S - 1) & 1));
            // bytes this round will deliver
            uint32_t expect = 0;
            for (int f = (jg == 0 ? 1 : 0); f < 3; ++f) {
                const int p = 3 * jg - 1 + f;
                const int j = min(j_first + p, T - 1);
                if (j < 0 && src.hist_valid) expect += (uint32_t)(ngroups * PPT);
                else expect += tile_rows > 0 ? (uint32_t)(tile_rows * gpr * TB) : bytes;
            }
            if (lane == 0) mbar_arrive_expect_tx(&full[st], expect);
            __syncwarp();                          // the expectation is posted before any copy can complete
            for (int f = (jg == 0 ? 1 : 0); f < 3; ++f) {
                const int p = 3 * jg - 1 + f;
                int j = min(j_first + p, T - 1);
                uint8_t* dst = smem + st * STAGE_BYTES + f * FRAME_BYTES;
                if (j < 0 && src.hist_valid) {    // carried history: compact gray frames
                    if (lane == 0) {
                        const uint8_t* hf = src.hist + (long long)(j + (N - 1)) * h * wa;
                        bulk_g2s(dst, hf + (long long)g0 * PPT, (uint32_t)(ngroups * PPT), &full[st]);
                    }
                    continue;
                }
                if (j < -src.n_inline_halo) j = -src.n_inline_halo;   // replicate the earliest frame
                const uint8_t* fr = src.cur + (long long)j * src.frame_stride;
                if (tile_rows > 0) {
                    tma_load_box(dst, &tmap, cb * tile_rows, j + src.n_inline_halo, &full[st]);
                } else if (contiguous) {
                    bulk_g2s(dst, fr + (long long)g0 * TB, bytes, &full[st]);
                } else {
                    for (int i = lane; i < nseg; i += 32) {
                        const int off = (i == 0) ? 0 : len0 + (i - 1) * gpr;
                        const int n = (i == 0) ? len0 : min(gpr, ngroups - off);
                        bulk_g2s(dst + off * TB, fr + (long long)(r0 + i) * src.pitch + (long long)(i == 0 ? c0 : 0) * TB,
                                 (uint32_t)(n * TB), &full[st]);
                    }
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int g = g0 + tid;
    const bool active = g < G && tid < cta_groups;
    const int row = active ? g / gpr : 0;
    const int col = active ? g - row * gpr : 0;
    const uint32_t neg_th = ((uint32_t)(-thresh) & 0xFFFFu) * 0x00010001u;
    FmaAdd fa;
    fa.one = one;
    fa.mone = 0u - one;
    const uint32_t k1001 = one * 0x1001u;
    uint8_t* out = raw_bits + ((long long)t_start * h + row) * gpr + col;    // one byte per thread and frame
    const uint32_t out_step = (uint32_t)h * (uint32_t)gpr;
    const bool lane0 = (tid & 31) == 0;
    const int hist_to = (T - 1) - j_first;
    const int hist_from = (src.hist_out != nullptr && t_end == T) ? hist_to - (N - 2) : 0x7FFFFFFF;
    const bool hist_src = (C == 3) && src.hist_valid && j_first < 0;         // some warm-up frames are carried gray frames

    // pipeline frame p (slot f of the stage at `sp`) as packed gray lanes
    auto take = [&](const uint8_t* sp, int f, int p, uint32_t (&dst)[L], bool maybe_hist) {
        if (maybe_hist && hist_src && j_first + p < 0) {                    // block-uniform
            uint32_t gw[L / 2];
            const uint32_t* sg = reinterpret_cast<const uint32_t*>(sp + f * FRAME_BYTES + tid * PPT);
#pragma unroll
            for (int i = 0; i < L / 2; ++i) gw[i] = sg[i];
            gray_to_lanes<L>(gw, dst);
        } else {
            uint32_t w[NWORDS];
            const uint2* sp2 = reinterpret_cast<const uint2*>(sp + f * FRAME_BYTES + tid * TB);
#pragma unroll
            for (int i = 0; i < TB / 8; ++i) {
                const uint2 v = sp2[i];
                w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
            if constexpr (C == 3) bgr_to_lanes_dp<L>(w, dst);
            else gray_to_lanes<L>(w, dst);
        }
    };
    auto keep_for_next = [&](int p, const uint32_t (&v)[L]) {
        if (p >= hist_from && p <= hist_to && active) {
            uint32_t hw[L / 2];
            lanes_to_gray<L>(v, hw);
            uint32_t* hp = reinterpret_cast<uint32_t*>(
                src.hist_out + (((long long)(p - hist_from) * h + row) * gpr + col) * PPT);
#pragma unroll
            for (int i = 0; i < L / 2; ++i) hp[i] = hw[i];
        }
    };
    auto fg_flag = [&](uint32_t x, uint32_t med) -> uint32_t {
        return __viaddmin_s16x2_relu(__vabsdiffu4(x, med), neg_th, 0x00010001u);
    };
    // acc = sum over lanes q of flag << q: bits 0..3 = pixels 0..3, bits 16..19 = pixels 4..7
    // `ok`: this thread owns pixels (and, in the last iteration of a sub-chunk, the output frame exists);
    // a predicated store, no branch
    auto emit_acc = [&](uint8_t* o, uint32_t acc, bool ok) {
        uint32_t y;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(acc), "r"(k1001), "r"(0u));   // acc | acc << 12 (disjoint bits)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u8 [%0], %1;\n\t}" ::"l"(o),
                     "r"(y >> 12), "r"((uint32_t)ok)
                     : "memory");
    };
    // mbarriers by shared-window address (computed once: the generic -> shared conversion reads a special register)
    const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
    auto wait_full = [&](int st, uint32_t parity) {
        const uint32_t addr = full_a + 8u * st;
        uint32_t done = 0;
        int spins = 0;
        const long long t_wait0 = clock64();
        while (true) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(addr), "r"(parity), "r"(MBAR_SUSPEND_NS)
                : "memory");
            if (done) break;
            // never hang the GPU on a pipeline bug: give up after ~2 s of SM clock (a try_wait may suspend for up to 1 ms)
        if ((++spins & 63) == 0 && clock64() - t_wait0 > (1ll << 32)) __trap();
        }
    };
    auto release = [&](int st, const uint32_t (&a)[L], const uint32_t (&b)[L], const uint32_t (&c)[L]) {
        // only after the loaded words have been consumed (the lanes depend on every LDS)
        asm volatile("" ::"r"(a[0]), "r"(a[L - 1]), "r"(b[0]), "r"(b[L - 1]), "r"(c[0]), "r"(c[L - 1]) : "memory");
        // elect.sync converges the warp (every lane has consumed its words by then) and picks the lane that arrives
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "elect.sync _|p, 0xffffffff;\n\t"
            "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(empty_a + 8u * st)
            : "memory");
    };

    const uint8_t* stage0 = smem;
    uint32_t st[2][3][L];         // sorted triples: st[m & 1] = A, st[(m + 1) & 1] = B
    {
        // ---- prologue: groups 0 (frames 0, 1), 1 (frames 2..4), 2 (frames 5..7)
        const bool short_tail = hist_from < N - 1;
        uint32_t a[L], b[L], c[L];
        wait_full(0, 0u);
        take(stage0, 1, 0, a, true);
        take(stage0, 2, 1, b, true);
        release(0, a, b, b);
        if (short_tail) { keep_for_next(0, a); keep_for_next(1, b); }
        rw_ring[0] = make_uint4(a[0] | (b[0] << 8), a[1] | (b[1] << 8), a[2] | (b[2] << 8), a[3] | (b[3] << 8));
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint8_t* sp = stage0 + (1 + i) * STAGE_BYTES;
            wait_full(1 + i, 0u);
            take(sp, 0, 2 + 3 * i, a, true);
            take(sp, 1, 3 + 3 * i, b, true);
            take(sp, 2, 4 + 3 * i, c, true);
            release(1 + i, a, b, c);
            if (short_tail) { keep_for_next(2 + 3 * i, a); keep_for_next(3 + 3 * i, b); keep_for_next(4 + 3 * i, c); }
#pragma unroll
            for (int q = 0; q < L; ++q) {
                const Sorted3 s3 = sort3_lanes(a[q], b[q], c[q], fa);
                st[i][0][q] = s3.lo; st[i][1][q] = s3.mid; st[i][2][q] = s3.hi;
            }
            rw_ring[(1 + i) * CONSUMERS] =
                make_uint4(b[0] | (c[0] << 8), b[1] | (c[1] << 8), b[2] | (c[2] << 8), b[3] | (c[3] << 8));
        }
    }

    // iteration m: outputs t_start + 3m .. + 2 from group m + 3 (stage (m + 3) % 3 = m % 3), ring slot m % 3
    uint8_t* optr = out;          // output byte of frame t_start + 3m
    int sidx = 0;                 // m % 3: stage and ring slot
    uint32_t par = 1u;            // parity of round (m + 3) / 3
    auto iterate = [&](auto UC, auto HC, int m) {
        constexpr int ia = decltype(UC)::value, ib = ia ^ 1;
        constexpr bool HIST = decltype(HC)::value != 0;
        const uint8_t* sp = stage0 + sidx * STAGE_BYTES;
        uint32_t x0[L], x1[L], x2[L];
        wait_full(sidx, par);
        take(sp, 0, 3 * m + 8, x0, false);
        take(sp, 1, 3 * m + 9, x1, false);
        take(sp, 2, 3 * m + 10, x2, false);
        release(sidx, x0, x1, x2);
        if constexpr (HIST) {
            keep_for_next(3 * m + 8, x0);
            keep_for_next(3 * m + 9, x1);
            keep_for_next(3 * m + 10, x2);
        }
        uint4* rwp = rw_ring + sidx * CONSUMERS;
        const uint4 e4 = *rwp;
        const uint32_t ep[L] = {e4.x, e4.y, e4.z, e4.w};
        uint32_t np[L];
        uint32_t acc0 = 0u, acc1 = 0u, acc2 = 0u;
#pragma unroll
        for (int q = 0; q < L; ++q) {
            // middle four of A U B
            const uint32_t pp = vmax2(st[ia][0][q], st[ib][0][q]);
            const uint32_t qq = vmin2(st[ia][2][q], st[ib][2][q]);
            const uint32_t uu = vmin2(st[ia][1][q], st[ib][1][q]);
            const uint32_t vv = fa.sub(fa.add(st[ia][1][q], st[ib][1][q]), uu);
            const uint32_t m2 = vmin2(pp, uu), hh = vmin2(qq, vv);
            // max(a, b) = a + b - min(a, b): two multiply-adds on the FMA pipe instead of one VIMNMX on the ALU pipe
            const uint32_t gg = FMA_MERGE ? fa.sub(fa.add(pp, uu), m2) : vmax2(pp, uu);
            const uint32_t m5 = FMA_MERGE ? fa.sub(fa.add(qq, vv), hh) : vmax2(qq, vv);
            const uint32_t m3 = vmin2(gg, hh);
            const uint32_t m4 = fa.sub(fa.add(gg, hh), m3);
            // (first | second << 8) per u16 half -> two lanes: one PRMT each (bytes 0, 2 / bytes 1, 3 into the low bytes)
            const uint32_t e0 = __byte_perm(ep[q], 0u, 0x4240), e1 = __byte_perm(ep[q], 0u, 0x4341);
            Sorted3 y = sort3_lanes(e0, e1, x0[q], fa);             // extras of window t
            acc0 += fg_flag(x0[q], select4of7(m2, m3, m4, m5, y)) << q;
            y = sort3_lanes(e1, x0[q], x1[q], fa);                  // extras of window t+1
            acc1 += fg_flag(x1[q], select4of7(m2, m3, m4, m5, y)) << q;
            y = sort3_lanes(x0[q], x1[q], x2[q], fa);               // extras of window t+2 = the new triple
            acc2 += fg_flag(x2[q], select4of7(m2, m3, m4, m5, y)) << q;
            st[ia][0][q] = y.lo; st[ia][1][q] = y.mid; st[ia][2][q] = y.hi;   // A is dead: the next B
            np[q] = fa.add(x1[q], x2[q] << 8);
        }
        *rwp = make_uint4(np[0], np[1], np[2], np[3]);
        if constexpr (HIST) {                                    // the last iteration of a sub-chunk may be partial
            emit_acc(optr, acc0, active && 3 * m < n_out);
            emit_acc(optr + out_step, acc1, active && 3 * m + 1 < n_out);
            emit_acc(optr + 2 * out_step, acc2, active && 3 * m + 2 < n_out);
        } else {
            emit_acc(optr, acc0, active);
            emit_acc(optr + out_step, acc1, active);
            emit_acc(optr + 2 * out_step, acc2, active);
        }
        optr += 3 * out_step;
        if (sidx == 2) { sidx = 0; par ^= 1u; } else { ++sidx; }
    };
    // iterations that cannot see one of the submit's last eight frames: 3m + 10 < hist_from
    // ... and whose three output frames all exist
    int m_plain = hist_from == 0x7FFFFFFF ? n_iter : min(n_iter, max(0, (hist_from - 10 + 2) / 3));
    m_plain = min(m_plain, n_out / 3) & ~1;                                              // pairs: the triple roles swap every iteration
    int m = 0;
    for (; m < m_plain; m += 2) {
        iterate(IntC<0>{}, IntC<0>{}, m);
        iterate(IntC<1>{}, IntC<0>{}, m + 1);
    }
    for (; m < n_iter; m += 2) {
        iterate(IntC<0>{}, IntC<1>{}, m);
        if (m + 1 < n_iter) iterate(IntC<1>{}, IntC<1>{}, m + 1);
    }
}

// Column blocks of the N = 9 kernel over the machine: with `slots` CTAs resident at a time, the blocks of
// the last, partial wave would run on a mostly idle GPU for as long as a full wave takes.  Those blocks are
// cut into temporal sub-chunks short enough to spread them over all slots.
template <int C, int OCC, bool FMA_MERGE>
cudaError_t launch_n9(cudaStream_t s, const FrameSrc& src, int T, const Geom& g, int thresh, uint16_t* raw_bits,
                      int gpu_share) {
    constexpr int N = 9, L = 4, TB = 2 * L * C;
    constexpr int SMEM = 3 * 3 * CONSUMERS * TB + 2 * 3 * 8 + 16 + 3 * CONSUMERS * 16;
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_fg_n9<C, OCC, FMA_MERGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
    }
    const int gpr = g.wa / (2 * L);
    const int G = g.h * gpr;
    int n_col_blocks = (G + CONSUMERS - 1) / CONSUMERS;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int tile_rows = 0;
    const long long row_bytes = (long long)gpr * TB;
    if (src.pitch != row_bytes && row_bytes <= 2048 && gpr <= CONSUMERS && tma_enabled()) {
        const int rows = CONSUMERS / gpr;
        const int frames = src.n_inline_halo + T;
        const uint8_t* base = src.cur - (long long)src.n_inline_halo * src.frame_stride;
        if (encode_roi_tensor(&tmap, base, row_bytes, g.h, frames, src.pitch, src.frame_stride, rows)) {
            tile_rows = rows;
            n_col_blocks = (g.h + rows - 1) / rows;
        }
    }
    const int Ts = tile_rows > 0 ? pick_ts_roi(T, n_col_blocks, N, gpu_share, OCC) : pick_ts(T, n_col_blocks, N, gpu_share);
    const int per_long = (T + Ts - 1) / Ts;
    // tail balancing (only when the context has the GPU to itself and nobody forced a sub-chunk length)
    int n_long = n_col_blocks, ts_tail = Ts;
    static const bool balance = [] { const char* e = getenv("SWB_K1_BALANCE"); return !(e && e[0] == '0'); }();
    if (balance && gpu_share <= 1 && g_forced_ts == 0) {
        const long long slots = 148ll * OCC;
        const long long ctas = (long long)n_col_blocks * per_long;
        const long long rem = ctas % slots;
        if (ctas > slots && rem != 0 && rem * 4 < slots * 3 && Ts >= 48) {
            const int tail_blocks = (int)(rem / per_long);       // whole column blocks of the partial wave
            if (tail_blocks > 0) {
                // enough pieces to occupy every slot about once, each at least 24 frames (warm-up: 8)
                int pieces = (int)std::min<long long>((slots + tail_blocks - 1) / tail_blocks, (long long)(Ts / 24));
                if (pieces > 1) {
                    ts_tail = ((Ts + pieces - 1) / pieces + 5) / 6 * 6;
                    n_long = n_col_blocks - tail_blocks;
                }
            }
        }
    }
    const int per_tail = (T + ts_tail - 1) / ts_tail;
    const long long grid = (long long)n_long * per_long + (long long)(n_col_blocks - n_long) * per_tail;
    k_fg_n9<C, OCC, FMA_MERGE><<<(unsigned)grid, V2_THREADS, SMEM, s>>>(src, T, Ts, g.h, g.wa, thresh, 1u,
                                                            reinterpret_cast<uint8_t*>(raw_bits), tmap, tile_rows,
                                                            n_long, ts_tail);
    return cudaGetLastError();
}

template <int N, int C, int OCC>
cudaError_t launch_v2_occ(cudaStream_t s, const FrameSrc& src, int T, const Geom& g, int thresh,
                          uint16_t* raw_bits, int gpu_share) {
    constexpr int L = (N <= 5) ? 8 : 4;
    constexpr int TB = 2 * L * C;
    constexpr int S = (N == 5) ? 4 : (N == 9 ? 6 : 8);   // grouped loops: stages = frames per unrolled body
    constexpr int SMEM = S * CONSUMERS * TB + 2 * S * 8 + (N == 9 ? 3 * CONSUMERS * 16 : 0);
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_fg_bits_v2<N, C, S, L, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
    }
    const int gpr = g.wa / (2 * L);
    const int G = g.h * gpr;
    int n_col_blocks = (G + CONSUMERS - 1) / CONSUMERS;
    // Rows of a cropped ROI are not adjacent in memory.  When a row is at most 2 KB (TMA boxes are limited to
    // 256 elements per dimension; the row is described in 8-byte elements) the CTA owns whole rows and
    // the producer fetches its tile with one tensor-map load per frame instead of one bulk copy per row.
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int tile_rows = 0;
    const long long row_bytes = (long long)gpr * TB;
    if (src.pitch != row_bytes && row_bytes <= 2048 && gpr <= CONSUMERS && tma_enabled()) {
        const int rows = CONSUMERS / gpr;
        const int frames = src.n_inline_halo + T;
        const uint8_t* base = src.cur - (long long)src.n_inline_halo * src.frame_stride;
        if (encode_roi_tensor(&tmap, base, row_bytes, g.h, frames, src.pitch, src.frame_stride, rows)) {
            tile_rows = rows;
            n_col_blocks = (g.h + rows - 1) / rows;
        }
    }
    const int Ts = tile_rows > 0 ? pick_ts_roi(T, n_col_blocks, N, gpu_share, OCC) : pick_ts(T, n_col_blocks, N, gpu_share);
    dim3 grid(n_col_blocks, (T + Ts - 1) / Ts);
    k_fg_bits_v2<N, C, S, L, OCC><<<grid, V2_THREADS, SMEM, s>>>(src, T, Ts, g.h, g.wa, thresh, 1u,
                                                                reinterpret_cast<uint8_t*>(raw_bits), tmap, tile_rows);
    return cudaGetLastError();
}

template <int N, int C>
cudaError_t launch_v2(cudaStream_t s, const FrameSrc& src, int T, const Geom& g, int thresh,
                      uint16_t* raw_bits, int gpu_share) {
    if constexpr (N == 9) {
        // the N = 9 loop fits 72 registers: three CTAs (24 consumer warps) per SM
        static const bool occ2 = [] { const char* e = getenv("SWB_K1_N9_OCC"); return e && e[0] == '2'; }();
        static const bool v2 = [] { const char* e = getenv("SWB_K1_N9_V2"); return e && e[0] == '1'; }();
        static const bool fma_merge = [] { const char* e = getenv("SWB_K1_N9_FMA_MERGE"); return !(e && e[0] == '0'); }();
        if (!v2) {
            if (occ2) return launch_n9<C, 2, false>(s, src, T, g, thresh, raw_bits, gpu_share);
            return fma_merge ? launch_n9<C, 3, true>(s, src, T, g, thresh, raw_bits, gpu_share)
                             : launch_n9<C, 3, false>(s, src, T, g, thresh, raw_bits, gpu_share);
        }
        if (!occ2) return launch_v2_occ<N, C, 3>(s, src, T, g, thresh, raw_bits, gpu_share);
    }
    return launch_v2_occ<N, C, 2>(s, src, T, g, thresh, raw_bits, gpu_share);
}

template <int N>
cudaError_t launch_n(cudaStream_t s, const FrameSrc& src, int channels, int T, const Geom& g,
                     int thresh, uint16_t* raw_bits, bool aligned, int gpu_share) {
    if (aligned) {   // 16-byte aligned rows: bulk-copy pipeline (v2)
        if (channels == 3) return launch_v2<N, 3>(s, src, T, g, thresh, raw_bits, gpu_share);
        return launch_v2<N, 1>(s, src, T, g, thresh, raw_bits, gpu_share);
    }
    const int G = g.h * (g.wa >> 4);
    const int Ts = pick_ts(T, (G + 255) / 256, N, gpu_share);
    dim3 grid((G + 255) / 256, (T + Ts - 1) / Ts);
    dim3 block(256);
    // odd pitches / frame widths: guarded byte loads (v1)
    if (channels == 3) k_fg_bits<N, 3, false><<<grid, block, 0, s>>>(src, T, Ts, g.h, g.wa, thresh, raw_bits);
    else k_fg_bits<N, 1, false><<<grid, block, 0, s>>>(src, T, Ts, g.h, g.wa, thresh, raw_bits);
    return cudaGetLastError();
}

}  // namespace

int last_temporal_subchunk() { return g_last_ts; }

cudaError_t launch_fg_bits(cudaStream_t s, const FrameSrc& src, int channels, int median_n, int T,
                           const Geom& g, int thresh, uint16_t* raw_bits, bool aligned, int* n_launches,
                           int gpu_share, int forced_ts) {
    if (n_launches) *n_launches += 1;
    g_forced_ts = forced_ts;
    switch (median_n) {
        case 1: return launch_n<1>(s, src, channels, T, g, thresh, raw_bits, aligned, gpu_share);
        case 3: return launch_n<3>(s, src, channels, T, g, thresh, raw_bits, aligned, gpu_share);
        case 5: return launch_n<5>(s, src, channels, T, g, thresh, raw_bits, aligned, gpu_share);
        case 7: return launch_n<7>(s, src, channels, T, g, thresh, raw_bits, aligned, gpu_share);
        case 9: return launch_n<9>(s, src, channels, T, g, thresh, raw_bits, aligned, gpu_share);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace swb
