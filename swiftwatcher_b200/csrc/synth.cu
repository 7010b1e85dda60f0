// Seeded synthetic chimney-swift video, the CUDA twin of oracle/synth.py (bench and
// test input; the reference ships no sample video).  Every pixel is a pure function
// of (seed, video, t, y, x) through a stateless 32-bit hash, so device-generated
// frames equal the numpy oracle's bit for bit.
#include "swb_internal.cuh"

namespace swb {

namespace {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

constexpr uint32_t GOLD = 0x9E3779B1u;
constexpr uint32_t K_T = 0x85EBCA6Bu;
constexpr uint32_t BIRD_TAG = 0xB1D50000u;

__device__ __forceinline__ int noise_c(uint32_t hsh, int c) {
    return (int)((((hsh >> (8 * c)) & 0xFFu) * 5u) >> 8) - 2;
}

__global__ void k_background(uint8_t* __restrict__ dst, uint32_t vkey, int t0, int n, int h, int w) {
    const long long npx = (long long)h * w;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (i >= npx) return;
    const uint32_t fk = mix32(vkey ^ ((uint32_t)(t0 + f) * K_T));
    const uint32_t hsh = mix32(fk + (uint32_t)i * GOLD);
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    const int grad = (64 * x) / w - (48 * y) / h;
    uint8_t* p = dst + ((long long)f * npx + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = (uint8_t)(140 + 10 * c + grad + noise_c(hsh, c));
}

// one CTA per (bird, frame): paint the footprint with the dark value (idempotent)
__global__ void k_birds(uint8_t* __restrict__ dst, uint32_t vkey, int t0, int n_birds, int h, int w) {
    const int b = blockIdx.x, f = blockIdx.y;
    uint32_t hk[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) hk[k] = mix32(vkey ^ (BIRD_TAG + (uint32_t)b * 8u + (uint32_t)k));
    const long long x0 = hk[0] % (uint32_t)w, y0 = hk[1] % (uint32_t)h;
    long long vx16 = 48 + hk[2] % 80u;
    if (hk[2] & 0x80000000u) vx16 = -vx16;
    const long long vy16 = (long long)(((int)(hk[3] % 129u) - 64) | 1);
    const int bw = 5 + (int)(hk[4] % 8u), bh = 7 + (int)(hk[5] % 10u);
    const int ell = (int)(hk[6] & 1u);
    const long long t = (long long)t0 + f;
    long long cx = ((x0 * 16 + vx16 * t) >> 4) % w;
    long long cy = ((y0 * 16 + vy16 * t) >> 4) % h;
    if (cx < 0) cx += w;
    if (cy < 0) cy += h;
    const uint32_t fk = mix32(vkey ^ ((uint32_t)t * K_T));
    const long long npx = (long long)h * w;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
        const int dy = i / bw, dx = i - dy * bw;
        if (ell) {
            const long long ex = 2 * dx + 1 - bw, ey = 2 * dy + 1 - bh;
            if (ex * ex * bh * bh + ey * ey * bw * bw > (long long)bw * bw * bh * bh) continue;
        }
        const int y = (int)((cy + dy) % h), x = (int)((cx + dx) % w);
        const long long pi = (long long)y * w + x;
        const uint32_t hsh = mix32(fk + (uint32_t)pi * GOLD);
        uint8_t* p = dst + ((long long)f * npx + pi) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) p[c] = (uint8_t)(40 + 5 * c + noise_c(hsh, c));
    }
}

}  // namespace

cudaError_t launch_synth(cudaStream_t s, uint8_t* dst, uint32_t seed, uint32_t video, int t0, int n, int h, int w,
                         int n_birds) {
    const uint32_t vkey = mix32(seed * GOLD + video);
    const long long npx = (long long)h * w;
    dim3 grid((unsigned)((npx + 255) / 256), n);
    k_background<<<grid, 256, 0, s>>>(dst, vkey, t0, n, h, w);
    if (n_birds > 0) {
        dim3 gb(n_birds, n);
        k_birds<<<gb, 64, 0, s>>>(dst, vkey, t0, n_birds, h, w);
    }
    return cudaGetLastError();
}

}  // namespace swb
