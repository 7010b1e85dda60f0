// Single-stage kernels: one reference function each (image_filtering.py), used by
// the per-function drop-in API (swb_stage_*).  Straightforward one-thread-per-pixel
// kernels; the fused path (fg_bits.cu / morph_mask.cu / ccl.cu) is the fast one.
#include "swb_internal.cuh"

namespace swb {

namespace {

// convert_grayscale, image_filtering.py:188-196 (cv2 4.13 fixed point).
__global__ void k_gray(const uint8_t* __restrict__ bgr, long long n, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    out[i] = (uint8_t)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
}

// temporal median of n (odd) frames: insertion sort in registers
__global__ void k_median(const uint8_t* __restrict__ stack, int n, long long npx, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    uint8_t v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = (k < n) ? stack[(long long)k * npx + i] : 255;
    // full sort of 9 slots; padding 255s sink to the top
#pragma unroll
    for (int a = 0; a < 9; ++a) {
#pragma unroll
        for (int b = 0; b + 1 < 9 - a; ++b) {
            uint8_t lo = min(v[b], v[b + 1]), hi = max(v[b], v[b + 1]);
            v[b] = lo;
            v[b + 1] = hi;
        }
    }
    uint8_t m = v[0];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (k == n / 2) m = v[k];
    out[i] = m;
}

__global__ void k_absdiff(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n,
                          uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = (int)a[i] - (int)b[i];
    out[i] = (uint8_t)(d < 0 ? -d : d);
}

// thresh_to_zero, image_filtering.py:310-316
__global__ void k_thresh(const uint8_t* __restrict__ in, long long n, int thresh, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t v = in[i];
    out[i] = ((int)v > thresh) ? v : (uint8_t)0;
}

// flat min / max filter, out-of-image pixels ignored (== scipy 'reflect' for odd sizes)
__global__ void k_minmax(const uint8_t* __restrict__ in, int h, int w, int se_h, int se_w, int is_max,
                         uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int ry = se_h / 2, rx = se_w / 2;
    int acc = is_max ? 0 : 255;
    for (int dy = -ry; dy <= ry; ++dy) {
        const int yy = y + dy;
        if ((unsigned)yy >= (unsigned)h) continue;
        for (int dx = -rx; dx <= rx; ++dx) {
            const int xx = x + dx;
            if ((unsigned)xx >= (unsigned)w) continue;
            const int v = in[(long long)yy * w + xx];
            acc = is_max ? max(acc, v) : min(acc, v);
        }
    }
    out[(long long)y * w + x] = (uint8_t)acc;
}

// regionprops accumulators indexed by label value - 1
template <typename LT>
__global__ void k_props(const LT* __restrict__ labels, int h, int w, swb_segment* acc, int cap) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const long long l = (long long)labels[(long long)y * w + x];
    if (l <= 0 || l > cap) return;
    swb_segment* s = acc + (l - 1);
    atomicAdd(&s->area, 1);
    atomicMin(&s->bbox[0], y);
    atomicMin(&s->bbox[1], x);
    atomicMax(&s->bbox[2], y + 1);
    atomicMax(&s->bbox[3], x + 1);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_row), (unsigned long long)y);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_col), (unsigned long long)x);
}

__global__ void k_props_init(swb_segment* acc, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    swb_segment s;
    s.frame = 0;
    s.label = i + 1;
    s.area = 0;
    s.bbox[0] = 0x7FFFFFFF;
    s.bbox[1] = 0x7FFFFFFF;
    s.bbox[2] = 0;
    s.bbox[3] = 0;
    s.reserved = 0;
    s.sum_row = 0;
    s.sum_col = 0;
    acc[i] = s;
}

inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

}  // namespace

cudaError_t launch_stage_gray(cudaStream_t s, const uint8_t* bgr, int h, int w, uint8_t* out) {
    const long long n = (long long)h * w;
    k_gray<<<nblk(n, 256), 256, 0, s>>>(bgr, n, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_median(cudaStream_t s, const uint8_t* stack, int n, int h, int w, uint8_t* out) {
    const long long npx = (long long)h * w;
    k_median<<<nblk(npx, 256), 256, 0, s>>>(stack, n, npx, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_absdiff(cudaStream_t s, const uint8_t* a, const uint8_t* b, long long n, uint8_t* out) {
    k_absdiff<<<nblk(n, 256), 256, 0, s>>>(a, b, n, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_thresh(cudaStream_t s, const uint8_t* in, long long n, int thresh, uint8_t* out) {
    k_thresh<<<nblk(n, 256), 256, 0, s>>>(in, n, thresh, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_minmax(cudaStream_t s, const uint8_t* in, int h, int w, int se_h, int se_w, int is_max,
                                uint8_t* out) {
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    k_minmax<<<grid, block, 0, s>>>(in, h, w, se_h, se_w, is_max, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_props(cudaStream_t s, const void* labels, int elem_size, int h, int w, swb_segment* acc,
                               int cap) {
    k_props_init<<<nblk(cap, 256), 256, 0, s>>>(acc, cap);
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    if (elem_size == 1) k_props<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)labels, h, w, acc, cap);
    else k_props<int32_t><<<grid, block, 0, s>>>((const int32_t*)labels, h, w, acc, cap);
    return cudaGetLastError();
}

}  // namespace swb
