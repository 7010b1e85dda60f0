// Glue kernels of the batched segment classifier (segment_classification.py: WindowedSqueezeNet.forward_buffered).
// The classifier's convolutions are library kernels (cuDNN / CUTLASS, as the reference's own SqueezeNet); what sits
// between them — writing a layer's (ReLU'd) output window into the next layer's halo'd patch buffer, and the 3x3 / 2
// max-pooling of such a patch — was ~55 % of the classifier's GPU time as strided PyTorch elementwise kernels
// (profiles/classifier_timing.py).  Both are plain NHWC (channels-last) float32 streaming kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "../../include/swb200.h"

namespace {

// dst[b][oy + y][ox + x][c] = relu?(src[b][y][x][c] + bias?[c]); a row of the window is w * C contiguous floats on both
// sides.  src == dst with identical geometry is allowed (in-place bias + ReLU: every element is read and written by
// the same thread), hence no __restrict__.
template <typename V>
__global__ void __launch_bounds__(256)
k_nhwc_paste(const V* src, V* dst, int h, int row_v, long long src_img_v, long long dst_img_v, int dst_row_v,
             long long dst_off_v, int relu, const V* __restrict__ bias, int c_v) {
    // one thread per vector of the (contiguous) source image: windows are small (5..36 rows of 10..1500 vectors), a block
    // per row would leave most of its threads idle
    const long long b = blockIdx.y;
    const V* s = src + b * src_img_v;
    V* dimg = dst + b * dst_img_v + dst_off_v;
    const int n = h * row_v;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / row_v, x = i - y * row_v;
        V* d = dimg + (long long)y * dst_row_v;
        V v = s[i];
        if (bias != nullptr) {
            const V bv = bias[x % c_v];
            if constexpr (sizeof(V) == 16) {
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            } else {
                v += bv;
            }
        }
        if (relu) {
            if constexpr (sizeof(V) == 16) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            } else {
                v = fmaxf(v, 0.f);
            }
        }
        d[x] = v;
    }
}

// out[b][oy][ox][c] = max over the k x k window at (oy * s, ox * s) of in[b][.][.][c]; in is H x W, out is OH x OW
// (the patch was built so that every window lies inside it: rows / columns outside the feature map hold -inf)
__global__ void __launch_bounds__(256)
k_nhwc_maxpool(const float4* __restrict__ in, float4* __restrict__ out, int W, int C4, int OH, int OW, int k, int s,
               long long in_img, long long out_img) {
    const long long b = blockIdx.z;
    const int oy = blockIdx.y;
    const int n = OW * C4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int ox = i / C4, c = i - ox * C4;
        const float4* p = in + b * in_img + ((long long)(oy * s) * W + ox * s) * C4 + c;
        float4 m = p[0];
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
                const float4 v = p[((long long)dy * W + dx) * C4];
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        out[b * out_img + (long long)oy * n + i] = m;
    }
    (void)OH;
}

}  // namespace

extern "C" {

int swb_nhwc_paste(const float* src, float* dst, int64_t batch, int32_t channels, int32_t h, int32_t w, int32_t dst_h,
                   int32_t dst_w, int32_t off_y, int32_t off_x, const float* bias, int32_t relu, void* stream) {
    if (!src || !dst || batch < 0 || channels < 1 || h < 1 || w < 1 || off_y < 0 || off_x < 0 || off_y + h > dst_h ||
        off_x + w > dst_w || batch > 65535 || (long long)h * w * channels > 0x7FFFFFFF)
        return SWB_ERR_INVALID;
    if (batch == 0) return SWB_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long row = (long long)w * channels, dst_row = (long long)dst_w * channels;
    const long long dst_off = ((long long)off_y * dst_w + off_x) * channels;
    const bool v4 = channels % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
    if (v4) {
        const int row_v = (int)(row / 4);
        dim3 grid((unsigned)std::min<long long>(((long long)h * row_v + 255) / 256, 64), (unsigned)batch);
        k_nhwc_paste<float4><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), h, row_v,
                                                  (long long)h * row / 4, (long long)dst_h * dst_row / 4, (int)(dst_row / 4),
                                                  dst_off / 4, relu, reinterpret_cast<const float4*>(bias), channels / 4);
    } else {
        dim3 grid((unsigned)std::min<long long>((h * row + 255) / 256, 64), (unsigned)batch);
        k_nhwc_paste<float><<<grid, 256, 0, s>>>(src, dst, h, (int)row, (long long)h * row, (long long)dst_h * dst_row,
                                                 (int)dst_row, dst_off, relu, bias, channels);
    }
    return cudaGetLastError() == cudaSuccess ? SWB_OK : SWB_ERR_CUDA;
}

int swb_nhwc_maxpool(const float* in, float* out, int64_t batch, int32_t channels, int32_t in_h, int32_t in_w,
                     int32_t kernel, int32_t stride, void* stream) {
    if (!in || !out || batch < 0 || channels < 4 || channels % 4 || kernel < 1 || stride < 1 || in_h < kernel ||
        in_w < kernel || batch > 65535 || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return SWB_ERR_INVALID;
    if (batch == 0) return SWB_OK;
    const int oh = (in_h - kernel) / stride + 1, ow = (in_w - kernel) / stride + 1;
    if (oh > 65535) return SWB_ERR_INVALID;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int c4 = channels / 4;
    dim3 grid((ow * c4 + 255) / 256, oh, (unsigned)batch);
    k_nhwc_maxpool<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), in_w, c4, oh, ow,
                                        kernel, stride, (long long)in_h * in_w * c4, (long long)oh * ow * c4);
    return cudaGetLastError() == cudaSuccess ? SWB_OK : SWB_ERR_CUDA;
}

}  // extern "C"
