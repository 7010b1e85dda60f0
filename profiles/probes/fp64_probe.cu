// FP64 on this B200: throughput of independent DFMA chains, latency of a dependent DFMA chain, throughput of the
// FP64 tensor-core path (mma.sync.m8n8k4.f64).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_tp(double* out, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma_lat(double* out, int iters, double a, double b, long long* cyc) {
    double x = threadIdx.x;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) x = fma(x, a, b);
    const long long c1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = c1 - c0;
}

__global__ void k_ffma_lat(float* out, int iters, float a, float b, long long* cyc) {
    float x = threadIdx.x;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) x = fmaf(x, a, b);
    const long long c1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = c1 - c0;
}

__global__ void k_f2f_lat(double* out, int iters, long long* cyc) {
    double x = 1.0 + threadIdx.x * 1e-3;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {            // double -> float -> double round trip, dependent
        float f = (float)x;
        asm volatile("" : "+f"(f));
        x = (double)f;
        asm volatile("" : "+d"(x));
    }
    const long long c1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = c1 - c0;
}

__global__ void k_rsq_lat(float* out, int iters, long long* cyc) {
    float x = 1.0f + threadIdx.x * 1e-3f;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        x = rsqrtf(x);
        asm volatile("" : "+f"(x));
    }
    const long long c1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = c1 - c0;
}

__global__ void k_lds_lat(int* out, int iters, long long* cyc) {
    __shared__ int idx[256];
    for (int i = threadIdx.x; i < 256; i += 32) idx[i] = (i * 7 + 3) & 255;
    __syncwarp();
    int x = threadIdx.x;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) x = idx[x];
    const long long c1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = c1 - c0;
}

__global__ void k_bar_lat(int* out, int iters, long long* cyc) {
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) __syncthreads();
    const long long c1 = clock64();
    if (threadIdx.x == 0) { *cyc = c1 - c0; out[0] = 1; }
}

__global__ void k_dmma_tp(double* out, int iters, double a, double b) {
    double c[4][2];
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma_lat(double* out, int iters, double a, double b, long long* cyc) {
    double c0 = threadIdx.x, c1 = 1.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1)
                     : "d"(a), "d"(b));
    const long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    cudaMallocManaged(&cyc, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = warps * 32;                      // one CTA per SM
        k_dfma_tp<<<148, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dfma_tp<<<148, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * threads * 8.0 * iters;
        printf("DFMA throughput, %2d warps/SM: %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM (at 1.965 GHz)\n", warps, ms,
               2 * fma / ms * 1e-9, fma / 148 / (ms * 1e-3 * 1.965e9));
        k_dmma_tp<<<148, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dmma_tp<<<148, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double mfma = 148.0 * warps * 4.0 * iters * 256.0;
        printf("DMMA m8n8k4 throughput, %2d warps/SM: %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM\n", warps, ms,
               2 * mfma / ms * 1e-9, mfma / 148 / (ms * 1e-3 * 1.965e9));
    }
    k_dfma_lat<<<1, 32>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    printf("DFMA dependent latency: %.1f cycles\n", (double)*cyc / iters);
    k_ffma_lat<<<1, 32>>>((float*)out, iters, 1.0000001f, 1e-9f, cyc);
    cudaDeviceSynchronize();
    printf("FFMA dependent latency: %.1f cycles\n", (double)*cyc / iters);
    k_dmma_lat<<<1, 32>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    printf("DMMA dependent latency: %.1f cycles\n", (double)*cyc / iters);
    k_f2f_lat<<<1, 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    printf("F2F double->float->double round trip, dependent: %.1f cycles\n", (double)*cyc / iters);
    k_rsq_lat<<<1, 32>>>((float*)out, iters, cyc);
    cudaDeviceSynchronize();
    printf("MUFU.RSQ dependent latency: %.1f cycles\n", (double)*cyc / iters);
    k_lds_lat<<<1, 32>>>((int*)out, iters, cyc);
    cudaDeviceSynchronize();
    printf("LDS dependent (pointer chase) latency: %.1f cycles\n", (double)*cyc / iters);
    k_bar_lat<<<1, 256>>>((int*)out, iters, cyc);
    cudaDeviceSynchronize();
    printf("__syncthreads, 8 warps, back to back: %.1f cycles\n", (double)*cyc / iters);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
