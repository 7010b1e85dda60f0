// Tracker cost matrix: formulate_cost_matrix, segment_tracking.py:46-102 with
// calculate_distance_cost (:190-198), calculate_angle_cost (:201-243), calculate_nonmatch_cost (:246-250).
//
// The reference fills the match block with a Python double loop — 250,000 iterations with a scipy call and
// two atan2 per pair at 500 segments per frame; here one thread per matrix element writes the WHOLE
// (n_prev + n_curr)^2 float64 matrix the Hungarian step consumes: the match block
//     cost[i, n_prev + j] = 0.5 * 2^(dist(i, j) - 25) + 0.5 * angle_cost(i, j)
// with the reference's formulas in the reference's operation order, the diagonal 1 and everything else
// 1 + DBL_EPSILON (intialize_cost_matrix, :179-187).  Float64 throughout: sqrt is correctly rounded, CUDA's
// atan2 / exp2 are within 2 ulp of libm's, so the matrix agrees with the reference to ~1e-15 relative
// (tests state 1e-12) and the assignments are the same.  The assignment problem itself stays on the host
// (scipy.optimize.linear_sum_assignment, :253-260): the matrix lands in page-locked memory owned by the
// workspace, one DMA transfer.
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "swb_internal.cuh"

struct swb_tracker {
    int device = 0;
    int cap = 0;                    // max n_prev + n_curr
    cudaStream_t stream = nullptr;
    double* d_in = nullptr;         // [3 * cap * 2]: prev_yx, first_yx, curr_yx
    uint8_t* d_has = nullptr;       // [cap]
    double* d_cost = nullptr;       // [cap * cap]
    double* h_in = nullptr;         // pinned staging of the inputs
    uint8_t* h_has = nullptr;
    double* h_cost = nullptr;       // pinned result
    int64_t launches = 0;
    std::string error;
};

namespace {

thread_local std::string g_track_error;

int tfail(swb_tracker* t, int code, const char* fmt, ...) {
    char buf[384];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (t) t->error = buf;
    g_track_error = buf;
    return code;
}

#define TCU(t, call)                                                                                        \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) return tfail(t, SWB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

__global__ void __launch_bounds__(256)
k_track_costs(const double* __restrict__ prev_yx, const double* __restrict__ first_yx,
              const uint8_t* __restrict__ has_history, const double* __restrict__ curr_yx, int n_prev, int n_curr,
              double* __restrict__ cost) {
    const int n = n_prev + n_curr;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;     // column
    const int i = blockIdx.y;                                // row
    if (j >= n) return;
    double v = 1.0 + DBL_EPSILON;                            // "impossible" cells (intialize_cost_matrix)
    if (i == j) {
        v = 1.0;                                             // calculate_nonmatch_cost
    } else if (i < n_prev && j >= n_prev) {
        const double py = prev_yx[2 * i], px = prev_yx[2 * i + 1];
        const double cy = curr_yx[2 * (j - n_prev)], cx = curr_yx[2 * (j - n_prev) + 1];
        const double dy = py - cy, dx = px - cx;             // prev_pos - curr_pos
        const double dist = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx)));
        const double d_cost = exp2(dist - 25.0);
        double a_cost = 1.0;
        if (has_history[i]) {
            const double k = 180.0 / 3.14159265358979323846;                        // math.degrees
            const double oy = first_yx[2 * i] - py, ox = first_yx[2 * i + 1] - px;  // initial_pos - prev_pos
            const double old_angle = __dmul_rn(atan2(oy, -1.0 * ox), k);
            const double new_angle = __dmul_rn(atan2(dy, -1.0 * dx), k);
            double diff = fabs(new_angle - old_angle);
            diff = fmin(diff, 360.0 - diff);
            a_cost = exp2(diff - 90.0);
        }
        v = __dadd_rn(__dmul_rn(0.5, d_cost), __dmul_rn(0.5, a_cost));
    }
    cost[(long long)i * n + j] = v;
}

}  // namespace

extern "C" {

const char* swb_tracker_last_error(const swb_tracker* t) { return t ? t->error.c_str() : g_track_error.c_str(); }

int swb_tracker_create(int32_t device, int32_t max_segments, swb_tracker** out) {
    if (!out || max_segments <= 0 || max_segments > 16384) return tfail(nullptr, SWB_ERR_INVALID, "bad argument");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        return tfail(nullptr, SWB_ERR_CUDA, "no CUDA device available; libswb200 has no CPU path");
    if (device < 0 || device >= n) return tfail(nullptr, SWB_ERR_INVALID, "device %d out of range", device);
    swb_tracker* t = new swb_tracker();
    t->device = device;
    t->cap = max_segments;
    const size_t cap = (size_t)max_segments;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_in), 6 * cap * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_has), cap);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_cost), cap * cap * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->h_in), 6 * cap * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->h_has), cap);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->h_cost), cap * cap * sizeof(double));
    if (e != cudaSuccess) {
        const int rc = tfail(nullptr, SWB_ERR_CUDA, "swb_tracker_create: %s", cudaGetErrorString(e));
        swb_tracker_destroy(t);
        return rc;
    }
    *out = t;
    return SWB_OK;
}

int swb_tracker_destroy(swb_tracker* t) {
    if (!t) return SWB_OK;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    cudaFree(t->d_in);
    cudaFree(t->d_has);
    cudaFree(t->d_cost);
    if (t->h_in) cudaFreeHost(t->h_in);
    if (t->h_has) cudaFreeHost(t->h_has);
    if (t->h_cost) cudaFreeHost(t->h_cost);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
    return SWB_OK;
}

int swb_tracker_costs(swb_tracker* t, const double* prev_yx, const double* first_yx, const uint8_t* has_history,
                      int32_t n_prev, const double* curr_yx, int32_t n_curr, double** matrix) {
    if (!t || !matrix) return tfail(t, SWB_ERR_INVALID, "null argument");
    *matrix = nullptr;
    if (n_prev < 0 || n_curr < 0) return tfail(t, SWB_ERR_INVALID, "negative segment count");
    const int n = n_prev + n_curr;
    if (n > t->cap) return tfail(t, SWB_ERR_CAPACITY, "%d segments in two frames exceed the tracker's max_segments (%d)", n, t->cap);
    if ((n_prev > 0 && (!prev_yx || !first_yx || !has_history)) || (n_curr > 0 && !curr_yx))
        return tfail(t, SWB_ERR_INVALID, "null centroid array");
    *matrix = t->h_cost;
    if (n == 0) return SWB_OK;
    TCU(t, cudaSetDevice(t->device));
    const size_t cap = (size_t)t->cap;
    // stage the (small) inputs in page-locked memory: one DMA each, nothing pageable on the stream
    for (int i = 0; i < 2 * n_prev; ++i) { t->h_in[i] = prev_yx[i]; t->h_in[2 * cap + i] = first_yx[i]; }
    for (int i = 0; i < n_prev; ++i) t->h_has[i] = has_history[i];
    for (int i = 0; i < 2 * n_curr; ++i) t->h_in[4 * cap + i] = curr_yx[i];
    TCU(t, cudaMemcpyAsync(t->d_in, t->h_in, 6 * cap * sizeof(double), cudaMemcpyHostToDevice, t->stream));
    if (n_prev > 0) TCU(t, cudaMemcpyAsync(t->d_has, t->h_has, (size_t)n_prev, cudaMemcpyHostToDevice, t->stream));
    dim3 grid((n + 255) / 256, n);
    k_track_costs<<<grid, 256, 0, t->stream>>>(t->d_in, t->d_in + 2 * cap, t->d_has, t->d_in + 4 * cap, n_prev, n_curr,
                                                t->d_cost);
    TCU(t, cudaGetLastError());
    t->launches += 1;
    TCU(t, cudaMemcpyAsync(t->h_cost, t->d_cost, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
    TCU(t, cudaStreamSynchronize(t->stream));
    return SWB_OK;
}

int64_t swb_tracker_launch_count(const swb_tracker* t) { return t ? t->launches : 0; }

}  // extern "C"
