// RPCA background model — the reference's own localisation step.
//
// rpca (image_filtering.py:220-253) stacks the n gray frames of a batch as the columns of
// a (P x n) matrix X (P = h*w pixels, n = 21) and splits it into low rank + sparse with the
// inexact augmented Lagrange multiplier method (:256-301); the "RPCA" image of a frame is
// clip(-E, 0, 255) truncated to uint8 (:243-245: what is darker than the background).
//
// Each IALM iteration is, per pixel row p (all float64, same operation order as the numpy code):
//     Eraw = (X - A) + (1/mu) Y                E = shrink(Eraw, lambda/mu)        (:282-283)
//     M    = (X - E) + (1/mu) Y                                                   (:284)
//     A'   = U diag(S - 1/mu) V^T  with  U S V^T = svd(M)   (svp == n always)     (:284-290)
//     Z    = (X - A') - E;   Y += mu Z;   mu *= rho;   stop when |Z|_F / |X|_F < tol
// M is tall and skinny, so the thin SVD goes through the n x n Gram matrix: G = M^T M =
// V diag(S^2) V^T and A' = M W with W = V diag((S - 1/mu) / S) V^T — two streaming passes over
// X, A, Y per iteration plus a 21 x 21 symmetric eigenproblem (cyclic Jacobi on the host):
//   pass 1  k_rpca_gram   M -> per-CTA partial Gram sums (fixed order: deterministic)
//   pass 2  k_rpca_apply  recomputes E, M; A' = M W; Z; Y; |Z|^2 partials; the uint8 image of E
// Bandwidth-bound float64 streaming (about 50 bytes per matrix element per iteration); the
// Gram products and the 441 multiply-adds per pixel row of A' = M W run from shared memory.
//
// Parity: the reference's LAPACK SVD and this Gram / Jacobi route differ in the last bits of A;
// the uint8 image only changes where -E lies within ~1e-12 of an integer.  On the golden batches
// it is bit-identical (tests/golden/rpca_*.npz, produced by the reference's own rpca()).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "swb_internal.cuh"

namespace swb {

namespace {

constexpr int RP_THREADS = 256;
constexpr int RP_NMAX = 32;

__device__ __forceinline__ double shrink(double v, double thr) {
    // np.maximum(v - thr, 0) + np.minimum(v + thr, 0)
    return fmax(v - thr, 0.0) + fmin(v + thr, 0.0);
}

// sum of squares (exact, integers) and maximum of the gray stack
__global__ void __launch_bounds__(RP_THREADS)
k_rpca_norms(const uint8_t* __restrict__ x, long long total, unsigned long long* __restrict__ sumsq,
             unsigned int* __restrict__ vmax) {
    unsigned long long s = 0;
    unsigned int m = 0;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long long)gridDim.x * blockDim.x;
    // 16 pixels per load where the stack is 16-byte aligned (it is: cudaMalloc), bytes for the tail
    const long long nvec = (reinterpret_cast<uintptr_t>(x) & 15) == 0 ? total / 16 : 0;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    for (long long i = gtid; i < nvec; i += gsize) {
        const uint4 q = __ldg(xv + i);
        const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
        unsigned int part = 0;                       // 16 squares of bytes: < 2^21
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            part = __dp4a(wds[k], wds[k], part);     // sum of the four byte squares
            m = max(m, max(max(wds[k] & 0xFFu, (wds[k] >> 8) & 0xFFu), max((wds[k] >> 16) & 0xFFu, wds[k] >> 24)));
        }
        s += part;
    }
    for (long long i = nvec * 16 + gtid; i < total; i += gsize) {
        const unsigned int v = x[i];
        s += (unsigned long long)(v * v);
        m = max(m, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, d);
        m = max(m, __shfl_down_sync(0xFFFFFFFFu, m, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sumsq, s);          // integer sums: order does not matter
        atomicMax(vmax, m);
    }
}

__global__ void __launch_bounds__(RP_THREADS)
k_rpca_init(const uint8_t* __restrict__ x, long long total, double dual_norm, double* __restrict__ A,
            double* __restrict__ Y, const RpcaState* __restrict__ st, uint8_t* __restrict__ out) {
    if (st) dual_norm = st->dual_norm;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        A[i] = 0.0;
        Y[i] = (double)x[i] / dual_norm;
        if (out) out[i] = 0;                       // what an all-black batch (no iteration at all) leaves behind
    }
}

// pass 1: partial Gram matrices.  A CTA walks tiles of RP_THREADS pixel rows: thread t computes
// M[p0 + t][0..n) into shared memory (padded rows: conflict-free column reads), then thread q
// accumulates the pair (i, j) = pairs[q] over the tile.  gpart[cta][q], q < n(n+1)/2.
__global__ void __launch_bounds__(RP_THREADS)
k_rpca_gram(const uint8_t* __restrict__ X, const double* __restrict__ A, const double* __restrict__ Y, int n,
            long long P, double inv_mu, double thr, double* __restrict__ gpart) {
    extern __shared__ double sm[];                 // [n][RP_THREADS + 1]
    constexpr int LD = RP_THREADS + 1;
    const int t = threadIdx.x;
    const int npairs = n * (n + 1) / 2;
    double acc[3] = {0.0, 0.0, 0.0};               // up to 3 pairs per thread (n <= 32: 528 pairs)
    int pi[3], pj[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int q = t + k * RP_THREADS;
        pi[k] = pj[k] = -1;
        if (q < npairs) {                          // row-major upper triangle: q -> (i, j), j >= i
            int i = 0;
            while (q >= n - i) { q -= n - i; ++i; }
            pi[k] = i;
            pj[k] = i + q;
        }
    }
    const long long ntiles = (P + RP_THREADS - 1) / RP_THREADS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * RP_THREADS + t;
        for (int i = 0; i < n; ++i) {
            double m = 0.0;
            if (p < P) {
                const long long idx = (long long)i * P + p;
                const double x = (double)X[idx];
                const double t2 = __dmul_rn(inv_mu, Y[idx]);           // numpy rounds the product, then the sum
                const double e = shrink(__dadd_rn(x - A[idx], t2), thr);
                m = __dadd_rn(x - e, t2);
            }
            sm[i * LD + t] = m;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (pi[k] < 0) continue;
            const double* a = sm + pi[k] * LD;
            const double* b = sm + pj[k] * LD;
            double s = 0.0;
#pragma unroll 8
            for (int r = 0; r < RP_THREADS; ++r) s = fma(a[r], b[r], s);
            acc[k] += s;
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (pi[k] >= 0) gpart[(long long)blockIdx.x * npairs + t + k * RP_THREADS] = acc[k];
}

// pass 1 for n = 21 (the reference's batch size, data_structures.py:120): the same tiles and the same
// shared-memory rows of M, but the products are register-tiled.  The 21 x 21 upper triangle is cut into
// 28 blocks of 3 x 3 pairs; thread (block b, slice s) accumulates its nine pairs over rows s, s + 8, ...
// of the tile: six shared-memory loads per nine multiply-adds instead of two per one (the plain kernel is
// bound by its shared-memory reads).  The eight slices of a block sit in adjacent lanes and are added
// with a fixed shuffle tree at the end, so the result is the same on every run.
__global__ void __launch_bounds__(RP_THREADS)
k_rpca_gram21(const uint8_t* __restrict__ X, const double* __restrict__ A, const double* __restrict__ Y,
              long long P, double inv_mu, double thr, double* __restrict__ gpart, const RpcaState* __restrict__ st) {
    // st != nullptr (graph loop): this is iteration 0, where A == 0 and Y == X / dual_norm are known without
    // reading them (k_rpca_init does not run in the graph: 0.7 GB less to write and 0.7 GB less to read at 1080p)
    const bool first = st != nullptr;
    double dual_norm = 1.0;
    if (st) { inv_mu = st->inv_mu; thr = st->thr; dual_norm = st->dual_norm; }
    constexpr int n = 21, NB = 7, NBLK = NB * (NB + 1) / 2, NS = 8;
    constexpr int npairs = n * (n + 1) / 2;
    extern __shared__ double sm[];                 // [n][LD]
    constexpr int LD = RP_THREADS + 8;             // 3 * LD = 8 mod 16: the two blocks of a half-warp read disjoint banks
    const int t = threadIdx.x;
    const int slice = t & (NS - 1), blk = t >> 3;
    const bool worker = blk < NBLK;
    int bi = 0, bj = 0;
    if (worker) {                                  // row-major upper triangle of blocks: blk -> (bi, bj), bj >= bi
        int q = blk;
        while (q >= NB - bi) { q -= NB - bi; ++bi; }
        bj = bi + q;
    }
    double acc[3][3];
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) acc[x][y] = 0.0;
    const double* a0 = sm + (3 * bi) * LD;
    const double* b0 = sm + (3 * bj) * LD;
    const long long ntiles = (P + RP_THREADS - 1) / RP_THREADS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * RP_THREADS + t;
#pragma unroll 3
        for (int i = 0; i < n; ++i) {
            double m = 0.0;
            if (p < P) {
                const long long idx = (long long)i * P + p;
                const double x = (double)X[idx];
                const double y = first ? __ddiv_rn(x, dual_norm) : Y[idx];
                const double a = first ? 0.0 : A[idx];
                const double t2 = __dmul_rn(inv_mu, y);                // numpy rounds the product, then the sum
                const double e = shrink(__dadd_rn(x - a, t2), thr);
                m = __dadd_rn(x - e, t2);
            }
            sm[i * LD + t] = m;
        }
        __syncthreads();
        if (worker) {
#pragma unroll 4
            for (int r = slice; r < RP_THREADS; r += NS) {
                const double a[3] = {a0[r], a0[LD + r], a0[2 * LD + r]};
                const double b[3] = {b0[r], b0[LD + r], b0[2 * LD + r]};
#pragma unroll
                for (int x = 0; x < 3; ++x)
#pragma unroll
                    for (int y = 0; y < 3; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
            }
        }
        __syncthreads();
    }
    // the eight slices of a block: adjacent lanes, fixed-order tree
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) {
            double v = acc[x][y];
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
            const int i = 3 * bi + x, j = 3 * bj + y;
            if (worker && slice == 0 && j >= i)
                gpart[(long long)blockIdx.x * npairs + i * n - (i * (i - 1)) / 2 + (j - i)] = v;
        }
}

// fixed-order sum of the per-CTA partials -> packed upper triangle
// (one warp per pair: lane l sums CTAs l, l + 32, ...; then a shuffle tree — the same order every run)
__global__ void k_rpca_gram_reduce(const double* __restrict__ gpart, int nctas, int npairs, double* __restrict__ G) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= npairs) return;
    double s = 0.0;
    for (int c = lane; c < nctas; c += 32) s += gpart[(long long)c * npairs + q];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
    if (lane == 0) G[q] = s;
}

// the same inside the graph loop: the first iteration reduces k_rpca_gram21's partials (n_first CTAs), the later
// ones the fused pass's (n_later CTAs)
__global__ void k_rpca_gram_reduce_st(const double* __restrict__ gpart, int n_first, int n_later, int npairs,
                                      double* __restrict__ G, const RpcaState* __restrict__ st) {
    const int nctas = st->itr == 0 ? n_first : n_later;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= npairs) return;
    double s = 0.0;
    for (int c = lane; c < nctas; c += 32) s += gpart[(long long)c * npairs + q];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
    if (lane == 0) G[q] = s;
}

// pass 2: A' = M W, Z, Y, |Z|^2, uint8 image of E.  Thread = pixel row; M and E of the row sit in
// shared memory columns (stride blockDim: conflict-free), W is read as a broadcast.
__global__ void __launch_bounds__(RP_THREADS)
k_rpca_apply(const uint8_t* __restrict__ X, const double* __restrict__ A, double* __restrict__ Anew,
             double* __restrict__ Y, int n, long long P, double inv_mu, double thr, double mu,
             const double* __restrict__ Wg, double* __restrict__ zpart, uint8_t* __restrict__ out) {
    extern __shared__ double sm[];
    double* sW = sm;                               // [n][n]
    double* sM = sW + n * n;                       // [n][RP_THREADS]
    double* sE = sM + n * RP_THREADS;              // [n][RP_THREADS]
    const int t = threadIdx.x;
    for (int i = t; i < n * n; i += RP_THREADS) sW[i] = Wg[i];
    __syncthreads();
    double zz = 0.0;
    const long long ntiles = (P + RP_THREADS - 1) / RP_THREADS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * RP_THREADS + t;
        if (p < P) {
            for (int i = 0; i < n; ++i) {
                const long long idx = (long long)i * P + p;
                const double x = (double)X[idx];
                const double t2 = __dmul_rn(inv_mu, Y[idx]);           // numpy rounds the product, then the sum
                const double e = shrink(__dadd_rn(x - A[idx], t2), thr);
                sE[i * RP_THREADS + t] = e;
                sM[i * RP_THREADS + t] = __dadd_rn(x - e, t2);
                // clip(-E, 0, 255).astype(uint8): truncation
                const double ne = fmin(fmax(-e, 0.0), 255.0);
                out[idx] = (uint8_t)ne;
            }
            for (int j = 0; j < n; ++j) {
                double a = 0.0;
                for (int i = 0; i < n; ++i) a = fma(sM[i * RP_THREADS + t], sW[i * n + j], a);
                const long long idx = (long long)j * P + p;
                const double z = ((double)X[idx] - a) - sE[j * RP_THREADS + t];
                Anew[idx] = a;
                Y[idx] = __dadd_rn(Y[idx], __dmul_rn(mu, z));
                zz = fma(z, z, zz);
            }
        }
    }
    // per-CTA |Z|^2 in a fixed order
    __shared__ double red[RP_THREADS / 32];
    for (int d = 16; d > 0; d >>= 1) zz += __shfl_down_sync(0xFFFFFFFFu, zz, d);
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = zz;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < RP_THREADS / 32; ++w) s += red[w];
        zpart[blockIdx.x] = s;
    }
}

// pass 2 for a compile-time batch size (the reference's queue holds 21 frames): the row's E and
// the accumulators of A' = M W stay in registers, W rows are read from shared memory as double2
// broadcasts (one load per two multiply-adds instead of two loads per multiply-add).
template <int N>
__global__ void __launch_bounds__(RP_THREADS, 2)
k_rpca_apply_n(const uint8_t* __restrict__ X, const double* __restrict__ A, double* __restrict__ Anew,
               double* __restrict__ Y, long long P, double inv_mu, double thr, double mu,
               const double* __restrict__ Wg, double* __restrict__ zpart, uint8_t* __restrict__ out) {
    constexpr int NP = (N + 1) & ~1;               // W rows padded to an even length
    __shared__ __align__(16) double sW[N * NP];
    extern __shared__ double sE[];                 // [N][RP_THREADS]: the row's E (registers hold the accumulators)
    const int t = threadIdx.x;
    for (int i = t; i < N * NP; i += RP_THREADS) {
        const int r = i / NP, c = i - r * NP;
        sW[i] = c < N ? Wg[r * N + c] : 0.0;
    }
    __syncthreads();
    double zz = 0.0;
    const long long ntiles = (P + RP_THREADS - 1) / RP_THREADS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * RP_THREADS + t;
        if (p >= P) continue;
        double acc[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[j] = 0.0;
#pragma unroll 3
        for (int i = 0; i < N; ++i) {
            const long long idx = (long long)i * P + p;
            const double x = (double)X[idx];
            const double t2 = __dmul_rn(inv_mu, Y[idx]);
            const double e = shrink(__dadd_rn(x - A[idx], t2), thr);
            sE[i * RP_THREADS + t] = e;
            const double m = __dadd_rn(x - e, t2);
            out[idx] = (uint8_t)fmin(fmax(-e, 0.0), 255.0);
            const double2* wr = reinterpret_cast<const double2*>(sW + i * NP);
#pragma unroll
            for (int j = 0; j < NP / 2; ++j) {
                const double2 w2 = wr[j];
                acc[2 * j] = fma(m, w2.x, acc[2 * j]);
                acc[2 * j + 1] = fma(m, w2.y, acc[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const long long idx = (long long)j * P + p;
            const double z = ((double)X[idx] - acc[j]) - sE[j * RP_THREADS + t];
            Anew[idx] = acc[j];
            Y[idx] = __dadd_rn(Y[idx], __dmul_rn(mu, z));
            zz = fma(z, z, zz);
        }
    }
    __shared__ double red[RP_THREADS / 32];
    for (int d = 16; d > 0; d >>= 1) zz += __shfl_down_sync(0xFFFFFFFFu, zz, d);
    if ((t & 31) == 0) red[t >> 5] = zz;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < RP_THREADS / 32; ++w) s += red[w];
        zpart[blockIdx.x] = s;
    }
}

// pass 2, two lanes per pixel row (N = 21).  The one-thread-per-row kernel above needs 128 registers
// (22 accumulators + the row) = 16 warps per SM, and is bound by the latency of its loads.  Here lane
// parity `par` of a lane pair loads the frames i = 2k + par of the row, computes their E / M values, and owns
// the output columns j = 2k + par: the M values cross the pair with shuffles, so every accumulator still
// sees m_0 .. m_20 in order (the same sums, bit for bit), with half the registers and twice the warps.
//
// FUSE: the pass also leaves the Gram partials of the NEXT iteration.  At the end of the update the lane holds
// x, A' and the new Y of its elements, which is all M of the next iteration needs (its mu is known in advance):
// the values replace E in shared memory and the CTA adds the tile's 21 x 21 products to its partial Gram
// matrix with the register-tiled scheme of k_rpca_gram21 (28 blocks of 3 x 3 pairs x 8 row slices).  The
// separate Gram pass (a second sweep over X, A, Y) and one of the two stream synchronisations per iteration
// disappear; only the first iteration still runs k_rpca_gram21.
template <int N, bool FUSE>
__global__ void __launch_bounds__(RP_THREADS, 3)
k_rpca_apply_pair(const uint8_t* __restrict__ X, const double* A, double* Anew,
                  double* __restrict__ Y, long long P, double inv_mu, double thr, double mu,
                  const double* __restrict__ Wg, double* __restrict__ zpart, uint8_t* __restrict__ out,
                  double inv_mu_next, double thr_next, double* __restrict__ gpart_next,
                  const RpcaState* __restrict__ st) {
    bool first = false;                            // graph loop, iteration 0: A == 0 and Y == X / dual_norm, not read
    double dual_norm = 1.0;
    if (st) {                                      // graph loop: parameters and the ping-pong role of the A buffers from the state
        first = st->itr == 0;
        dual_norm = st->dual_norm;
        inv_mu = st->inv_mu; thr = st->thr; mu = st->mu;
        inv_mu_next = st->inv_mu_next; thr_next = st->thr_next;
        if (st->itr & 1) { const double* t = A; A = Anew; Anew = const_cast<double*>(t); }
    }
    constexpr int KH = (N + 1) / 2;                // frames / columns per lane (11)
    constexpr int KP = (KH + 1) & ~1;              // padded to an even count (12): double2 loads of W
    constexpr int ROWS = RP_THREADS / 2;           // pixel rows per CTA step
    constexpr int LD = ROWS + 8;                   // sE row pitch: the two parities of a half-warp use disjoint banks
    __shared__ __align__(16) double sW[N * 2 * KP];   // [i][par][k] = W[i][2k + par] (0 past column N-1)
    constexpr int NPAIRS = N * (N + 1) / 2;
    __shared__ double sG[FUSE ? NPAIRS : 1];       // the CTA's partial Gram matrix of the next iteration
    extern __shared__ double sE[];                 // [N][LD] the rows' E, then [N][LD] their Y, then [N][LD] bytes of X
    double* sY = sE + N * LD;
    uint8_t* sX = reinterpret_cast<uint8_t*>(sY + N * LD);
    const int t = threadIdx.x, par = t & 1, rl = t >> 1;
    for (int q = t; q < N * 2 * KP; q += RP_THREADS) {
        const int i = q / (2 * KP), r = q - i * 2 * KP, pp = r / KP, k = r - pp * KP;
        const int j = 2 * k + pp;
        sW[q] = j < N ? Wg[i * N + j] : 0.0;
    }
    // Gram work split (FUSE): thread = (block of 3 x 3 pairs, slice of every 8th row)
    constexpr int NB = N / 3, NBLK = NB * (NB + 1) / 2, NS = 8;
    static_assert(!FUSE || (N % 3 == 0 && NBLK * NS <= RP_THREADS), "28 blocks x 8 slices");
    const int gslice = t & (NS - 1), gblk = t >> 3;
    const bool gworker = FUSE && gblk < NBLK;
    int bi = 0, bj = 0;
    if (gworker) {
        int q = gblk;
        while (q >= NB - bi) { q -= NB - bi; ++bi; }
        bj = bi + q;
    }
    if (FUSE)
        for (int q = t; q < NPAIRS; q += RP_THREADS) sG[q] = 0.0;
    __syncthreads();
    double zz = 0.0;
    const long long ntiles = (P + ROWS - 1) / ROWS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * ROWS + rl;
        const bool live = p < P;                   // both lanes of a pair agree; dead pairs still shuffle
        // every load of the row first (they are all independent: the kernel lives on loads in flight) ...
        double av[KH], yv[KH];
        uint32_t xv[KH];
#pragma unroll
        for (int k = 0; k < KH; ++k) {
            const int i = 2 * k + par;
            av[k] = yv[k] = 0.0;
            xv[k] = 0u;
            if (live && i < N) {
                const long long idx = (long long)i * P + p;
                xv[k] = X[idx];
                if (!first) {
                    av[k] = A[idx];
                    yv[k] = Y[idx];
                }
            }
        }
        if (first) {
#pragma unroll
            for (int k = 0; k < KH; ++k) yv[k] = __ddiv_rn((double)xv[k], dual_norm);
        }
        // ... then E and M; X, Y and E wait in shared memory for the update below (each lane reads back
        // only what it wrote: it owns frame i = 2k + par and column j = 2k + par)
        double mloc[KH];
#pragma unroll
        for (int k = 0; k < KH; ++k) {
            const int i = 2 * k + par;
            double m = 0.0;
            if (live && i < N) {
                const double x = (double)xv[k];
                const double t2 = __dmul_rn(inv_mu, yv[k]);
                const double e = shrink(__dadd_rn(x - av[k], t2), thr);
                sE[i * LD + rl] = e;
                sY[i * LD + rl] = yv[k];
                sX[i * LD + rl] = (uint8_t)xv[k];
                m = __dadd_rn(x - e, t2);
                out[(long long)i * P + p] = (uint8_t)fmin(fmax(-e, 0.0), 255.0);
            }
            mloc[k] = m;
        }
        double acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double own = mloc[i >> 1];
            const double other = __shfl_xor_sync(0xFFFFFFFFu, own, 1);
            const double m = ((i & 1) == par) ? own : other;
            const double2* wr = reinterpret_cast<const double2*>(sW + (i * 2 + par) * KP);
#pragma unroll
            for (int k = 0; k < KP / 2; ++k) {
                const double2 w2 = wr[k];
                acc[2 * k] = fma(m, w2.x, acc[2 * k]);
                acc[2 * k + 1] = fma(m, w2.y, acc[2 * k + 1]);
            }
        }
        __syncwarp();                              // the pair's E values are in shared memory
        if (live) {
#pragma unroll
            for (int k = 0; k < KH; ++k) {
                const int j = 2 * k + par;
                if (j < N) {
                    const long long idx = (long long)j * P + p;
                    const double x = (double)sX[j * LD + rl];
                    const double z = (x - acc[k]) - sE[j * LD + rl];
                    Anew[idx] = acc[k];
                    const double ynew = __dadd_rn(sY[j * LD + rl], __dmul_rn(mu, z));
                    Y[idx] = ynew;
                    zz = fma(z, z, zz);
                    if (FUSE) {                    // M of the next iteration, exactly as k_rpca_gram21 would compute it
                        const double t2 = __dmul_rn(inv_mu_next, ynew);
                        const double e = shrink(__dadd_rn(x - acc[k], t2), thr_next);
                        sE[j * LD + rl] = __dadd_rn(x - e, t2);
                    }
                }
            }
        } else if (FUSE) {
#pragma unroll
            for (int k = 0; k < KH; ++k)
                if (2 * k + par < N) sE[(2 * k + par) * LD + rl] = 0.0;
        }
        if (FUSE) {
            __syncthreads();                       // the tile's M values are complete
            if (gworker) {
                const double* a0 = sE + (3 * bi) * LD;
                const double* b0 = sE + (3 * bj) * LD;
                double g9[3][3];
#pragma unroll
                for (int x = 0; x < 3; ++x)
#pragma unroll
                    for (int y = 0; y < 3; ++y) g9[x][y] = 0.0;
#pragma unroll 4
                for (int r = gslice; r < ROWS; r += NS) {
                    const double a[3] = {a0[r], a0[LD + r], a0[2 * LD + r]};
                    const double b[3] = {b0[r], b0[LD + r], b0[2 * LD + r]};
#pragma unroll
                    for (int x = 0; x < 3; ++x)
#pragma unroll
                        for (int y = 0; y < 3; ++y) g9[x][y] = fma(a[x], b[y], g9[x][y]);
                }
#pragma unroll
                for (int x = 0; x < 3; ++x)
#pragma unroll
                    for (int y = 0; y < 3; ++y) {
                        double v = g9[x][y];       // the eight slices of a block: adjacent lanes, fixed-order tree
                        v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
                        v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
                        v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
                        const int i = 3 * bi + x, j = 3 * bj + y;
                        if (gslice == 0 && j >= i) sG[i * N - (i * (i - 1)) / 2 + (j - i)] += v;
                    }
            }
            __syncthreads();                       // sE is rewritten by the next tile
        } else {
            __syncwarp();                          // sE is rewritten by the next tile
        }
    }
    if (FUSE) {
        __syncthreads();
        for (int q = t; q < NPAIRS; q += RP_THREADS) gpart_next[(long long)blockIdx.x * NPAIRS + q] = sG[q];
    }
    __shared__ double red[RP_THREADS / 32];
    for (int d = 16; d > 0; d >>= 1) zz += __shfl_down_sync(0xFFFFFFFFu, zz, d);
    if ((t & 31) == 0) red[t >> 5] = zz;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < RP_THREADS / 32; ++w) s += red[w];
        zpart[blockIdx.x] = s;
    }
}

// crop + gray into the compact [n][h*w] stack (convert_grayscale, image_filtering.py:188-196;
// crop_frame :199-203); column order = `order[k]` picks the source frame of column k
__global__ void __launch_bounds__(RP_THREADS)
k_crop_gray(const uint8_t* __restrict__ frames, long long frame_stride, long long pitch, int channels, int x0,
            int y0, int h, int w, int n, int newest_first, uint8_t* __restrict__ out) {
    // grid = (groups of four pixels along the row, rows, frames): no index divisions, one 32-bit store per thread
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, k = blockIdx.z;
    if (x4 >= w) return;
    const int f = newest_first ? n - 1 - k : k;
    const uint8_t* px = frames + (long long)f * frame_stride + (long long)(y0 + y) * pitch + (long long)(x0 + x4) * channels;
    uint8_t* o = out + ((long long)k * h + y) * w + x4;
    const int cnt = min(4, w - x4);
    uint32_t packed = 0;
    for (int i = 0; i < cnt; ++i) {
        uint32_t v;
        if (channels == 3) v = (3735u * px[3 * i] + 19235u * px[3 * i + 1] + 9798u * px[3 * i + 2] + 16384u) >> 15;
        else v = px[i];
        packed |= v << (8 * i);
    }
    if (cnt == 4 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
        *reinterpret_cast<uint32_t*>(o) = packed;
    } else {
        for (int i = 0; i < cnt; ++i) o[i] = (uint8_t)(packed >> (8 * i));
    }
}

// symmetric eigenproblem, cyclic Jacobi (n <= 32): a = V diag(d) V^T, a is destroyed
void jacobi_eigh(int n, std::vector<double>& a, std::vector<double>& d, std::vector<double>& v) {
    v.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += a[(size_t)i * n + i] * a[(size_t)i * n + i];
            for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j];
        }
        if (off <= 1e-34 * diag || off == 0.0) break;   // off-diagonal mass below eps^2 of the diagonal
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
                for (int k = 0; k < n; ++k) {          // rotate rows / columns p, q of a
                    const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq;
                    a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk;
                    a[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
                    v[(size_t)k * n + p] = c * vkp - s * vkq;
                    v[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    d.resize(n);
    for (int i = 0; i < n; ++i) d[i] = a[(size_t)i * n + i];
}

}  // namespace

cudaError_t rpca_alloc(RpcaWork& w, long long P, int nmax) {
    memset(&w, 0, sizeof(w));
    if (nmax > RP_NMAX) return cudaErrorInvalidValue;
    w.P = P;
    w.nmax = nmax;
    w.nctas = 148 * 4;
    w.device_loop = -1;
    const size_t elems = (size_t)P * nmax;
    cudaError_t e;
    if ((e = cudaMalloc(&w.A0, elems * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.A1, elems * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.Y, elems * sizeof(double))) != cudaSuccess) return e;
    const int npairs = nmax * (nmax + 1) / 2;
    if ((e = cudaMalloc(&w.gpart, (size_t)w.nctas * npairs * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.G, (size_t)npairs * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.W, (size_t)nmax * nmax * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.zpart, (size_t)w.nctas * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.sumsq, 16)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.state, 256)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.Vprev, (size_t)nmax * nmax * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMemset(w.gpart, 0, (size_t)w.nctas * npairs * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMallocHost(&w.h_buf, (size_t)(npairs + nmax * nmax + w.nctas + 4) * sizeof(double))) != cudaSuccess) return e;
    return cudaSuccess;
}

void rpca_free(RpcaWork& w) {
    cudaFree(w.A0);
    cudaFree(w.A1);
    cudaFree(w.Y);
    cudaFree(w.gpart);
    cudaFree(w.G);
    cudaFree(w.W);
    cudaFree(w.zpart);
    cudaFree(w.sumsq);
    cudaFree(w.state);
    cudaFree(w.Vprev);
    if (w.graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)w.graph_exec);
    if (w.graph) cudaGraphDestroy((cudaGraph_t)w.graph);
    if (w.h_buf) cudaFreeHost(w.h_buf);
    memset(&w, 0, sizeof(w));
}

cudaError_t launch_crop_gray(cudaStream_t s, const uint8_t* frames, long long frame_stride, long long pitch,
                             int channels, int x0, int y0, int h, int w, int n, int newest_first, uint8_t* out) {
    if (h < 1 || w < 1 || n < 1 || h > 65535 || n > 65535) return cudaErrorInvalidValue;
    const int groups = (w + 3) / 4;
    const int threads = groups >= 128 ? 128 : 32 * ((groups + 31) / 32);
    const dim3 grid((groups + threads - 1) / threads, h, n);
    k_crop_gray<<<grid, threads, 0, s>>>(frames, frame_stride, pitch, channels, x0, y0, h, w, n, newest_first, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// The IALM loop without the host (n = 21, the reference's batch).  One CUDA graph per work area:
//   norms -> setup -> init -> gram21 -> WHILE (not converged) { gram_reduce -> eigen -> fused apply pass -> check }
// The WHILE node is a conditional graph node; k_rpca_check ends the loop with cudaGraphSetConditional.
// Nothing is copied to the host and nothing synchronises: swb_submit stays asynchronous in RPCA mode and
// the per-iteration cost is the kernels plus a few microseconds of graph scheduling instead of two
// stream synchronisations, three small copies and the host eigenproblem.
// ------------------------------------------------------------------------------------------
__global__ void k_rpca_setup(const unsigned long long* __restrict__ sumsq, RpcaState* __restrict__ st, double lmbda,
                             cudaGraphConditionalHandle loop) {
    if (threadIdx.x != 0) return;
    const double norm_two = sqrt((double)sumsq[0]);                       // norm(Y.ravel(), 2) == norm(X, 'fro')
    const double norm_inf = (double)(unsigned int)sumsq[1] / lmbda;
    RpcaState r;
    r.zero = norm_two == 0.0;
    r.dual_norm = r.zero ? 1.0 : fmax(norm_two, norm_inf);
    r.dnorm = norm_two;
    r.mu = r.zero ? 1.0 : 1.25 / norm_two;
    r.inv_mu = 1 / r.mu;
    r.thr = lmbda / r.mu;
    const double mu_next = fmin(r.mu * 1.5, r.mu * 1e7);
    r.inv_mu_next = 1 / mu_next;
    r.thr_next = lmbda / mu_next;
    r.itr = 0;
    r.done = r.zero;
    r.sweeps = 0;
    r.cyc_eigen = r.cyc_jacobi = 0;
    for (int i = 0; i < 8; ++i) r.sweeps_hist[i] = 0;
    *st = r;
    cudaGraphSetConditional(loop, r.zero ? 0u : 1u);                      // an all-black batch: E == 0, no iteration
}

// 1 / sqrt(x) from a float32 seed that is good to ~2^-22: with x y^2 = 1 - e, x^(-1/2) = y (1 + e/2 + 3 e^2 / 8 + O(e^3)),
// i.e. ONE step with an error below 2^-64, four dependent double operations.  (A dependent FP64 operation costs
// ~40 cycles on a B200: the rotation's chain of them, not its instruction count, is what a Jacobi round waits for.)
__device__ __forceinline__ double rsqrt_refine(double x, float seed) {
    const double y = (double)seed;
    const double e = fma(-(x * y), y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}

// shared memory through explicit 32-bit shared-space addresses (see the Jacobi rounds of k_rpca_eigen21)
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_s32(uint32_t addr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// G (packed upper triangle) -> W = V diag((S - 1/mu) / S) V^T with G = V diag(S^2) V^T, for n = 21.
// One CTA.  The eigenproblem is a parallel-order Jacobi iteration on shared memory (round-robin schedule: 21 rounds
// of 10 disjoint rotations per sweep), warm-started in the eigenbasis of the previous IALM iteration, where G
// is nearly diagonal (four to six sweeps instead of seven).  Rotation formulas and stopping rule are those of the
// host solver (jacobi_eigh).
__global__ void __launch_bounds__(256)
k_rpca_eigen21(const double* __restrict__ G, double* __restrict__ Wg, double* __restrict__ Vprev,
               RpcaState* __restrict__ st, double jtol) {
    constexpr int n = 21, LD = 23;
    __shared__ double a[n * LD], v[n * LD], vp[n * LD], tmp[n * LD], dsc[n];
    __shared__ double cs[10][2], red[8][2];
    __shared__ int pq[10][2];
    __shared__ int s_go;
    const int t = threadIdx.x, lane = t & 31;
    const long long c_start = clock64();
    const bool warm = st->itr > 0;
    const double inv_mu = st->inv_mu;
    for (int q = t; q < n * n; q += 256) {
        const int i = q / n, j = q - i * n;
        const int lo = min(i, j), hi = max(i, j);
        a[i * LD + j] = G[lo * n - (lo * (lo - 1)) / 2 + (hi - lo)];
        vp[i * LD + j] = warm ? Vprev[q] : (i == j ? 1.0 : 0.0);
        v[i * LD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (warm) {
        // B = Vp^T G Vp (symmetric by construction of the second product: upper triangle mirrored)
        for (int q = t; q < n * n; q += 256) {
            const int i = q / n, j = q - i * n;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(a[i * LD + k], vp[k * LD + j]));
            tmp[i * LD + j] = acc;
        }
        __syncthreads();
        double mine[2] = {0.0, 0.0};
        for (int r = 0, q = t; q < n * n; q += 256, ++r) {
            const int i = q / n, j = q - i * n;
            const int lo = min(i, j), hi = max(i, j);
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(vp[k * LD + lo], tmp[k * LD + hi]));
            mine[r] = acc;
        }
        __syncthreads();
        for (int r = 0, q = t; q < n * n; q += 256, ++r) a[(q / n) * LD + (q % n)] = mine[r];
        __syncthreads();
    }
    // ---- Jacobi sweeps, all 256 threads.  A round has two steps separated by __syncthreads: ten lanes form the
    // round's rotations; then thread (row, pair) rotates its two columns of v and thread (pair_i, pair_j) takes
    // one 2 x 2 block of a through the column rotation of pair_j and the row rotation of pair_i (the same
    // operations in the same order as the one-warp version, which walked columns then rows with a lane per
    // row: a block needs nothing from outside itself, so the two passes fuse).  The bye of a round is index r.
    const long long c_j = clock64();
    uint32_t a_s = (uint32_t)__cvta_generic_to_shared(a), v_s = (uint32_t)__cvta_generic_to_shared(v);
    uint32_t cs_s = (uint32_t)__cvta_generic_to_shared(cs), pq_s = (uint32_t)__cvta_generic_to_shared(pq);
    asm volatile("" : "+r"(a_s), "+r"(v_s), "+r"(cs_s), "+r"(pq_s));   // opaque: kept in registers, not re-derived per use
    int sweeps = 0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        {
            // stopping test: thread (i, j) squares one element of the upper triangle, warps add up with a fixed
            // shuffle tree, warp 0 adds the eight partial sums
            double off = 0.0, diag = 0.0;
            if (t < n * (n + 1) / 2) {
                int i = 0, q = t;
                while (q >= n - i) { q -= n - i; ++i; }
                const double x = a[i * LD + i + q];
                if (q == 0) diag = x * x; else off = x * x;
            }
            for (int d = 16; d > 0; d >>= 1) {
                off += __shfl_xor_sync(0xFFFFFFFFu, off, d);
                diag += __shfl_xor_sync(0xFFFFFFFFu, diag, d);
            }
            if (lane == 0) { red[t >> 5][0] = off; red[t >> 5][1] = diag; }
            __syncthreads();
            if (t == 0) {
                off = diag = 0.0;
                for (int wq = 0; wq < 8; ++wq) { off += red[wq][0]; diag += red[wq][1]; }
                s_go = !(off <= jtol * diag || off == 0.0);   // off-diagonal mass below jtol of the diagonal (squares)
            }
        }
        __syncthreads();
        if (!s_go) break;
        ++sweeps;
        for (int r = 0; r < n; ++r) {
            // round r of the circle schedule on 22 players (player 21 = the bye, paired with r).  Shared memory is
            // addressed through 32-bit shared-space addresses formed once (a_s, v_s, cs_s, pq_s): with plain array
            // accesses the compiler re-reads SR_CgaCtaId (S2R) in every divergent step of every round to rebuild the
            // shared window base, right at the head of the round's critical path.
            if (t < 10) {
                int p = (r + t + 1) % n, q = (r + n - t - 1) % n;
                if (p > q) { const int x = p; p = q; q = x; }
                const double apq = lds_f64(a_s + 8 * (p * LD + q));
                double c = 1.0, sn = 0.0;
                if (apq != 0.0) {
                    // The rotation that zeroes a_pq: tan(2 phi) = b / d with d = aqq - app, b = 2 apq, |phi| <= pi / 4.
                    // With r = 1 / sqrt(d^2 + b^2): cos(2 phi) = |d| r, cos^2(phi) = (1 + |d| r) / 2 = h,
                    // c = h / sqrt(h), s = sgn(d b) |b| r / (2 sqrt(h)); h is in [1/2, 1], so nothing cancels.  The two
                    // reciprocal square roots have arguments of known range (d and b are scaled by a power of two
                    // first), so they are float32 MUFU seeds (computed ahead, in float32) + one cubic correction step
                    // in double (rsqrt_refine) instead of the library's sqrt / div / rsqrt.
                    const double d = lds_f64(a_s + 8 * (q * LD + q)) - lds_f64(a_s + 8 * (p * LD + p)), b2 = 2.0 * apq;
                    // scale by the power of two that brings max(|d|, |b|) to [1, 2): exponent bits, two multiplies
                    const int ef = min(max((__double2hiint(fmax(fabs(d), fabs(b2))) >> 20) & 0x7FF, 1), 2045);
                    const double sc = __hiloint2double((2046 - ef) << 20, 0);
                    const double ds = fabs(d) * sc, bs = fabs(b2) * sc;
                    // float32 first (a few cycles per operation): the seeds of both reciprocal square roots
                    const float dsf = (float)ds, bsf = (float)bs;
                    const float rf = rsqrtf(fmaf(dsf, dsf, bsf * bsf));
                    const float ihf = rsqrtf(fmaf(0.5f * dsf, rf, 0.5f));
                    const double rr = rsqrt_refine(fma(ds, ds, bs * bs), rf);  // the argument is in [1, 8)
                    const double h = fma(0.5 * ds, rr, 0.5);                   // [1/2, 1]
                    const double ih = rsqrt_refine(h, ihf);
                    c = h * ih;
                    sn = copysign(0.5 * bs * rr * ih, (d >= 0.0) == (b2 >= 0.0) ? 1.0 : -1.0);
                }
                sts_f64(cs_s + 16 * t, c);
                sts_f64(cs_s + 16 * t + 8, sn);
                sts_s32(pq_s + 8 * t, p);
                sts_s32(pq_s + 8 * t + 4, q);
            }
            __syncthreads();
            if (t < n * 10) {                                  // columns p, q of v, one row
                const int row = t / 10, i = t - row * 10;
                const int p = lds_s32(pq_s + 8 * i), q = lds_s32(pq_s + 8 * i + 4);
                const double c = lds_f64(cs_s + 16 * i), sn = lds_f64(cs_s + 16 * i + 8);
                const uint32_t ap = v_s + 8 * (row * LD + p), aq = v_s + 8 * (row * LD + q);
                const double vp_ = lds_f64(ap), vq_ = lds_f64(aq);
                sts_f64(ap, c * vp_ - sn * vq_);
                sts_f64(aq, sn * vp_ + c * vq_);
            }
            const int blk = 255 - t;                           // 11 x 11 blocks of a (pair 10 = the bye: one index, no rotation)
            if (blk < 121) {
                const int bi = blk / 11, bj = blk - bi * 11;
                const bool ri = bi < 10, rj = bj < 10;
                const int pi = ri ? lds_s32(pq_s + 8 * bi) : r, qi = ri ? lds_s32(pq_s + 8 * bi + 4) : r;
                const int pj = rj ? lds_s32(pq_s + 8 * bj) : r, qj = rj ? lds_s32(pq_s + 8 * bj + 4) : r;
                const uint32_t a00 = a_s + 8 * (pi * LD + pj), a01 = a_s + 8 * (pi * LD + qj);
                const uint32_t a10 = a_s + 8 * (qi * LD + pj), a11 = a_s + 8 * (qi * LD + qj);
                double x00 = lds_f64(a00), x01 = lds_f64(a01), x10 = lds_f64(a10), x11 = lds_f64(a11);
                if (rj) {                                      // columns pj, qj (rows pi, qi)
                    const double c = lds_f64(cs_s + 16 * bj), sn = lds_f64(cs_s + 16 * bj + 8);
                    const double u0 = c * x00 - sn * x01, u1 = sn * x00 + c * x01;
                    const double w0 = c * x10 - sn * x11, w1 = sn * x10 + c * x11;
                    x00 = u0; x01 = u1; x10 = w0; x11 = w1;
                }
                if (ri) {                                      // rows pi, qi (columns pj, qj)
                    const double c = lds_f64(cs_s + 16 * bi), sn = lds_f64(cs_s + 16 * bi + 8);
                    const double u0 = c * x00 - sn * x10, u1 = sn * x00 + c * x10;
                    const double w0 = c * x01 - sn * x11, w1 = sn * x01 + c * x11;
                    x00 = u0; x10 = u1; x01 = w0; x11 = w1;
                }
                sts_f64(a00, x00);
                if (rj) sts_f64(a01, x01);
                if (ri) sts_f64(a10, x10);
                if (ri && rj) sts_f64(a11, x11);
            }
            __syncthreads();
        }
    }
    if (t < n) {
        const double d = a[t * LD + t];
        const double sv = d > 0.0 ? sqrt(d) : 0.0;
        dsc[t] = sv > 0.0 ? __ddiv_rn(sv - inv_mu, sv) : 0.0;   // exactly dependent columns are skipped (see header)
    }
    if (t == 0) {
        st->sweeps += sweeps;
        st->cyc_jacobi += clock64() - c_j;
        if (st->itr < 8) st->sweeps_hist[st->itr] = sweeps;
    }
    __syncthreads();
    // V = Vp Vb (kept for the next warm start), then W = V diag(f) V^T
    for (int q = t; q < n * n; q += 256) {
        const int i = q / n, j = q - i * n;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(vp[i * LD + k], v[k * LD + j]));
        tmp[i * LD + j] = acc;
        Vprev[q] = acc;
    }
    __syncthreads();
    for (int q = t; q < n * n; q += 256) {
        const int i = q / n, j = q - i * n;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(tmp[i * LD + k], dsc[k]), tmp[j * LD + k]));
        Wg[q] = acc;
    }
    __syncthreads();
    if (t == 0) st->cyc_eigen += clock64() - c_start;
}

// |Z|_F^2 from the per-CTA partials (fixed order), the stopping test (image_filtering.py:296-298), mu *= rho
__global__ void k_rpca_check(const double* __restrict__ zpart, int napply, RpcaState* __restrict__ st, double lmbda,
                             double tol, double rho, int maxiter, cudaGraphConditionalHandle loop) {
    if (threadIdx.x != 0) return;
    double zz = 0.0;
    for (int c = 0; c < napply; ++c) zz += zpart[c];
    RpcaState r = *st;
    r.mu = fmin(r.mu * rho, r.mu * 1e7);
    r.inv_mu = 1 / r.mu;
    r.thr = lmbda / r.mu;
    const double mu_next = fmin(r.mu * rho, r.mu * 1e7);
    r.inv_mu_next = 1 / mu_next;
    r.thr_next = lmbda / mu_next;
    r.itr += 1;
    r.done = (sqrt(zz) / r.dnorm < tol) || r.itr >= maxiter;
    *st = r;
    cudaGraphSetConditional(loop, r.done ? 0u : 1u);
}

// Builds (once per work area and buffer set) and launches the graph.  Returns cudaErrorNotSupported when the
// graph cannot be built (the caller then runs the host loop).
static cudaError_t rpca_run_graph(cudaStream_t s, const uint8_t* X, long long P, RpcaWork& w, uint8_t* out,
                                  int* n_launches) {
    constexpr int n = 21;
    const double lmbda = 0.01, tol = 0.001, rho = 1.5;
    const int maxiter = 100;
    const long long total = (long long)n * P;
    const int npairs = n * (n + 1) / 2;
    const int nctas = (int)std::min<long long>(w.nctas, (P + RP_THREADS - 1) / RP_THREADS);
    const int napply = std::min(nctas, 148 * 3);
    if (w.graph_failed) return cudaErrorNotSupported;
    if (!w.graph_exec || w.g_X != X || w.g_out != out || w.g_P != P) {
        if (w.graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)w.graph_exec); w.graph_exec = nullptr; }
        if (w.graph) { cudaGraphDestroy((cudaGraph_t)w.graph); w.graph = nullptr; }
        cudaStream_t c1 = nullptr, c2 = nullptr;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool capturing1 = false, capturing2 = false;
        auto fail = [&](cudaError_t e) {
            cudaGraph_t junk = nullptr;
            if (capturing2) cudaStreamEndCapture(c2, &junk);
            if (capturing1) { junk = nullptr; cudaStreamEndCapture(c1, &junk); if (junk) cudaGraphDestroy(junk); }
            if (c1) cudaStreamDestroy(c1);
            if (c2) cudaStreamDestroy(c2);
            cudaGetLastError();
            w.graph_failed = 1;
            (void)e;
            return cudaErrorNotSupported;
        };
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&c1, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
        if ((e = cudaStreamCreateWithFlags(&c2, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
        if ((e = cudaStreamBeginCapture(c1, cudaStreamCaptureModeRelaxed)) != cudaSuccess) return fail(e);
        capturing1 = true;
        RpcaState* st = reinterpret_cast<RpcaState*>(w.state);
        // the loop handle is needed by k_rpca_setup, which comes before the loop node: create it first
        cudaStreamCaptureStatus status;
        const cudaGraphNode_t* deps = nullptr;
        size_t ndeps = 0;
        if ((e = cudaStreamGetCaptureInfo(c1, &status, nullptr, &graph, &deps, &ndeps)) != cudaSuccess) return fail(e);
        cudaGraphConditionalHandle loop;
        if ((e = cudaGraphConditionalHandleCreate(&loop, graph, 1, cudaGraphCondAssignDefault)) != cudaSuccess) return fail(e);
        cudaMemsetAsync(w.sumsq, 0, 16, c1);
        k_rpca_norms<<<nctas, RP_THREADS, 0, c1>>>(X, total, w.sumsq, reinterpret_cast<unsigned int*>(w.sumsq + 1));
        k_rpca_setup<<<1, 32, 0, c1>>>(w.sumsq, st, lmbda, loop);
        cudaMemsetAsync(out, 0, (size_t)total, c1);       // what an all-black batch (no iteration at all) leaves behind
        k_rpca_gram21<<<nctas, RP_THREADS, 21 * (RP_THREADS + 8) * sizeof(double), c1>>>(X, w.A0, w.Y, P, 0.0, 0.0, w.gpart, st);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
        if ((e = cudaStreamGetCaptureInfo(c1, &status, nullptr, &graph, &deps, &ndeps)) != cudaSuccess) return fail(e);
        cudaGraphNodeParams cp = {};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = loop;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        cudaGraphNode_t loop_node;
        if ((e = cudaGraphAddNode(&loop_node, graph, deps, ndeps, &cp)) != cudaSuccess) return fail(e);
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        if ((e = cudaStreamUpdateCaptureDependencies(c1, &loop_node, 1, cudaStreamSetCaptureDependencies)) != cudaSuccess)
            return fail(e);
        // ---- loop body: the first iteration's Gram partials come from k_rpca_gram21 (nctas CTAs), the later ones
        // from the fused pass (napply CTAs): both are padded to nctas rows of partials (zeros) so one reduce fits
        if ((e = cudaStreamBeginCaptureToGraph(c2, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed)) != cudaSuccess)
            return fail(e);
        capturing2 = true;
        k_rpca_gram_reduce_st<<<(npairs * 32 + 127) / 128, 128, 0, c2>>>(w.gpart, nctas, napply, npairs, w.G, st);
        static const double jtol = [] { const char* e = getenv("SWB_RPCA_JTOL"); return e ? atof(e) : 1e-28; }();   // off-diagonal norm below 1e-14 of the diagonal norm
        k_rpca_eigen21<<<1, 256, 0, c2>>>(w.G, w.W, w.Vprev, st, jtol);
        k_rpca_apply_pair<21, true><<<napply, RP_THREADS, 21 * (RP_THREADS / 2 + 8) * 17, c2>>>(
            X, w.A0, w.A1, w.Y, P, 0.0, 0.0, 0.0, w.W, w.zpart, out, 0.0, 0.0, w.gpart, st);
        k_rpca_check<<<1, 32, 0, c2>>>(w.zpart, napply, st, lmbda, tol, rho, maxiter, loop);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
        cudaGraph_t body_out = nullptr;
        e = cudaStreamEndCapture(c2, &body_out);
        capturing2 = false;
        if (e != cudaSuccess) return fail(e);
        e = cudaStreamEndCapture(c1, &graph);
        capturing1 = false;
        if (e != cudaSuccess) return fail(e);
        if ((e = cudaGraphInstantiate(&exec, graph, 0)) != cudaSuccess) { cudaGraphDestroy(graph); return fail(e); }
        cudaStreamDestroy(c1);
        cudaStreamDestroy(c2);
        w.graph = graph;
        w.graph_exec = exec;
        w.g_X = X;
        w.g_out = out;
        w.g_P = P;
    }
    cudaError_t e = cudaGraphLaunch((cudaGraphExec_t)w.graph_exec, s);
    if (e != cudaSuccess) return e;
    if (n_launches) *n_launches += 3;          // norms, setup, first Gram pass; the loop's launches are counted when the state is read back
    return cudaSuccess;
}

// inexact_augmented_lagrange_multiplier (image_filtering.py:256-301) on the device.
// X: [n][P] uint8 (column k of the reference's matrix = X[k]); out: [n][P] uint8 = clip(-E, 0, 255).
// Synchronises the stream every iteration (the stopping test and the 21 x 21 eigenproblem run on the host).
cudaError_t rpca_run(cudaStream_t s, const uint8_t* X, int n, long long P, RpcaWork& w, uint8_t* out, int* iters,
                     int* n_launches) {
    if (n < 1 || n > w.nmax || P > w.P) return cudaErrorInvalidValue;
    w.last_mode = 0;
    // Where the iteration loop runs.  Device (CUDA-graph WHILE node): nothing blocks, nothing is copied, the 21 x 21
    // eigenproblem is one CTA's work (60-90 us per iteration); host: two stream synchronisations per iteration,
    // eigenproblem in ~40 us.  Measured on a B200 (round 2, all 256 threads in the Jacobi rounds): 1080p full frame
    // 9.26 ms (device) vs 9.67 (host) per batch, 320x160 ROI 2.10 vs 2.59 — the device loop is the default for the
    // reference's batch size; SWB_RPCA_HOST_LOOP=1 / swb_set_option("rpca_device_loop", 0) select the host loop.
    static const int env_loop = [] { const char* e = getenv("SWB_RPCA_HOST_LOOP"); return e ? (e[0] == '1' ? 0 : 1) : -1; }();
    const int want = w.device_loop >= 0 ? w.device_loop : (env_loop >= 0 ? env_loop : 1);
    if (n == 21 && want == 1) {
        static PerDeviceOnce once_g;
        if (once_g.need()) {
            cudaFuncSetAttribute(k_rpca_gram21, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 RP_NMAX * (RP_THREADS + 1) * (int)sizeof(double));
            cudaFuncSetAttribute(k_rpca_apply_pair<21, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 21 * (RP_THREADS / 2 + 8) * 17);
        }
        const cudaError_t eg = rpca_run_graph(s, X, P, w, out, n_launches);
        if (eg == cudaSuccess) {
            w.last_mode = 1;
            if (iters) *iters = -1;                 // not known yet: rpca_read_state() after the stream has drained
            return cudaSuccess;
        }
        if (eg != cudaErrorNotSupported) return eg;
    }
    const double lmbda = 0.01, tol = 0.001, rho = 1.5;
    const int maxiter = 100;
    const long long total = (long long)n * P;
    const int npairs = n * (n + 1) / 2;
    const int nctas = (int)std::min<long long>(w.nctas, (P + RP_THREADS - 1) / RP_THREADS);
    double* hG = w.h_buf;
    double* hW = hG + w.nmax * (w.nmax + 1) / 2;
    double* hZ = hW + w.nmax * w.nmax;
    unsigned long long* hS = reinterpret_cast<unsigned long long*>(hZ + w.nctas);
    cudaError_t e;
    int launches = 0;

    static PerDeviceOnce once;
    const int smem_gram = RP_NMAX * (RP_THREADS + 1) * (int)sizeof(double);
    const int smem_apply = (RP_NMAX * RP_NMAX + 2 * RP_NMAX * RP_THREADS) * (int)sizeof(double);
    if (once.need()) {
        // a failed shared-memory opt-in is reported here, not as a generic launch error later
        cudaError_t ea[6] = {
            cudaFuncSetAttribute(k_rpca_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_gram),
            cudaFuncSetAttribute(k_rpca_gram21, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_gram),
            cudaFuncSetAttribute(k_rpca_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_apply),
            cudaFuncSetAttribute(k_rpca_apply_n<21>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 21 * RP_THREADS * (int)sizeof(double)),
            cudaFuncSetAttribute(k_rpca_apply_pair<21, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 21 * (RP_THREADS / 2 + 8) * 17),
            cudaFuncSetAttribute(k_rpca_apply_pair<21, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 21 * (RP_THREADS / 2 + 8) * 17)};
        for (cudaError_t x : ea)
            if (x != cudaSuccess) return x;
    }

    cudaMemsetAsync(w.sumsq, 0, 16, s);
    k_rpca_norms<<<nctas, RP_THREADS, 0, s>>>(X, total, w.sumsq, reinterpret_cast<unsigned int*>(w.sumsq + 1));
    cudaMemcpyAsync(hS, w.sumsq, 16, cudaMemcpyDeviceToHost, s);
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    const double norm_two = std::sqrt((double)hS[0]);                 // norm(Y.ravel(), 2) == norm(X, 'fro')
    const double norm_inf = (double)(unsigned int)hS[1] / lmbda;
    if (norm_two == 0.0) {                                            // all-black batch: E == 0
        cudaMemsetAsync(out, 0, (size_t)total, s);
        if (iters) *iters = 0;
        if (n_launches) *n_launches += 1;
        return cudaGetLastError();
    }
    const double dual_norm = std::max(norm_two, norm_inf);
    const double dnorm = norm_two;
    double mu = 1.25 / norm_two;
    k_rpca_init<<<nctas, RP_THREADS, 0, s>>>(X, total, dual_norm, w.A0, w.Y, nullptr, nullptr);
    launches += 2;

    double* Aold = w.A0;
    double* Anew = w.A1;
    std::vector<double> g((size_t)n * n), d, v, vprev, tmp((size_t)n * n), vb;
    int itr = 0;
    bool have_G = false;                           // hG already holds this iteration's Gram matrix (fused pass)
    while (true) {
        const double inv_mu = 1 / mu;
        const double thr = lmbda / mu;
        if (!have_G) {                             // first iteration (or no fused pass): the Gram pass on its own
            if (n == 21)
                k_rpca_gram21<<<nctas, RP_THREADS, 21 * (RP_THREADS + 8) * sizeof(double), s>>>(X, Aold, w.Y, P, inv_mu, thr, w.gpart, nullptr);
            else
                k_rpca_gram<<<nctas, RP_THREADS, n * (RP_THREADS + 1) * sizeof(double), s>>>(X, Aold, w.Y, n, P, inv_mu, thr, w.gpart);
            k_rpca_gram_reduce<<<(npairs * 32 + 127) / 128, 128, 0, s>>>(w.gpart, nctas, npairs, w.G);
            cudaMemcpyAsync(hG, w.G, (size_t)npairs * sizeof(double), cudaMemcpyDeviceToHost, s);
            launches += 2;
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
        }
        for (int i = 0, q = 0; i < n; ++i)
            for (int j = i; j < n; ++j, ++q) g[(size_t)i * n + j] = g[(size_t)j * n + i] = hG[q];
        if (vprev.empty()) {
            jacobi_eigh(n, g, d, v);
        } else {
            // warm start: in the eigenbasis of the previous iteration G is nearly diagonal, so the
            // Jacobi sweeps that remain are two or three instead of seven (the eigenproblem is the
            // bulk of an iteration for ROI-sized frames).  B = Vp^T G Vp;  G = (Vp Vb) diag(d) (Vp Vb)^T
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc += g[(size_t)i * n + k] * vprev[(size_t)k * n + j];
                    tmp[(size_t)i * n + j] = acc;
                }
            for (int i = 0; i < n; ++i)
                for (int j = i; j < n; ++j) {
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc += vprev[(size_t)k * n + i] * tmp[(size_t)k * n + j];
                    g[(size_t)i * n + j] = g[(size_t)j * n + i] = acc;
                }
            jacobi_eigh(n, g, d, vb);
            v.assign((size_t)n * n, 0.0);
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < n; ++k) {
                    const double a = vprev[(size_t)i * n + k];
                    for (int j = 0; j < n; ++j) v[(size_t)i * n + j] += a * vb[(size_t)k * n + j];
                }
        }
        vprev = v;
        // W = V diag((S - 1/mu) / S) V^T  (svp == n: every singular value is shifted, none is dropped)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) {
                    const double sv = d[k] > 0.0 ? std::sqrt(d[k]) : 0.0;
                    if (sv <= 0.0) continue;       // exactly dependent columns: U is undefined there (see header)
                    acc += v[(size_t)i * n + k] * ((sv - inv_mu) / sv) * v[(size_t)j * n + k];
                }
                hW[(size_t)i * n + j] = acc;
            }
        cudaMemcpyAsync(w.W, hW, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, s);
        static const bool apply_one = [] { const char* e = getenv("SWB_RPCA_APPLY1"); return e && e[0] == '1'; }();
        int napply = nctas;                        // CTAs of the apply pass = |Z|^2 partials to add up
        static const bool no_fuse = [] { const char* e = getenv("SWB_RPCA_FUSE"); return e && e[0] == '0'; }();
        if (n == 21 && !apply_one) {
            napply = std::min(nctas, 148 * 3);     // three CTAs per SM are resident: one round of equal shares
            const size_t smem = 21 * (RP_THREADS / 2 + 8) * 17;        // E, Y (doubles), X (bytes)
            if (no_fuse) {
                k_rpca_apply_pair<21, false><<<napply, RP_THREADS, smem, s>>>(X, Aold, Anew, w.Y, P, inv_mu, thr, mu, w.W,
                                                                             w.zpart, out, 0.0, 0.0, nullptr, nullptr);
            } else {
                // the pass also leaves the Gram partials of the next iteration (its mu is known now)
                const double mu_next = std::min(mu * rho, mu * 1e7);
                k_rpca_apply_pair<21, true><<<napply, RP_THREADS, smem, s>>>(X, Aold, Anew, w.Y, P, inv_mu, thr, mu, w.W,
                                                                            w.zpart, out, 1 / mu_next, lmbda / mu_next, w.gpart, nullptr);
                k_rpca_gram_reduce<<<(npairs * 32 + 127) / 128, 128, 0, s>>>(w.gpart, napply, npairs, w.G);
                cudaMemcpyAsync(hG, w.G, (size_t)npairs * sizeof(double), cudaMemcpyDeviceToHost, s);
                have_G = true;
                launches += 1;
            }
        }
        else if (n == 21)
            k_rpca_apply_n<21><<<nctas, RP_THREADS, 21 * RP_THREADS * sizeof(double), s>>>(X, Aold, Anew, w.Y, P, inv_mu, thr, mu,
                                                                                           w.W, w.zpart, out);
        else
            k_rpca_apply<<<nctas, RP_THREADS, (size_t)(n * n + 2 * n * RP_THREADS) * sizeof(double), s>>>(
                X, Aold, Anew, w.Y, n, P, inv_mu, thr, mu, w.W, w.zpart, out);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;          // a failed launch of this iteration
        if ((e = cudaMemcpyAsync(hZ, w.zpart, (size_t)napply * sizeof(double), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        launches += 1;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
        double zz = 0.0;
        for (int c = 0; c < napply; ++c) zz += hZ[c];
        std::swap(Aold, Anew);
        mu = std::min(mu * rho, mu * 1e7);
        ++itr;
        if (std::sqrt(zz) / dnorm < tol || itr >= maxiter) break;
    }
    if (iters) *iters = itr;
    w.host_iters = itr;
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}

// Iterations / Jacobi sweeps of the last run (synchronises the stream when the graph loop ran).
cudaError_t rpca_read_state(cudaStream_t s, RpcaWork& w, int* iters, int* sweeps, int* mode) {
    if (mode) *mode = w.last_mode;
    if (w.last_mode == 1) {
        RpcaState r;
        cudaError_t e = cudaMemcpyAsync(w.h_buf, w.state, sizeof(RpcaState), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        memcpy(&r, w.h_buf, sizeof(r));
        static const bool dbg = [] { const char* e = getenv("SWB_RPCA_DEBUG"); return e && e[0] == '1'; }();
        if (dbg)
            fprintf(stderr, "rpca: %d iterations, %d sweeps (first eight: %d %d %d %d %d %d %d %d), eigen kernel %.1f us / iteration "
                            "of which Jacobi %.1f us (at 1.965 GHz)\n", r.itr, r.sweeps, r.sweeps_hist[0], r.sweeps_hist[1],
                    r.sweeps_hist[2], r.sweeps_hist[3], r.sweeps_hist[4], r.sweeps_hist[5], r.sweeps_hist[6], r.sweeps_hist[7],
                    r.cyc_eigen / 1965.0 / std::max(r.itr, 1), r.cyc_jacobi / 1965.0 / std::max(r.itr, 1));
        if (iters) *iters = r.itr;
        if (sweeps) *sweeps = r.sweeps;
    } else {
        if (iters) *iters = w.host_iters;
        if (sweeps) *sweeps = 0;
    }
    return cudaSuccess;
}


// ------------------------------------------------------------------------------------------
// bilateral_blur = cv2.bilateralFilter(frame, 7, 15, 1) (image_filtering.py:304-307,
// data_structures.py:194).  OpenCV's definition for 8-bit single-channel images
// (modules/imgproc bilateral_filter, 4.x): radius = d / 2; the taps are the offsets with
// sqrt(i^2 + j^2) <= radius visited i-outer / j-inner (29 for d = 7, centre included);
// space weight = (float)exp(-r^2 / (2 sigma_space^2)), colour weight = (float)exp(-dv^2 /
// (2 sigma_color^2)) from a 256-entry table; BORDER_REFLECT_101; float32 accumulation of
// w and v * w in tap order; result = cvRound(sum / wsum).  The tables are built on the host with
// the same double-precision exp.  This is bit-identical to cv2 with setUseOptimized(False);
// OpenCV's SIMD body differs from its own scalar code on about one pixel per million (exact .5
// ties of the float quotient), so against the optimised build parity is "equal up to 1 grey level
// on <= 2 ppm of the pixels" (tests).
// ------------------------------------------------------------------------------------------
void bilateral_lut(int d, double sigma_color, double sigma_space, BilateralLut& lut) {
    int radius = d / 2;
    if (radius < 1) radius = 1;
    if (radius > 3) radius = 3;                    // 64-tap table: d <= 7 (the reference uses 7)
    const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    for (int i = 0; i < 256; ++i) lut.color[i] = (float)std::exp((double)i * i * gc);
    int k = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const double r = std::sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            lut.space[k] = (float)std::exp(r * r * gs);
            lut.dy[k] = i;
            lut.dx[k] = j;
            ++k;
        }
    lut.ntaps = k;
    lut.radius = radius;
}

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// thread = pixel; a warp covers 32 consecutive pixels of a row so that the thresholded result
// is one ballot per word
__global__ void __launch_bounds__(256)
k_bilateral(const uint8_t* __restrict__ in, int n, int h, int w, const BilateralLut* __restrict__ lutp, int reverse,
            uint8_t* __restrict__ out, int thresh, uint32_t* __restrict__ bits, int wpr_bits) {
    __shared__ BilateralLut lut;
    for (int i = threadIdx.x; i < (int)(sizeof(BilateralLut) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&lut)[i] = reinterpret_cast<const uint32_t*>(lutp)[i];
    __syncthreads();
    const int wpr = (w + 31) / 32;
    const int lane = threadIdx.x & 31;
    const long long nwords = (long long)n * h * wpr;
    for (long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; word < nwords;
         word += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int f = (int)(word / ((long long)h * wpr));
        const long long rem = word - (long long)f * h * wpr;
        const int y = (int)(rem / wpr), x = (int)(rem - (long long)y * wpr) * 32 + lane;
        const uint8_t* img = in + (long long)(reverse ? n - 1 - f : f) * h * w;
        int res = 0;
        if (x < w) {
            const int c = img[(long long)y * w + x];
            float sum = 0.f, wsum = 0.f;
            for (int k = 0; k < lut.ntaps; ++k) {
                const int yy = reflect101(y + lut.dy[k], h), xx = reflect101(x + lut.dx[k], w);
                const int v = img[(long long)yy * w + xx];
                const float wk = __fmul_rn(lut.space[k], lut.color[abs(v - c)]);
                wsum = __fadd_rn(wsum, wk);
                sum = __fadd_rn(sum, __fmul_rn((float)v, wk));
            }
            res = __float2int_rn(__fdiv_rn(sum, wsum));
            res = min(max(res, 0), 255);
            if (out) out[((long long)f * h + y) * w + x] = (uint8_t)res;
        }
        if (bits) {
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, res > thresh);
            if (lane == 0 && x / 32 < wpr_bits) bits[((long long)f * h + y) * wpr_bits + x / 32] = b;
        }
    }
}

// The same filter for radius 3 (d = 7, what the reference calls): the 29 taps are unrolled at compile time in
// OpenCV's order (offsets become immediates, the tap index a constant) and pixels at least 3 away from every
// border skip the reflection.  Same operations in the same order as k_bilateral: bit-identical results.
__global__ void __launch_bounds__(256)
k_bilateral_r3(const uint8_t* __restrict__ in, int n, int h, int w, const BilateralLut* __restrict__ lutp, int reverse,
               uint8_t* __restrict__ out, int thresh, uint32_t* __restrict__ bits, int wpr_bits) {
    __shared__ float s_color[256];
    __shared__ float s_space[32];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_color[i] = lutp->color[i];
    if (threadIdx.x < 32) s_space[threadIdx.x] = lutp->space[threadIdx.x];
    __syncthreads();
    const int wpr = (w + 31) / 32;
    const int lane = threadIdx.x & 31;
    const long long nwords = (long long)n * h * wpr;
    for (long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; word < nwords;
         word += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int f = (int)(word / ((long long)h * wpr));
        const long long rem = word - (long long)f * h * wpr;
        const int y = (int)(rem / wpr), x = (int)(rem - (long long)y * wpr) * 32 + lane;
        const uint8_t* img = in + (long long)(reverse ? n - 1 - f : f) * h * w;
        int res = 0;
        if (x < w) {
            const uint8_t* ctr = img + (long long)y * w + x;
            const int c = *ctr;
            float sum = 0.f, wsum = 0.f;
            if (y >= 3 && y < h - 3 && x >= 3 && x < w - 3) {
                int k = 0;
#pragma unroll
                for (int i = -3; i <= 3; ++i)
#pragma unroll
                    for (int j = -3; j <= 3; ++j) {
                        if (i * i + j * j > 9) continue;
                        const int v = ctr[i * w + j];
                        const float wk = __fmul_rn(s_space[k], s_color[abs(v - c)]);
                        wsum = __fadd_rn(wsum, wk);
                        sum = __fadd_rn(sum, __fmul_rn((float)v, wk));
                        ++k;
                    }
            } else {
                int k = 0;
#pragma unroll
                for (int i = -3; i <= 3; ++i)
#pragma unroll
                    for (int j = -3; j <= 3; ++j) {
                        if (i * i + j * j > 9) continue;
                        const int v = img[(long long)reflect101(y + i, h) * w + reflect101(x + j, w)];
                        const float wk = __fmul_rn(s_space[k], s_color[abs(v - c)]);
                        wsum = __fadd_rn(wsum, wk);
                        sum = __fadd_rn(sum, __fmul_rn((float)v, wk));
                        ++k;
                    }
            }
            res = __float2int_rn(__fdiv_rn(sum, wsum));
            res = min(max(res, 0), 255);
            if (out) out[((long long)f * h + y) * w + x] = (uint8_t)res;
        }
        if (bits) {
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, res > thresh);
            if (lane == 0 && x / 32 < wpr_bits) bits[((long long)f * h + y) * wpr_bits + x / 32] = b;
        }
    }
}

}  // namespace

// d = 7 (radius 3, 29 taps), rows that are multiples of four pixels: a thread filters FOUR adjacent pixels.
// The one-pixel kernel above issues 29 byte loads + 58 shared-memory reads per pixel and is bound by the
// load/store pipe (1.04 ms per 21 frames of 1080p); here a window row is three aligned 32-bit loads for the four
// pixels (21 loads per thread instead of 116), the spatial weights sit in registers, and only the colour-weight
// look-ups still go to shared memory.  Per pixel the taps are visited in the same order with the same float32
// operations as k_bilateral: identical results.  block = (32, 8): a warp covers 128 pixels of a row, a CTA eight
// adjacent rows (their windows overlap in L1); grid = (ceil(w / 128), ceil(h / 8), frames): no index divisions.
__global__ void __launch_bounds__(256)
k_bilateral_r3x4(const uint8_t* __restrict__ in, int n, int h, int w, const BilateralLut* __restrict__ lutp, int reverse,
                 uint8_t* __restrict__ out, int thresh, uint32_t* __restrict__ bits, int wpr_bits) {
    __shared__ float s_color[256];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    s_color[tid] = lutp->color[tid];
    float sp[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) sp[k] = __ldg(&lutp->space[k]);
    __syncthreads();
    const int lane = threadIdx.x;
    const int x4 = (blockIdx.x * 32 + lane) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int f = blockIdx.z;
    if (y >= h) return;                                        // warp-uniform
    const uint8_t* img = in + (long long)(reverse ? n - 1 - f : f) * h * w;
    uint32_t res4 = 0;                                         // the four results, one per byte
    if (x4 < w) {
        if (y >= 3 && y < h - 3 && x4 >= 4 && x4 + 8 <= w) {
            const uint8_t* row = img + (long long)y * w + x4;
            const uint32_t cw = __ldg(reinterpret_cast<const uint32_t*>(row));
            int c[4];
            float sum[4], wsum[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                c[p] = (cw >> (8 * p)) & 0xFFu;
                sum[p] = wsum[p] = 0.f;
            }
            int k0 = 0;                                        // tap index of the row's first tap
#pragma unroll
            for (int i = -3; i <= 3; ++i) {
                const uint32_t* rp = reinterpret_cast<const uint32_t*>(row + (long long)i * w);
                const uint32_t wd[3] = {__ldg(rp - 1), __ldg(rp), __ldg(rp + 1)};   // pixels x4 - 4 .. x4 + 7
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    int k = k0;
#pragma unroll
                    for (int j = -3; j <= 3; ++j) {
                        if (i * i + j * j > 9) continue;
                        const int b = 4 + p + j;               // byte of the 12-byte window (compile-time)
                        const int v = (wd[b >> 2] >> (8 * (b & 3))) & 0xFFu;
                        const float wk = __fmul_rn(sp[k], s_color[abs(v - c[p])]);
                        wsum[p] = __fadd_rn(wsum[p], wk);
                        sum[p] = __fadd_rn(sum[p], __fmul_rn((float)v, wk));
                        ++k;
                    }
                }
#pragma unroll
                for (int j = -3; j <= 3; ++j)
                    if (i * i + j * j <= 9) ++k0;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                int r = __float2int_rn(__fdiv_rn(sum[p], wsum[p]));
                r = min(max(r, 0), 255);
                res4 |= (uint32_t)r << (8 * p);
            }
        } else {
            for (int p = 0; p < 4 && x4 + p < w; ++p) {        // frame border: reflected taps, one pixel at a time
                const int x = x4 + p;
                const int c = img[(long long)y * w + x];
                float sum = 0.f, wsum = 0.f;
                int k = 0;
                for (int i = -3; i <= 3; ++i)
                    for (int j = -3; j <= 3; ++j) {
                        if (i * i + j * j > 9) continue;
                        const int v = img[(long long)reflect101(y + i, h) * w + reflect101(x + j, w)];
                        const float wk = __fmul_rn(__ldg(&lutp->space[k]), s_color[abs(v - c)]);
                        wsum = __fadd_rn(wsum, wk);
                        sum = __fadd_rn(sum, __fmul_rn((float)v, wk));
                        ++k;
                    }
                int r = __float2int_rn(__fdiv_rn(sum, wsum));
                r = min(max(r, 0), 255);
                res4 |= (uint32_t)r << (8 * p);
            }
        }
        if (out) *reinterpret_cast<uint32_t*>(out + ((long long)f * h + y) * w + x4) = res4;
    }
    if (bits) {
        // lane -> nibble (lane & 7) of word lane / 8 of the warp's 128 pixels; OR over the eight lanes of a word
        uint32_t nib = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) nib |= ((int)((res4 >> (8 * p)) & 0xFFu) > thresh ? 1u : 0u) << p;
        uint32_t word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
        const int wi = x4 >> 5;
        if ((lane & 7) == 0 && x4 < w && wi < wpr_bits) bits[((long long)f * h + y) * wpr_bits + wi] = word;
    }
}

cudaError_t launch_bilateral(cudaStream_t s, const uint8_t* in, int n, int h, int w, const BilateralLut* d_lut,
                             int reverse, uint8_t* out, int thresh, uint32_t* bits, int wpr_bits, int radius) {
    const long long nwords = (long long)n * h * ((w + 31) / 32);
    const int grid = (int)std::min<long long>((nwords * 32 + 255) / 256, 148 * 32);
    static const bool x4 = [] { const char* e = getenv("SWB_BILATERAL_X4"); return !(e && e[0] == '0'); }();
    if (radius == 3 && h >= 7 && w >= 16 && (w & 3) == 0 && x4 && h <= 65535 * 8 && n <= 65535 &&
        (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (out == nullptr || (reinterpret_cast<uintptr_t>(out) & 3) == 0))
        k_bilateral_r3x4<<<dim3((w + 127) / 128, (h + 7) / 8, n), dim3(32, 8), 0, s>>>(in, n, h, w, d_lut, reverse, out, thresh,
                                                                                  bits, wpr_bits);
    else if (radius == 3 && h >= 7 && w >= 7)
        k_bilateral_r3<<<grid, 256, 0, s>>>(in, n, h, w, d_lut, reverse, out, thresh, bits, wpr_bits);
    else
        k_bilateral<<<grid, 256, 0, s>>>(in, n, h, w, d_lut, reverse, out, thresh, bits, wpr_bits);
    return cudaGetLastError();
}

}  // namespace swb
