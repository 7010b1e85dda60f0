// K3..K6 — 8-connected component labelling on the bit-packed mask, fused with
// the regionprops reduction.
//
// Replaces cc_labeling (image_filtering.py:325-329: cv2.connectedComponents
// called with `connectivity` in the `labels` slot, hence 8-connectivity) and
// get_segment_properties (image_filtering.py:332-335: skimage regionprops;
// label / area / bbox / centroid are what swiftwatcher consumes).
//
// Label numbering is OpenCV's: components are numbered 1..n by ascending
// minimum 2x2-block raster index.  The union-find therefore runs over 2x2
// pixel blocks (all foreground pixels of a block are mutually 8-adjacent),
// links larger roots under smaller ones (so a root IS the minimum block of
// its component) and the final label is 1 + the rank of the root among all
// roots in raster order — a bit count over root flags, no sort.
//
// Foreground is sparse (birds), so every array indexed by block is allocated
// dense but only touched where the mask is set; the only dense traffic is the
// final label image write.
//
// Tiled path (frames up to 4096 pixels wide; one CTA = full-width tile of 8..256
// block rows held in shared memory):
//   ccl_local     tile-local union-find in shared memory (init, merge, compress);
//                 writes parent[b] = tile-local root and one regionprops partial
//                 sum per tile-local component (shared-memory atomics)
//   ccl_boundary  global unions between the first block row of a tile and the
//                 last block row of the tile above (atomicMin union-find)
// Wide path (wider frames): ccl_init + ccl_merge do everything globally and
//   ccl_label also accumulates the regionprops with global atomics.
// Then, both paths:
//   ccl_roots     roots get parent = -(1 + rank in their block row); per-row counts
//   ccl_scan      per-frame exclusive scan of row counts -> segments per frame
//   ccl_offsets   exclusive scan over frames -> row offsets of the segment table
//   seg_init      initialise the table rows (frame, label, empty bbox)
//   ccl_label     per block: root -> label (dense-allocated, sparsely written)
//   props_final   adds every partial sum to its component's table row
//   write_labels  dense int32 / uint8 label image from bits + block labels
#include "swb_internal.cuh"

namespace swb {

namespace {

constexpr uint32_t EVEN = 0x55555555u;

// One thread scans a group of 4 consecutive words (128 pixels) of a block row:
// rows 2by (A) and 2by+1 (B) with one 16-byte load each.  fbits rows are padded
// to wpr4 (a multiple of 4 words) with zero words.
struct Group {
    uint32_t A[4], B[4];
    __device__ __forceinline__ uint32_t any() const {
        return A[0] | A[1] | A[2] | A[3] | B[0] | B[1] | B[2] | B[3];
    }
};

__device__ __forceinline__ void load_group(Group& gr, const uint32_t* fb, const Geom& g, int by, int q) {
    const uint4* p = reinterpret_cast<const uint4*>(fb + (long long)(2 * by) * g.wpr4) + q;
    const uint4 a = __ldg(p);
    uint4 b = make_uint4(0u, 0u, 0u, 0u);
    if (2 * by + 1 < g.h) b = __ldg(p + (g.wpr4 >> 2));
    gr.A[0] = a.x; gr.A[1] = a.y; gr.A[2] = a.z; gr.A[3] = a.w;
    gr.B[0] = b.x; gr.B[1] = b.y; gr.B[2] = b.z; gr.B[3] = b.w;
}

// thread <-> (frame, block row, group): blockDim = (bx, 256 / bx), grid = (ceil(Q / bx), ceil(BH / by), T)
__device__ __forceinline__ bool my_group(const Geom& g, int& f, int& by, int& q) {
    q = blockIdx.x * blockDim.x + threadIdx.x;
    by = blockIdx.y * blockDim.y + threadIdx.y;
    f = blockIdx.z;
    return q < (g.wpr4 >> 2) && by < g.BH;
}

__global__ void __launch_bounds__(256)
k_ccl_init(const uint32_t* __restrict__ fbits, Geom g, int* __restrict__ parent) {
    int f, by, q;
    if (!my_group(g, f, by, q)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    Group gr;
    load_group(gr, fb, g, by, q);
    if (!gr.any()) return;
    int* par = parent + (long long)f * g.BH * g.BW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t P = gr.A[i] | gr.B[i];
        if (!P) continue;
        uint32_t O = (P | (P >> 1)) & EVEN;         // bit 2k: block k occupied
        const uint32_t H = P & (P << 1) & EVEN;     // bit 2k: block k touches block k-1 (same word)
        const int base = by * g.BW + 16 * (4 * q + i);
        int start = 0;
        while (O) {
            const int b = __ffs(O) - 1;
            O &= O - 1;
            const int k = b >> 1;
            if (!((H >> b) & 1u)) start = k;
            par[base + k] = base + start;
        }
    }
}

__device__ __forceinline__ int find_root_volatile(int* par, int x) {
    int p = ((volatile int*)par)[x];
    while (p != x) {
        x = p;
        p = ((volatile int*)par)[x];
    }
    return x;
}

// Lock-free union; the larger root is linked under the smaller one, so the root
// of a finished component is its minimum block index.
__device__ void unite(int* par, int a, int b) {
    while (true) {
        a = find_root_volatile(par, a);
        b = find_root_volatile(par, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // a > b
        const int old = atomicMin(&par[a], b);
        if (old == a) return;
        a = old;
    }
}

// Unions between the blocks of group (by, q) and the block row above: only the
// bottom pixel row (2by - 1) of that row matters.
__device__ void vertical_links(const Group& gr, const uint32_t* fb, const Geom& g, int by, int q, int* par) {
    const uint32_t* up = fb + (long long)(2 * by - 1) * g.wpr4 + 4 * q;
    const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(up));
    uint32_t U[6];
    U[1] = u4.x; U[2] = u4.y; U[3] = u4.z; U[4] = u4.w;
    if ((U[1] | U[2] | U[3] | U[4]) == 0u) {
        // only the diagonal neighbours outside the group can still touch
        U[0] = (q > 0 && (gr.A[0] & 1u)) ? up[-1] : 0u;
        U[5] = (4 * q + 4 < g.wpr4 && (gr.A[3] >> 31)) ? up[4] : 0u;
        if ((U[0] | U[5]) == 0u) return;
    } else {
        U[0] = q > 0 ? up[-1] : 0u;
        U[5] = 4 * q + 4 < g.wpr4 ? up[4] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t A = gr.A[i];
        if (!A) continue;
        const uint32_t Bp = U[i + 1];
        const uint32_t Bp_l = __funnelshift_l(U[i], Bp, 1);      // bit x = Bp[x-1]
        const uint32_t Bp_r = __funnelshift_r(Bp, U[i + 2], 1);  // bit x = Bp[x+1]
        const uint32_t P = A | gr.B[i];
        uint32_t UP = (Bp | (Bp >> 1)) & (A | (A >> 1)) & EVEN;  // block k <-> up block k
        uint32_t UL = Bp_l & A & EVEN;                           // pixel 2k   <-> up pixel 2k-1
        uint32_t UR = ((Bp_r & A) >> 1) & EVEN;                  // pixel 2k+1 <-> up pixel 2k+2
        // drop links implied by others
        const uint32_t H = P & (P << 1) & EVEN;
        UL &= ~(UP & Bp);                    // up blocks k-1,k already joined through Bp[2k-1],Bp[2k]
        UR &= ~(UP & (Bp >> 1));             // up blocks k,k+1 already joined through Bp[2k+1],Bp[2k+2]
        UP &= ~(H & (UP << 2) & Bp_l & Bp);  // cur k-1~k, up k-1~k and cur k-1 ~ up k-1
        const int base = by * g.BW + 16 * (4 * q + i);
        const int upbase = base - g.BW;
        while (UP) {
            const int k = (__ffs(UP) - 1) >> 1;
            UP &= UP - 1;
            unite(par, base + k, upbase + k);
        }
        while (UL) {
            const int k = (__ffs(UL) - 1) >> 1;
            UL &= UL - 1;
            unite(par, base + k, upbase + k - 1);
        }
        while (UR) {
            const int k = (__ffs(UR) - 1) >> 1;
            UR &= UR - 1;
            unite(par, base + k, upbase + k + 1);
        }
    }
}


__global__ void __launch_bounds__(256)
k_ccl_merge(const uint32_t* __restrict__ fbits, Geom g, int* parent) {
    int f, by, q;
    if (!my_group(g, f, by, q)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    Group gr;
    load_group(gr, fb, g, by, q);
    if (!gr.any()) return;
    int* par = parent + (long long)f * g.BH * g.BW;

    // horizontal links across word boundaries
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t P = gr.A[i] | gr.B[i];
        if (!(P & 1u)) continue;
        uint32_t Pl;
        if (i > 0) {
            Pl = gr.A[i > 0 ? i - 1 : 0] | gr.B[i > 0 ? i - 1 : 0];
        } else {
            if (q == 0) continue;
            const uint32_t* r0 = fb + (long long)(2 * by) * g.wpr4 + 4 * q - 1;
            Pl = r0[0] | ((2 * by + 1 < g.h) ? r0[g.wpr4] : 0u);
        }
        const int base = by * g.BW + 16 * (4 * q + i);
        if (Pl >> 31) unite(par, base, base - 1);
    }
    if (by == 0 || (gr.A[0] | gr.A[1] | gr.A[2] | gr.A[3]) == 0u) return;
    vertical_links(gr, fb, g, by, q, par);
}

// warp per (frame, block row): every root (parent[b] == b) gets parent[b] =
// -(1 + its rank among the roots of this block row); rowcount = roots in the row.
__global__ void __launch_bounds__(256)
k_ccl_roots(const uint32_t* __restrict__ fbits, int T, Geom g, int* __restrict__ parent,
            uint32_t* __restrict__ rowcount) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)T * g.BH) return;
    const int f = (int)(wid / g.BH);
    const int by = (int)(wid - (long long)f * g.BH);
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    int* par = parent + (long long)f * g.BH * g.BW;
    const int Q = g.wpr4 >> 2;
    uint32_t running = 0;
    for (int q0 = 0; q0 < Q; q0 += 32) {
        const int q = q0 + lane;
        uint32_t RB[4] = {0u, 0u, 0u, 0u};
        if (q < Q) {
            Group gr;
            load_group(gr, fb, g, by, q);
            if (gr.any()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t P = gr.A[i] | gr.B[i];
                    uint32_t o = (P | (P >> 1)) & EVEN;
                    const int base = by * g.BW + 16 * (4 * q + i);
                    while (o) {
                        const int b = __ffs(o) - 1;
                        o &= o - 1;
                        if (par[base + (b >> 1)] == base + (b >> 1)) RB[i] |= 1u << b;
                    }
                }
            }
        }
        const uint32_t cnt = __popc(RB[0]) + __popc(RB[1]) + __popc(RB[2]) + __popc(RB[3]);
        const uint32_t any = __ballot_sync(0xFFFFFFFFu, cnt != 0u);
        if (any == 0u) continue;                      // warp-uniform
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        uint32_t rank = running + incl - cnt;
        if (cnt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t r = RB[i];
                const int base = by * g.BW + 16 * (4 * q + i);
                while (r) {
                    const int b = __ffs(r) - 1;
                    r &= r - 1;
                    par[base + (b >> 1)] = -(int)(1u + rank);
                    ++rank;
                }
            }
        }
        running += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) rowcount[(long long)f * g.BH + by] = running;
}

// warp per frame: rowcount -> exclusive row base; nseg[f] = total
__global__ void __launch_bounds__(256)
k_ccl_scan(int T, Geom g, uint32_t* __restrict__ rowcount, int32_t* __restrict__ nseg) {
    const int lane = threadIdx.x & 31;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= T) return;
    uint32_t* rc = rowcount + (long long)f * g.BH;
    uint32_t running = 0;
    for (int r0 = 0; r0 < g.BH; r0 += 32) {
        const int r = r0 + lane;
        const uint32_t cnt = r < g.BH ? rc[r] : 0u;
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (r < g.BH) rc[r] = running + incl - cnt;
        running += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) nseg[f] = (int32_t)running;
}

// one CTA: exclusive scan of nseg over frames
__global__ void __launch_bounds__(1024)
k_ccl_offsets(int T, const int32_t* __restrict__ nseg, int32_t* __restrict__ segoff, int cap_rows,
              int32_t* __restrict__ overflow) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int f0 = 0; f0 < T; f0 += 1024) {
        const int f = f0 + tid;
        const int32_t cnt = f < T ? nseg[f] : 0;
        int32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warp_tot[lane];
            int32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += v;
            }
            warp_tot[lane] = wi - w;   // exclusive
        }
        __syncthreads();
        const int32_t base = carry + warp_tot[warp];
        if (f < T) segoff[f] = base + incl - cnt;
        __syncthreads();
        if (tid == 1023) carry = base + incl;
        __syncthreads();
    }
    if (tid == 0) {
        segoff[T] = carry;
        if (carry > cap_rows) *overflow = 1;
    }
}

__global__ void __launch_bounds__(256)
k_seg_init(int T, const int32_t* __restrict__ nseg, const int32_t* __restrict__ segoff,
           swb_segment* __restrict__ rows, int cap_rows) {
    const int f = blockIdx.y;
    const int n = nseg[f];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const long long r = (long long)segoff[f] + i;
        if (r >= cap_rows) return;
        swb_segment s;
        s.frame = f;
        s.label = i + 1;
        s.area = 0;
        s.bbox[0] = 0x7FFFFFFF;
        s.bbox[1] = 0x7FFFFFFF;
        s.bbox[2] = 0;
        s.bbox[3] = 0;
        s.reserved = 0;
        s.sum_row = 0;
        s.sum_col = 0;
        rows[r] = s;
    }
}

struct Acc {
    int label;        // 0 = empty
    int area;
    int minr, minc, maxr, maxc;   // inclusive max here; +1 applied at flush
    long long sr, sc;
};

__device__ __forceinline__ void flush(const Acc& a, swb_segment* rows, long long off, int cap_rows) {
    if (a.label == 0) return;
    const long long r = off + a.label - 1;
    if (r >= cap_rows) return;
    swb_segment* s = rows + r;
    atomicAdd(&s->area, a.area);
    atomicMin(&s->bbox[0], a.minr);
    atomicMin(&s->bbox[1], a.minc);
    atomicMax(&s->bbox[2], a.maxr + 1);
    atomicMax(&s->bbox[3], a.maxc + 1);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_row), (unsigned long long)a.sr);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_col), (unsigned long long)a.sc);
}

// per occupied block: follow parents to the root (negative entry = -(1 + rank in
// its row)), label = rowbase[root row] + 1 + rank; regionprops partial sums of
// consecutive blocks with the same label are flushed with one set of atomics.
template <bool PROPS>
__global__ void __launch_bounds__(256)
k_ccl_label(const uint32_t* __restrict__ fbits, Geom g, const int* __restrict__ parent,
            const uint32_t* __restrict__ rowbase, const int32_t* __restrict__ segoff,
            int* __restrict__ blocklabel, swb_segment* rows, int cap_rows) {
    int f, by, q;
    if (!my_group(g, f, by, q)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    Group gr;
    load_group(gr, fb, g, by, q);
    if (!gr.any()) return;
    const int* par = parent + (long long)f * g.BH * g.BW;
    int* bl = blocklabel + (long long)f * g.BH * g.BW;
    const uint32_t* rbase = rowbase + (long long)f * g.BH;
    const long long off = segoff[f];

    Acc acc;
    acc.label = 0;
    int last_parent = 0x7FFFFFFF, last_label = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t A = gr.A[i], B = gr.B[i];
        const uint32_t P = A | B;
        if (!P) continue;
        const int base = by * g.BW + 16 * (4 * q + i);
        uint32_t O = (P | (P >> 1)) & EVEN;
        while (O) {
            const int b = __ffs(O) - 1;
            O &= O - 1;
            const int k = b >> 1;
            const int p0 = par[base + k];
            int label;
            if (p0 == last_parent) {
                label = last_label;
            } else {
                int x = base + k, p = p0;
                while (p >= 0) {
                    x = p;
                    p = par[x];
                }
                label = (int)rbase[x / g.BW] - p;
                last_parent = p0;
                last_label = label;
            }
            bl[base + k] = label;
            if constexpr (PROPS) {
                const int a0 = (A >> b) & 1, a1 = (A >> (b + 1)) & 1;
                const int c0 = (B >> b) & 1, c1 = (B >> (b + 1)) & 1;
                const int y0 = 2 * by, x0 = 32 * (4 * q + i) + b;
                const int area = a0 + a1 + c0 + c1;
                const int minr = (a0 | a1) ? y0 : y0 + 1;
                const int maxr = (c0 | c1) ? y0 + 1 : y0;
                const int minc = (a0 | c0) ? x0 : x0 + 1;
                const int maxc = (a1 | c1) ? x0 + 1 : x0;
                const long long sr = (long long)(a0 + a1) * y0 + (long long)(c0 + c1) * (y0 + 1);
                const long long sc = (long long)(a0 + c0) * x0 + (long long)(a1 + c1) * (x0 + 1);
                if (label != acc.label) {
                    flush(acc, rows, off, cap_rows);
                    acc.label = label;
                    acc.area = area;
                    acc.minr = minr; acc.maxr = maxr; acc.minc = minc; acc.maxc = maxc;
                    acc.sr = sr; acc.sc = sc;
                } else {
                    acc.area += area;
                    acc.minr = min(acc.minr, minr); acc.maxr = max(acc.maxr, maxr);
                    acc.minc = min(acc.minc, minc); acc.maxc = max(acc.maxc, maxc);
                    acc.sr += sr; acc.sc += sc;
                }
            }
        }
    }
    if constexpr (PROPS) flush(acc, rows, off, cap_rows);
}

// ------------------------------------------------------------------------------------
// Tiled path
// ------------------------------------------------------------------------------------
constexpr int MAXR = 256;          // tile-local components with a shared-memory accumulator
constexpr uint32_t TAG = 0x8000u;  // sp[] entry of a claimed root: TAG | slot
constexpr uint32_t NOSLOT = 0x7FFFu;

__device__ __forceinline__ uint32_t ld_s(const unsigned short* sp, int i) {
    return ((const volatile unsigned short*)sp)[i];
}
__device__ __forceinline__ int find_s(const unsigned short* sp, int x) {
    int p = (int)ld_s(sp, x);
    while (p != x) {
        x = p;
        p = (int)ld_s(sp, x);
    }
    return x;
}
// union in shared memory: link the larger root under the smaller with a 16-bit CAS
__device__ void unite_s(unsigned short* sp, int a, int b) {
    while (true) {
        a = find_s(sp, a);
        b = find_s(sp, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // a > b
        const unsigned short old = atomicCAS(&sp[a], (unsigned short)a, (unsigned short)b);
        if (old == (unsigned short)a) return;
    }
}

struct PartAcc {
    uint32_t slot;   // NOSLOT + 1 = empty
    int root;
    uint32_t area, sr, sc;
    int minr, minc, maxr, maxc;
};

__device__ __forceinline__ void emit_partial(Partial* parts, int idx, int f, int root, uint32_t area, int minr,
                                             int minc, int maxr, int maxc, uint32_t sr, uint32_t sc) {
    Partial p;
    p.frame = f; p.root = root; p.area = (int)area;
    p.minr = minr; p.minc = minc; p.maxr = maxr; p.maxc = maxc;
    p.sr = sr; p.sc = sc; p.pad[0] = p.pad[1] = p.pad[2] = 0;
    parts[idx] = p;
}

// BX = groups (of 4 words) per tile row, a power of two >= Q = wpr4 / 4; BY = 256 / BX block rows.
template <int BX>
__global__ void __launch_bounds__(256)
k_ccl_local(const uint32_t* __restrict__ fbits, Geom g, int* __restrict__ parent, Partial* __restrict__ parts,
            int* __restrict__ pcount, int cap_parts, int32_t* __restrict__ overflow) {
    constexpr int BY = 256 / BX;
    constexpr int ROWB = BX * 64;            // blocks per tile row
    constexpr int SBW = BX * 4 + 2;          // bottom-row words + one halo word each side
    __shared__ unsigned short sp[256 * 64];  // tile-local parents (32 KB)
    __shared__ uint32_t sB[BY * SBW];
    __shared__ uint32_t sPw3[256];
    __shared__ uint32_t st_area[MAXR], st_sr[MAXR], st_sc[MAXR];
    __shared__ int st_minr[MAXR], st_minc[MAXR], st_maxr[MAXR], st_maxc[MAXR];
    __shared__ unsigned short sroot[MAXR];
    __shared__ int s_n, s_base;

    const int tid = threadIdx.x;
    const int tx = tid % BX, ty = tid / BX;
    const int f = blockIdx.y;
    const int by0 = blockIdx.x * BY;
    const int by = by0 + ty;
    const int Q = g.wpr4 >> 2;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;

    Group gr;
#pragma unroll
    for (int i = 0; i < 4; ++i) gr.A[i] = gr.B[i] = 0u;
    if (tx < Q && by < g.BH) load_group(gr, fb, g, by, tx);
    const uint32_t any = gr.any();
#pragma unroll
    for (int i = 0; i < 4; ++i) sB[ty * SBW + 1 + 4 * tx + i] = gr.B[i];
    if (tx == 0) sB[ty * SBW] = 0u;
    if (tx == BX - 1) sB[ty * SBW + SBW - 1] = 0u;
    sPw3[tid] = gr.A[3] | gr.B[3];
    if (tid == 0) s_n = 0;

    // ---- phase 1: parent = first block of the horizontal run (inside the thread's 128 pixels)
    const int l0 = tid * 64;
    if (any) {
        int start = 0;
        uint32_t Pprev = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t P = gr.A[i] | gr.B[i];
            uint32_t O = (P | (P >> 1)) & EVEN;
            const uint32_t H = (P & ((P << 1) | (Pprev >> 31))) & EVEN;   // bit 2k: block k touches block k-1
            while (O) {
                const int b = __ffs(O) - 1;
                O &= O - 1;
                const int pos = i * 16 + (b >> 1);
                if (!((H >> b) & 1u)) start = pos;
                sp[l0 + pos] = (unsigned short)(l0 + start);
            }
            Pprev = P;
        }
    }
    if (!__syncthreads_or((int)any)) return;    // empty tile: nothing to write anywhere

    // ---- phase 2: unions with the left neighbour thread and with the block row above (same tile)
    if (any) {
        if (((gr.A[0] | gr.B[0]) & 1u) && tx > 0 && (sPw3[tid - 1] >> 31)) unite_s(sp, l0, l0 - 1);
        if (ty > 0) {
            const uint32_t* up = &sB[(ty - 1) * SBW + 4 * tx];   // up[0] = word 4tx-1 of the row above
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t A = gr.A[i];
                if (!A) continue;
                const uint32_t Bp = up[i + 1];
                const uint32_t Bp_l = __funnelshift_l(up[i], Bp, 1);
                const uint32_t Bp_r = __funnelshift_r(Bp, up[i + 2], 1);
                const uint32_t P = A | gr.B[i];
                uint32_t UP = (Bp | (Bp >> 1)) & (A | (A >> 1)) & EVEN;
                uint32_t UL = Bp_l & A & EVEN;
                uint32_t UR = ((Bp_r & A) >> 1) & EVEN;
                const uint32_t H = P & (P << 1) & EVEN;
                UL &= ~(UP & Bp);
                UR &= ~(UP & (Bp >> 1));
                UP &= ~(H & (UP << 2) & Bp_l & Bp);
                const int base = l0 + i * 16;
                while (UP) {
                    const int k = (__ffs(UP) - 1) >> 1;
                    UP &= UP - 1;
                    unite_s(sp, base + k, base + k - ROWB);
                }
                while (UL) {
                    const int k = (__ffs(UL) - 1) >> 1;
                    UL &= UL - 1;
                    unite_s(sp, base + k, base + k - ROWB - 1);
                }
                while (UR) {
                    const int k = (__ffs(UR) - 1) >> 1;
                    UR &= UR - 1;
                    unite_s(sp, base + k, base + k - ROWB + 1);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 3a: full path compression (every entry points at its root)
    if (any) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t P = gr.A[i] | gr.B[i];
            uint32_t O = (P | (P >> 1)) & EVEN;
            while (O) {
                const int pos = i * 16 + ((__ffs(O) - 1) >> 1);
                O &= O - 1;
                const int r = find_s(sp, l0 + pos);
                sp[l0 + pos] = (unsigned short)r;
            }
        }
    }
    __syncthreads();

    // ---- phase 3b: roots claim an accumulator slot
    if (any) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t P = gr.A[i] | gr.B[i];
            uint32_t O = (P | (P >> 1)) & EVEN;
            while (O) {
                const int pos = i * 16 + ((__ffs(O) - 1) >> 1);
                O &= O - 1;
                const int l = l0 + pos;
                if ((int)sp[l] == l) {
                    const int slot = atomicAdd(&s_n, 1);
                    if (slot < MAXR) {
                        sroot[slot] = (unsigned short)l;
                        st_area[slot] = 0u; st_sr[slot] = 0u; st_sc[slot] = 0u;
                        st_minr[slot] = 0x7FFFFFFF; st_minc[slot] = 0x7FFFFFFF;
                        st_maxr[slot] = -1; st_maxc[slot] = -1;
                        sp[l] = (unsigned short)(TAG | (uint32_t)slot);
                    } else {
                        sp[l] = (unsigned short)(TAG | NOSLOT);
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 3c: global parent = tile-local root; regionprops partial sums per root
    if (any) {
        int* par = parent + (long long)f * g.BH * g.BW;
        const int grow = by * g.BW;              // global id of this block row's block 0
        PartAcc acc;
        acc.slot = NOSLOT + 1;
        auto flush_acc = [&]() {
            if (acc.slot == NOSLOT + 1) return;
            if (acc.slot < (uint32_t)MAXR) {
                atomicAdd(&st_area[acc.slot], acc.area);
                atomicAdd(&st_sr[acc.slot], acc.sr);
                atomicAdd(&st_sc[acc.slot], acc.sc);
                atomicMin(&st_minr[acc.slot], acc.minr);
                atomicMin(&st_minc[acc.slot], acc.minc);
                atomicMax(&st_maxr[acc.slot], acc.maxr);
                atomicMax(&st_maxc[acc.slot], acc.maxc);
            } else {   // more than MAXR components in this tile: one partial per run, straight to global
                const int idx = atomicAdd(pcount, 1);
                if (idx < cap_parts) emit_partial(parts, idx, f, acc.root, acc.area, acc.minr, acc.minc, acc.maxr,
                                                  acc.maxc, acc.sr, acc.sc);
                else *overflow = 1;
            }
        };
        int last_e = -1;
        uint32_t last_slot = 0;
        int last_root = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t A = gr.A[i], B = gr.B[i];
            const uint32_t P = A | B;
            uint32_t O = (P | (P >> 1)) & EVEN;
            while (O) {
                const int b = __ffs(O) - 1;
                O &= O - 1;
                const int pos = i * 16 + (b >> 1);
                const int l = l0 + pos;
                const int e = (int)sp[l];
                uint32_t slot;
                int rl;                               // tile-local index of the root
                if (e == last_e && !(e & TAG)) {
                    slot = last_slot; rl = last_root;
                } else if (e & TAG) {
                    slot = (uint32_t)e & NOSLOT; rl = l;
                } else {
                    slot = (uint32_t)sp[e] & NOSLOT; rl = e;
                    last_e = e; last_slot = slot; last_root = rl;
                }
                // tile-local index -> global block id: rows are contiguous runs of ROWB blocks
                const int rgid = (by0 + rl / ROWB) * g.BW + (rl % ROWB);
                par[grow + tx * 64 + pos] = rgid;
                const int a0 = (A >> b) & 1, a1 = (A >> (b + 1)) & 1;
                const int c0 = (B >> b) & 1, c1 = (B >> (b + 1)) & 1;
                const int y0 = 2 * by, x0 = 2 * (tx * 64 + pos);
                const uint32_t area = a0 + a1 + c0 + c1;
                const int minr = (a0 | a1) ? y0 : y0 + 1;
                const int maxr = (c0 | c1) ? y0 + 1 : y0;
                const int minc = (a0 | c0) ? x0 : x0 + 1;
                const int maxc = (a1 | c1) ? x0 + 1 : x0;
                const uint32_t sr = (uint32_t)((a0 + a1) * y0 + (c0 + c1) * (y0 + 1));
                const uint32_t sc = (uint32_t)((a0 + c0) * x0 + (a1 + c1) * (x0 + 1));
                if (slot != acc.slot || (slot == NOSLOT && rgid != acc.root)) {
                    flush_acc();
                    acc.slot = slot; acc.root = rgid;
                    acc.area = area; acc.sr = sr; acc.sc = sc;
                    acc.minr = minr; acc.maxr = maxr; acc.minc = minc; acc.maxc = maxc;
                } else {
                    acc.area += area; acc.sr += sr; acc.sc += sc;
                    acc.minr = min(acc.minr, minr); acc.maxr = max(acc.maxr, maxr);
                    acc.minc = min(acc.minc, minc); acc.maxc = max(acc.maxc, maxc);
                }
            }
        }
        flush_acc();
    }
    __syncthreads();

    // ---- phase 3d: one partial per tile-local component -> global list
    const int n = min(s_n, MAXR);
    if (tid == 0) s_base = atomicAdd(pcount, n);
    __syncthreads();
    for (int s = tid; s < n; s += 256) {
        const int idx = s_base + s;
        if (idx < cap_parts) {
            const int rl = (int)sroot[s];
            const int rgid = (by0 + rl / ROWB) * g.BW + (rl % ROWB);
            emit_partial(parts, idx, f, rgid, st_area[s], st_minr[s], st_minc[s], st_maxr[s], st_maxc[s], st_sr[s],
                         st_sc[s]);
        } else {
            *overflow = 1;
        }
    }
}

// Global unions between the first block row of every tile (tile_rows apart) and the
// row above it.  grid = (ceil(Q / 32), n_boundaries, T), 32 threads.
__global__ void __launch_bounds__(32)
k_ccl_boundary(const uint32_t* __restrict__ fbits, Geom g, int tile_rows, int* parent) {
    const int q = blockIdx.x * 32 + threadIdx.x;
    const int by = (blockIdx.y + 1) * tile_rows;
    const int f = blockIdx.z;
    if (q >= (g.wpr4 >> 2) || by >= g.BH) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(fb + (long long)(2 * by) * g.wpr4) + q);
    uint32_t Aw[4] = {a4.x, a4.y, a4.z, a4.w};
    if ((Aw[0] | Aw[1] | Aw[2] | Aw[3]) == 0u) return;
    Group gr;
    load_group(gr, fb, g, by, q);
    int* par = parent + (long long)f * g.BH * g.BW;
    vertical_links(gr, fb, g, by, q, par);
}

// every partial sum -> its component's row of the segment table
__global__ void __launch_bounds__(256)
k_props_final(const Partial* __restrict__ parts, const int* __restrict__ pcount, int cap_parts, Geom g,
              const int* __restrict__ parent, const uint32_t* __restrict__ rowbase,
              const int32_t* __restrict__ segoff, swb_segment* rows, int cap_rows) {
    const int n = min(*pcount, cap_parts);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Partial p = parts[i];
        const int* par = parent + (long long)p.frame * g.BH * g.BW;
        int x = p.root, v = par[x];
        while (v >= 0) {
            x = v;
            v = par[x];
        }
        const int label = (int)rowbase[(long long)p.frame * g.BH + x / g.BW] - v;
        const long long r = (long long)segoff[p.frame] + label - 1;
        if (r >= cap_rows) continue;
        swb_segment* s = rows + r;
        atomicAdd(&s->area, p.area);
        atomicMin(&s->bbox[0], p.minr);
        atomicMin(&s->bbox[1], p.minc);
        atomicMax(&s->bbox[2], p.maxr + 1);
        atomicMax(&s->bbox[3], p.maxc + 1);
        atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_row), (unsigned long long)p.sr);
        atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_col), (unsigned long long)p.sc);
    }
}

// Dense label image.  A thread owns 4 (int32) or 16 (uint8) consecutive pixels
// of one row and walks down RPT rows; stores are 16 bytes, contiguous per warp.
template <typename LT, int PX>
__global__ void __launch_bounds__(256)
k_write_labels(const uint32_t* __restrict__ fbits, Geom g, const int* __restrict__ blocklabel,
               LT* __restrict__ labels, int rows_per_thread) {
    const int f = blockIdx.z;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;   // pixel group within the row
    const int x = gx * PX;
    if (x >= g.mpitch) return;
    const int yb = blockIdx.y * rows_per_thread;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    const int* bl = blocklabel + (long long)f * g.BH * g.BW;
    LT* out = labels + (long long)f * g.h * g.mpitch;
    const int j = x >> 5, sh = x & 31;
    for (int yy = 0; yy < rows_per_thread; ++yy) {
        const int y = yb + yy;
        if (y >= g.h) break;
        const uint32_t bits = (fb[(long long)y * g.wpr4 + j] >> sh) & ((PX == 32) ? 0xFFFFFFFFu : ((1u << PX) - 1u));
        LT v[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) v[i] = 0;
        if (bits) {
            const int* blrow = bl + (long long)(y >> 1) * g.BW + (x >> 1);
#pragma unroll
            for (int i = 0; i < PX; i += 2) {
                if ((bits >> i) & 3u) {
                    const int lab = blrow[i >> 1];
                    if ((bits >> i) & 1u) v[i] = (LT)lab;
                    if ((bits >> (i + 1)) & 1u) v[i + 1] = (LT)lab;
                }
            }
        }
        uint4 o;
        if constexpr (sizeof(LT) == 4) {
            o = make_uint4((uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                w[q] = (uint32_t)v[4 * q] | ((uint32_t)v[4 * q + 1] << 8) | ((uint32_t)v[4 * q + 2] << 16) |
                       ((uint32_t)v[4 * q + 3] << 24);
            o = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __stcs(reinterpret_cast<uint4*>(out + (long long)y * g.mpitch + x), o);
    }
}

__global__ void __launch_bounds__(256)
k_pack_bits(const uint8_t* __restrict__ img, int h, int w, uint32_t* __restrict__ fbits, int wpr4) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= h * wpr4) return;
    const int y = idx / wpr4, j = idx - y * wpr4;
    uint32_t bits = 0;
    const uint8_t* row = img + (long long)y * w;
    for (int i = 0; i < 32; ++i) {
        const int x = 32 * j + i;
        if (x < w && row[x] != 0) bits |= 1u << i;
    }
    fbits[idx] = bits;
}

// extract_segment_images (image_filtering.py:338-369) as fixed crop x crop tiles:
// bbox grown symmetrically to crop x crop (floor/ceil split), shifted by the ROI
// origin, read from the full frame; pixels outside the frame are 0.  Segments
// whose bbox exceeds the crop are centre-cropped (the reference would hand the
// larger crop to the classifier's Resize).
__global__ void __launch_bounds__(256)
k_gather_crops(const uint8_t* __restrict__ frames, long long frame_stride, long long pitch, int channels,
               int frame_h, int frame_w, int roi_x0, int roi_y0, const swb_segment* __restrict__ rows,
               int n_rows, int crop, uint8_t* __restrict__ dst) {
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    const swb_segment s = rows[r];
    const int bh = s.bbox[2] - s.bbox[0], bw = s.bbox[3] - s.bbox[1];
    // floor((crop - dim) / 2) also for negative differences (centre crop)
    const int dy = crop - bh, dxx = crop - bw;
    const int oy = s.bbox[0] - ((dy >= 0) ? dy / 2 : -((-dy + 1) / 2)) + roi_y0;
    const int ox = s.bbox[1] - ((dxx >= 0) ? dxx / 2 : -((-dxx + 1) / 2)) + roi_x0;
    const uint8_t* fr = frames + (long long)s.frame * frame_stride;
    uint8_t* out = dst + (long long)r * crop * crop * channels;
    const int n = crop * crop * channels;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i % channels;
        const int px = (i / channels) % crop;
        const int py = i / (channels * crop);
        const int y = oy + py, x = ox + px;
        uint8_t v = 0;
        if ((unsigned)y < (unsigned)frame_h && (unsigned)x < (unsigned)frame_w)
            v = fr[(long long)y * pitch + (long long)x * channels + c];
        out[i] = v;
    }
}

}  // namespace

template <int BX>
static void launch_local(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b) {
    constexpr int BY = 256 / BX;
    dim3 grid((g.BH + BY - 1) / BY, T);
    k_ccl_local<BX><<<grid, 256, 0, s>>>(fbits, g, b.parent, b.parts, b.pcount, b.cap_parts, b.overflow);
    const int n_boundaries = (g.BH + BY - 1) / BY - 1;
    if (n_boundaries > 0) {
        const int Q = g.wpr4 >> 2;
        dim3 bgrid((Q + 31) / 32, n_boundaries, T);
        k_ccl_boundary<<<bgrid, 32, 0, s>>>(fbits, g, BY, b.parent);
    }
}

cudaError_t launch_ccl(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b,
                       void* labels, int label_elem_size, int* n_launches, cudaEvent_t* ev, int n_ev) {
    int evi = 0;
    auto mark = [&]() {
        if (ev && evi < n_ev) cudaEventRecord(ev[evi++], s);
    };
    // (bx, 256 / bx) threads: bx groups of 4 words along the row, 256 / bx block rows
    const int Q = g.wpr4 >> 2;
    int bx = 1;
    while (bx < Q && bx < 32) bx <<= 1;
    dim3 gblock(bx, 256 / bx);
    dim3 ggrid((Q + bx - 1) / bx, (g.BH + gblock.y - 1) / gblock.y, T);
    const bool tiled = Q <= 32;      // a tile spans the full width: frames up to 4096 pixels wide
    int launches = 0;
    if (tiled) {
        cudaMemsetAsync(b.pcount, 0, sizeof(int), s);
        switch (bx) {
            case 1: launch_local<1>(s, fbits, T, g, b); break;
            case 2: launch_local<2>(s, fbits, T, g, b); break;
            case 4: launch_local<4>(s, fbits, T, g, b); break;
            case 8: launch_local<8>(s, fbits, T, g, b); break;
            case 16: launch_local<16>(s, fbits, T, g, b); break;
            default: launch_local<32>(s, fbits, T, g, b); break;
        }
        launches += 2;
    } else {
        k_ccl_init<<<ggrid, gblock, 0, s>>>(fbits, g, b.parent);
        k_ccl_merge<<<ggrid, gblock, 0, s>>>(fbits, g, b.parent);
        launches += 2;
    }
    mark();
    const long long n_rows_w = (long long)T * g.BH;
    k_ccl_roots<<<(int)((n_rows_w * 32 + 255) / 256), 256, 0, s>>>(fbits, T, g, b.parent, b.rowcount);
    k_ccl_scan<<<(T * 32 + 255) / 256, 256, 0, s>>>(T, g, b.rowcount, b.nseg);
    k_ccl_offsets<<<1, 1024, 0, s>>>(T, b.nseg, b.segoff, b.cap_rows, b.overflow);
    {
        dim3 grid(4, T);
        k_seg_init<<<grid, 256, 0, s>>>(T, b.nseg, b.segoff, b.rows, b.cap_rows);
    }
    launches += 4;
    mark();
    if (tiled) {
        k_ccl_label<false><<<ggrid, gblock, 0, s>>>(fbits, g, b.parent, b.rowcount, b.segoff, b.blocklabel, b.rows,
                                                    b.cap_rows);
        k_props_final<<<296, 256, 0, s>>>(b.parts, b.pcount, b.cap_parts, g, b.parent, b.rowcount, b.segoff, b.rows,
                                          b.cap_rows);
        launches += 2;
    } else {
        k_ccl_label<true><<<ggrid, gblock, 0, s>>>(fbits, g, b.parent, b.rowcount, b.segoff, b.blocklabel, b.rows,
                                                   b.cap_rows);
        launches += 1;
    }
    mark();
    if (labels != nullptr) {
        const int rpt = 8;
        if (label_elem_size == 4) {
            dim3 grid((g.mpitch / 4 + 255) / 256, (g.h + rpt - 1) / rpt, T);
            k_write_labels<int32_t, 4><<<grid, 256, 0, s>>>(fbits, g, b.blocklabel, (int32_t*)labels, rpt);
        } else {
            dim3 grid((g.mpitch / 16 + 255) / 256, (g.h + rpt - 1) / rpt, T);
            k_write_labels<uint8_t, 16><<<grid, 256, 0, s>>>(fbits, g, b.blocklabel, (uint8_t*)labels, rpt);
        }
        launches += 1;
    }
    mark();
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}

cudaError_t launch_pack_bits(cudaStream_t s, const uint8_t* img, int h, int w, uint32_t* fbits, int wpr4) {
    k_pack_bits<<<(h * wpr4 + 255) / 256, 256, 0, s>>>(img, h, w, fbits, wpr4);
    return cudaGetLastError();
}

cudaError_t launch_gather_crops_n(cudaStream_t s, const uint8_t* frames, long long frame_stride, long long pitch,
                                  int channels, int frame_h, int frame_w, int roi_x0, int roi_y0,
                                  const swb_segment* rows, int n_rows, int crop, uint8_t* dst) {
    if (n_rows <= 0) return cudaSuccess;
    k_gather_crops<<<n_rows, 256, 0, s>>>(frames, frame_stride, pitch, channels, frame_h, frame_w, roi_x0, roi_y0,
                                          rows, n_rows, crop, dst);
    return cudaGetLastError();
}

}  // namespace swb
