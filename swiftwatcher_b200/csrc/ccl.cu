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
// roots in raster order — counts and prefix sums, no sort.
//
// Foreground is sparse (birds), so every array indexed by block is allocated
// dense but only touched where the mask is set; the only dense traffic is the
// final label image write.  The unit of work is a RUN: a maximal chain of
// horizontally touching blocks inside one 32-pixel word.  Only run starts take
// part in the union-find; per-run regionprops come from popcounts (area, row
// sums) and bit-plane popcounts (column sums), independent of the run length.
//
// Tiled path (frames up to 4096 pixels wide; one CTA = full-width tile of 8..256
// block rows held in shared memory):
//   ccl_local     the tile's runs are numbered in raster order (block-wide scan) and
//                 listed in shared memory; union-find over run ids, one thread per
//                 run; writes parent[b] = tile-local root for every block and one
//                 regionprops partial sum per tile-local component (smem atomics)
//   ccl_boundary  global unions between the first block row of a tile and the
//                 last block row of the tile above (atomicMin union-find)
// Wide path (wider frames): ccl_init + ccl_merge do everything globally and
//   ccl_label also accumulates the regionprops with global atomics.
// Then, both paths:
// Ranking: tiled path from the compact partial list (root_count / root_place / root_rank:
//   roots per block row, then label = 1 + roots in earlier rows + smaller roots in the
//   same row, parent[root] = -label); wide path by a dense warp-per-row scan (ccl_roots).
//   ccl_scan      per-frame exclusive scan of row counts -> segments per frame
//   ccl_offsets   exclusive scan over frames -> row offsets of the segment table
//   seg_init      initialise the table rows (frame, label, empty bbox)
//   props_final   adds every partial sum to its component's table row
//   write_labels  dense int32 / uint8 label image: bits + parent walk to the tagged root
#include <algorithm>
#include <cstdlib>

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

// ---- runs inside one word ---------------------------------------------------------
// P = A | B (pixel columns occupied in the block row); block k = bits 2k, 2k+1.
__device__ __forceinline__ uint32_t occ_bits(uint32_t P) { return (P | (P >> 1)) & EVEN; }      // bit 2k: block k occupied
__device__ __forceinline__ uint32_t link_bits(uint32_t P) { return P & (P << 1) & EVEN; }      // bit 2k: block k touches k-1
__device__ __forceinline__ uint32_t run_starts(uint32_t P) { return occ_bits(P) & ~link_bits(P); }
// first block of the run that contains (occupied) block k
__device__ __forceinline__ int run_start_of(uint32_t RS, int k) {
    return (31 - __clz((int)(RS & ((2u << (2 * k)) - 1u)))) >> 1;
}
// pixel mask (both pixel columns of every block) of the run starting at block k0
__device__ __forceinline__ uint32_t run_mask(uint32_t H, int k0) {
    const uint32_t hs = (k0 == 15) ? 0u : (H >> (2 * k0 + 2));
    const uint32_t t = ~(hs | 0xAAAAAAAAu);                 // lowest set even bit = first block not linked
    const int len = 1 + ((__ffs((int)t) - 1) >> 1);        // t != 0: zeros are shifted in at the top
    const uint32_t m = (len >= 16) ? 0xFFFFFFFFu : ((1u << (2 * len)) - 1u);
    return m << (2 * k0);
}
// sum of the bit positions of the set bits of w
__device__ __forceinline__ uint32_t pos_sum(uint32_t w) {
    return __popc(w & 0xAAAAAAAAu) + 2 * __popc(w & 0xCCCCCCCCu) + 4 * __popc(w & 0xF0F0F0F0u) +
           8 * __popc(w & 0xFF00FF00u) + 16 * __popc(w & 0xFFFF0000u);
}

struct RunStats {
    uint32_t area, sr, sc;
    int minr, minc, maxr, maxc;   // inclusive
};
// regionprops partial sums of the pixels Am (row y0) and Bm (row y0 + 1), word origin column x0
__device__ __forceinline__ RunStats run_stats(uint32_t Am, uint32_t Bm, int y0, int x0) {
    RunStats s;
    const uint32_t na = __popc(Am), nb = __popc(Bm);
    const uint32_t Pm = Am | Bm;
    s.area = na + nb;
    s.sr = na * (uint32_t)y0 + nb * (uint32_t)(y0 + 1);
    s.sc = pos_sum(Am) + pos_sum(Bm) + s.area * (uint32_t)x0;
    s.minr = Am ? y0 : y0 + 1;
    s.maxr = Bm ? y0 + 1 : y0;
    s.minc = x0 + __ffs((int)Pm) - 1;
    s.maxc = x0 + 31 - __clz((int)Pm);
    return s;
}

// Iterate the runs of a 4-word group in run-major order, so that the n-th runs of all
// lanes of a warp are processed together whatever word they sit in.  RS[] holds the
// remaining run-start flags; selects keep everything in registers.
struct RunIter {
    uint32_t RS[4];
    __device__ __forceinline__ void init(const Group& gr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) RS[i] = run_starts(gr.A[i] | gr.B[i]);
    }
    // next run: word i, first block k0, that word's rows A / B
    __device__ __forceinline__ bool next(const Group& gr, int& i, int& k0, uint32_t& A, uint32_t& B) {
        if ((RS[0] | RS[1] | RS[2] | RS[3]) == 0u) return false;
        i = RS[0] ? 0 : (RS[1] ? 1 : (RS[2] ? 2 : 3));
        const uint32_t rs = i == 0 ? RS[0] : (i == 1 ? RS[1] : (i == 2 ? RS[2] : RS[3]));
        k0 = (__ffs((int)rs) - 1) >> 1;
        const uint32_t rest = rs & (rs - 1);
        RS[0] = i == 0 ? rest : RS[0];
        RS[1] = i == 1 ? rest : RS[1];
        RS[2] = i == 2 ? rest : RS[2];
        RS[3] = i == 3 ? rest : RS[3];
        A = i == 0 ? gr.A[0] : (i == 1 ? gr.A[1] : (i == 2 ? gr.A[2] : gr.A[3]));
        B = i == 0 ? gr.B[0] : (i == 1 ? gr.B[1] : (i == 2 ? gr.B[2] : gr.B[3]));
        return true;
    }
};

// ------------------------------------------------------------------------------------
// Global union-find (wide path and tile boundaries)
// ------------------------------------------------------------------------------------
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
        uint32_t O = occ_bits(P);
        const uint32_t H = link_bits(P);
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

// contact bits between word A (pixel row 2by) and the pixel row above it (Bp, with its
// left / right neighbour words BpL / BpR); links implied by others are dropped
struct Contacts {
    uint32_t UP, UL, UR;   // bit 2k: block k <-> up block k / k-1 / k+1
};
__device__ __forceinline__ Contacts contacts(uint32_t A, uint32_t P, uint32_t BpL, uint32_t Bp, uint32_t BpR) {
    Contacts c;
    const uint32_t Bp_l = __funnelshift_l(BpL, Bp, 1);       // bit x = Bp[x-1]
    const uint32_t Bp_r = __funnelshift_r(Bp, BpR, 1);       // bit x = Bp[x+1]
    c.UP = (Bp | (Bp >> 1)) & (A | (A >> 1)) & EVEN;         // block k      <-> up block k
    c.UL = Bp_l & A & EVEN;                                  // pixel 2k     <-> up pixel 2k-1
    c.UR = ((Bp_r & A) >> 1) & EVEN;                         // pixel 2k+1   <-> up pixel 2k+2
    const uint32_t H = link_bits(P);
    c.UL &= ~(c.UP & Bp);                    // up blocks k-1,k already joined through Bp[2k-1],Bp[2k]
    c.UR &= ~(c.UP & (Bp >> 1));             // up blocks k,k+1 already joined through Bp[2k+1],Bp[2k+2]
    c.UP &= ~(H & (c.UP << 2) & Bp_l & Bp);  // cur k-1~k, up k-1~k and cur k-1 ~ up k-1
    return c;
}

// Unions between the blocks of group (by, q) and the block row above: only the
// bottom pixel row (2by - 1) of that row matters.
// U[1..4] = the pixel row above the group (2by - 1), U[0] / U[5] = its neighbour words
__device__ void vertical_links_with(const Group& gr, const uint32_t (&U)[6], const Geom& g, int by, int q, int* par) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t A = gr.A[i];
        if (!A) continue;
        Contacts c = contacts(A, A | gr.B[i], U[i], U[i + 1], U[i + 2]);
        const int base = by * g.BW + 16 * (4 * q + i);
        const int upbase = base - g.BW;
        while (c.UP) {
            const int k = (__ffs(c.UP) - 1) >> 1;
            c.UP &= c.UP - 1;
            unite(par, base + k, upbase + k);
        }
        while (c.UL) {
            const int k = (__ffs(c.UL) - 1) >> 1;
            c.UL &= c.UL - 1;
            unite(par, base + k, upbase + k - 1);
        }
        while (c.UR) {
            const int k = (__ffs(c.UR) - 1) >> 1;
            c.UR &= c.UR - 1;
            unite(par, base + k, upbase + k + 1);
        }
    }
}

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
    vertical_links_with(gr, U, g, by, q, par);
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

// Global unions between the first block row of every tile (tile_rows apart) and the
// row above it.  grid = (ceil(Q / 32), ceil(n_boundaries / 8), T), block = (32, 8).
__global__ void __launch_bounds__(256)
k_ccl_boundary(const uint32_t* __restrict__ fbits, Geom g, int tile_rows, int* parent) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int q = blockIdx.x * 32 + threadIdx.x;
    const int by = (blockIdx.y * 8 + threadIdx.y + 1) * tile_rows;
    const int f = blockIdx.z;
    if (q >= (g.wpr4 >> 2) || by >= g.BH) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    // every word this thread can need, in one round of independent loads (the kernel is pure latency)
    Group gr;
    load_group(gr, fb, g, by, q);
    const uint32_t* up = fb + (long long)(2 * by - 1) * g.wpr4 + 4 * q;
    const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(up));
    uint32_t U[6];
    U[0] = q > 0 ? __ldg(up - 1) : 0u;
    U[5] = 4 * q + 4 < g.wpr4 ? __ldg(up + 4) : 0u;
    U[1] = u4.x; U[2] = u4.y; U[3] = u4.z; U[4] = u4.w;
    if ((gr.A[0] | gr.A[1] | gr.A[2] | gr.A[3]) == 0u) return;
    if ((U[0] | U[1] | U[2] | U[3] | U[4] | U[5]) == 0u) return;
    int* par = parent + (long long)f * g.BH * g.BW;
    vertical_links_with(gr, U, g, by, q, par);
}

// ------------------------------------------------------------------------------------
// Tiled path: tile-local union-find in shared memory
// ------------------------------------------------------------------------------------
constexpr uint32_t TAG = 0x8000u;  // sp[] entry of a claimed root: TAG | slot
constexpr uint32_t NOSLOT = 0x7FFFu;

__device__ __forceinline__ int find_s(const unsigned short* sp, int x) {
    int p = (int)((const volatile unsigned short*)sp)[x];
    while (p != x) {
        x = p;
        p = (int)((const volatile unsigned short*)sp)[x];
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

__device__ __forceinline__ void emit_partial(Partial* parts, int idx, int f, int root, uint32_t area, int minr,
                                             int minc, int maxr, int maxc, uint32_t sr, uint32_t sc, int flags) {
    Partial p;
    p.frame = f; p.root = root; p.area = (int)area;
    p.minr = minr; p.minc = minc; p.maxr = maxr; p.maxc = maxc;
    p.sr = sr; p.sc = sc; p.flags = flags; p.pad[0] = p.pad[1] = 0;
    parts[idx] = p;
}

// BX = groups (of 4 words) per tile row, a power of two >= Q = wpr4 / 4; BY = 256 / BX block rows;
// the tile always holds 1024 words (32 Ki pixels x 2 rows each).
//
// Sparse-to-dense: the tile's runs are numbered in raster order with a block-wide prefix sum
// (at most 8 runs per word, 8192 per tile) and listed in shared memory; the union-find,
// slot claiming and regionprops phases then run one THREAD PER RUN over that dense list, so
// warps are full instead of having a lane or two busy.  Run ids grow with the block index of
// the run start, so "link the larger id under the smaller" keeps the minimum block as root.
//
// Two sizes of the run list: the common kernel holds up to RUNS_FAST runs (24-28 KB of shared
// memory, eight CTAs per SM -- the tile is latency-bound on a few threads' union chains, so
// resident tiles per SM are what buys throughput); a tile with more runs (dense noise) puts
// itself on a list and is redone by the RUNS_MAX variant (8 runs per word, the worst case).
struct LocalSmem {
    static constexpr int RUNS_FAST = 1536;
    static constexpr int MAXR = 256;            // components with a shared-memory accumulator
    // words of one pixel-row plane of the tile: BY rows of 4 * BX words + a halo word each side
    // (1536 at BX = 1 ... 1056 at BX = 16: 24 KB in all for the common kernel, eight CTAs per SM)
    __host__ __device__ static constexpr int ab_words(int bx, int gpt = 1) { return gpt * (256 / bx) * (4 * bx + 2); }
    // gpt = groups (of four words) per thread: the tile holds 1024 * gpt words.  Frames small enough to fit ONE
    // tile of up to 4096 words (gpt <= 4) are labelled AND ranked by a single CTA (see SINGLE below).
    static constexpr size_t bytes(int maxruns, int bx, int gpt = 1) {
        return (size_t)2 * ab_words(bx, gpt) * 4 + (size_t)1024 * gpt * 2 + (size_t)2 * maxruns * 2 + MAXR * 30 + 64 +
               (size_t)((maxruns + 31) / 32) * 8;
    }
    __host__ __device__ static constexpr int runs_max(int gpt) { return 8192 * gpt; }    // 8 runs per word: the worst case
};

// One tile (frame f, block rows tile * BY ...).  Returns false (block-uniform, nothing written)
// when the tile has more than MAXRUNS runs.
// CAN_EXIT: the caller does nothing after the tile, so the warps that have no run to work on leave after the
// run list is built (the later phases are one thread per run: with the usual few dozen runs per tile seven of
// the eight warps would only walk from barrier to barrier) and the remaining phases synchronise on a named
// barrier sized for the warps that stay.
// SINGLE: the tile is the whole frame (tiles per frame == 1).  Then a tile-local root is a component and its
// rank among the tile's roots — run ids grow with the first block's raster index, i.e. in OpenCV's label order —
// is its label: the kernel tags the roots itself (parent[root] = -label), writes the frame's segment count, and
// the boundary / root_count / scan / root_place / root_rank launches disappear from the chain.
template <int BX, int GPT, int MAXRUNS, bool CAN_EXIT, bool SINGLE>
__device__ __forceinline__ bool ccl_local_tile(const uint32_t* __restrict__ fbits, const Geom& g, int* __restrict__ parent,
                                               Partial* __restrict__ parts, int* __restrict__ pcount, int cap_parts,
                                               int32_t* __restrict__ overflow, int32_t* __restrict__ nseg, int f, int tile) {
    constexpr int BY = GPT * 256 / BX;
    constexpr int WR = BX * 4;               // words per tile row (power of two)
    constexpr int SW = WR + 2;               // + one halo word each side
    constexpr int MAXR = LocalSmem::MAXR;
    constexpr int NW = 1024 * GPT;           // words per tile
    constexpr int RBW = (MAXRUNS + 31) / 32; // words of the root bitmap (SINGLE)
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* sA = reinterpret_cast<uint32_t*>(smem_raw);                 // [BY][SW] pixel row 2by
    uint32_t* sB = sA + LocalSmem::ab_words(BX, GPT);                        // [BY][SW] pixel row 2by + 1
    unsigned short* swpre = reinterpret_cast<unsigned short*>(sB + LocalSmem::ab_words(BX, GPT));   // [NW] runs before word
    unsigned short* srun = swpre + NW;                                    // [MAXRUNS] (word << 4) | first block
    unsigned short* spar = srun + MAXRUNS;                                // [MAXRUNS] parent run id / TAG | slot
    uint32_t* st_area = reinterpret_cast<uint32_t*>(spar + MAXRUNS);
    uint32_t* st_sr = st_area + MAXR;
    uint32_t* st_sc = st_sr + MAXR;
    int* st_minr = reinterpret_cast<int*>(st_sc + MAXR);
    int* st_minc = st_minr + MAXR;
    int* st_maxr = st_minc + MAXR;
    int* st_maxc = st_maxr + MAXR;
    unsigned short* sroot = reinterpret_cast<unsigned short*>(st_maxc + MAXR);
    int* s_misc = reinterpret_cast<int*>(sroot + MAXR);                   // [0..7] warp totals, [8] slots, [9] base
    uint32_t* rootbits = reinterpret_cast<uint32_t*>(s_misc + 16);        // [RBW] roots by run id (SINGLE)
    uint32_t* rootpre = rootbits + RBW;                                   // [RBW] roots in earlier words

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int by0 = tile * BY;
    const int Q = g.wpr4 >> 2;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;

    // ---- phase A: load, publish the rows, number the runs in raster order (a thread's GPT groups are
    // consecutive in tile-raster order, so thread order == raster order)
    uint32_t RS[GPT][4];
    int cnt = 0;
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
        const int idx = tid * GPT + gi;
        const int tx = idx % BX, ty = idx / BX;
        const int by = by0 + ty;
        Group gr;
#pragma unroll
        for (int i = 0; i < 4; ++i) gr.A[i] = gr.B[i] = 0u;
        if (tx < Q && by < g.BH) load_group(gr, fb, g, by, tx);
        const int w0 = ty * SW + 1 + 4 * tx;     // smem index of this group's word 0
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sA[w0 + i] = gr.A[i];
            sB[w0 + i] = gr.B[i];
            RS[gi][i] = run_starts(gr.A[i] | gr.B[i]);
            cnt += __popc(RS[gi][i]);
        }
        if (tx == 0) { sA[ty * SW] = 0u; sB[ty * SW] = 0u; }
        if (tx == BX - 1) { sA[ty * SW + SW - 1] = 0u; sB[ty * SW + SW - 1] = 0u; }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_misc[warp] = incl;
    if (tid == 0) s_misc[8] = 0;
    __syncthreads();
    int base = incl - cnt, nruns = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int t = s_misc[w];
        if (w < warp) base += t;
        nruns += t;
    }
    if (nruns == 0) {                        // empty tile (block-uniform): nothing to write anywhere ...
        if (SINGLE && tid == 0) nseg[f] = 0; // ... but the frame's (zero) segment count
        return true;
    }
    if (nruns > MAXRUNS) return false;       // block-uniform: left to the RUNS_MAX variant
    {
        int r = base;
#pragma unroll
        for (int gi = 0; gi < GPT; ++gi) {
            const int idx = tid * GPT + gi;
            const int wi0 = (idx / BX) * WR + 4 * (idx % BX);    // tile-raster word index
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                swpre[wi0 + i] = (unsigned short)r;
                uint32_t rs = RS[gi][i];
                while (rs) {
                    const int k0 = (__ffs((int)rs) - 1) >> 1;
                    rs &= rs - 1;
                    srun[r] = (unsigned short)(((wi0 + i) << 4) | k0);
                    spar[r] = (unsigned short)r;
                    ++r;
                }
            }
        }
    }
    if (SINGLE)
        for (int w = tid; w < RBW; w += 256) rootbits[w] = 0u;
    __syncthreads();

    const int nact = CAN_EXIT ? min(256, (nruns + 31) & ~31) : 256;   // threads of the per-run phases (block-uniform)
    if (CAN_EXIT && tid >= nact) return true;
    auto phase_barrier = [&]() {
        if (CAN_EXIT) asm volatile("bar.sync 1, %0;" ::"r"(nact) : "memory");
        else __syncthreads();
    };

    // id of the run that contains (occupied) block k of the word at smem position sw / raster index wi
    auto run_id = [&](int sw, int wi, int k) -> int {
        const uint32_t rsu = run_starts(sA[sw] | sB[sw]);
        return (int)swpre[wi] + __popc(rsu & ((2u << (2 * k)) - 1u)) - 1;
    };

    // ---- phase B: one thread per run: unions with the word to the left and the block row above
    for (int r = tid; r < nruns; r += 256) {
        const int d = (int)srun[r];
        const int wi = d >> 4, k0 = d & 15;
        const int rty = wi / WR, wx = wi % WR;
        const int sw = rty * SW + 1 + wx;
        const uint32_t A = sA[sw], B = sB[sw];
        const uint32_t P = A | B;
        if (k0 == 0 && (P & 1u)) {
            const uint32_t Pl = sA[sw - 1] | sB[sw - 1];
            if (Pl >> 31) unite_s(spar, r, run_id(sw - 1, wi - 1, 15));
        }
        if (rty == 0) continue;
        const uint32_t m = run_mask(link_bits(P), k0);
        const uint32_t Am = A & m;
        if (!Am) continue;
        const int su = sw - SW, wu = wi - WR;            // the word above
        Contacts c = contacts(Am, P & m, sB[su - 1], sB[su], sB[su + 1]);
        int last_b = -1;
        while (c.UP) {
            const int k = (__ffs((int)c.UP) - 1) >> 1;
            c.UP &= c.UP - 1;
            const int b = run_id(su, wu, k);
            if (b != last_b) unite_s(spar, r, b);
            last_b = b;
        }
        while (c.UL) {
            const int k = (__ffs((int)c.UL) - 1) >> 1;
            c.UL &= c.UL - 1;
            const int b = (k > 0) ? run_id(su, wu, k - 1) : run_id(su - 1, wu - 1, 15);
            if (b != last_b) unite_s(spar, r, b);
            last_b = b;
        }
        while (c.UR) {
            const int k = (__ffs((int)c.UR) - 1) >> 1;
            c.UR &= c.UR - 1;
            const int b = (k < 15) ? run_id(su, wu, k + 1) : run_id(su + 1, wu + 1, 0);
            if (b != last_b) unite_s(spar, r, b);
            last_b = b;
        }
    }
    phase_barrier();

    // ---- phase C: roots claim an accumulator slot (entry becomes TAG | slot)
    for (int r = tid; r < nruns; r += 256) {
        if ((int)spar[r] != r) continue;
        if (SINGLE) atomicOr(&rootbits[r >> 5], 1u << (r & 31));
        const int slot = atomicAdd(&s_misc[8], 1);
        if (slot < MAXR) {
            sroot[slot] = (unsigned short)r;
            st_area[slot] = 0u; st_sr[slot] = 0u; st_sc[slot] = 0u;
            st_minr[slot] = 0x7FFFFFFF; st_minc[slot] = 0x7FFFFFFF;
            st_maxr[slot] = -1; st_maxc[slot] = -1;
            spar[r] = (unsigned short)(TAG | (uint32_t)slot);
        } else {
            spar[r] = (unsigned short)(TAG | NOSLOT);
        }
    }
    phase_barrier();
    if (SINGLE) {
        // roots in earlier bitmap words (one warp, fixed order); label of root r = 1 + roots with a smaller run id
        if (tid < 32) {
            uint32_t running = 0;
            const int nwords = (nruns + 31) >> 5;
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const int w = w0 + lane;
                const uint32_t c = w < nwords ? __popc(rootbits[w]) : 0u;
                uint32_t inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= d) inc += v;
                }
                if (w < nwords) rootpre[w] = running + inc - c;
                running += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
            if (lane == 0) nseg[f] = (int32_t)running;
        }
        phase_barrier();
    }
    auto root_label = [&](int r) -> int {
        return 1 + (int)rootpre[r >> 5] + __popc(rootbits[r >> 5] & ((1u << (r & 31)) - 1u));
    };

    // tile-local run id -> global block id of its first block
    auto run_gid = [&](int r) -> int {
        const int d = (int)srun[r];
        const int wi = d >> 4;
        return (by0 + wi / WR) * g.BW + (wi % WR) * 16 + (d & 15);
    };

    // ---- phase D: one thread per run: root, global parent of every block, regionprops partial sums
    int* par_f = parent + (long long)f * g.BH * g.BW;
    for (int r = tid; r < nruns; r += 256) {
        int x = r;
        uint32_t p = spar[x];
        while (!(p & TAG)) {
            x = (int)p;
            p = spar[x];
        }
        const uint32_t slot = p & NOSLOT;
        const int rgid = run_gid(x);
        const int d = (int)srun[r];
        const int wi = d >> 4, k0 = d & 15;
        const int rty = wi / WR, wx = wi % WR;
        const int sw = rty * SW + 1 + wx;
        const uint32_t A = sA[sw], B = sB[sw];
        const uint32_t P = A | B;
        const uint32_t m = run_mask(link_bits(P), k0);
        int* par = par_f + (long long)(by0 + rty) * g.BW + wx * 16;
        uint32_t O = occ_bits(P & m);
        while (O) {
            par[(__ffs((int)O) - 1) >> 1] = rgid;
            O &= O - 1;
        }
        if (SINGLE && x == r) par[k0] = -root_label(r);          // the root's own block carries the tag: -label
        const RunStats s = run_stats(A & m, B & m, 2 * (by0 + rty), 32 * wx);
        if (slot < (uint32_t)MAXR) {
            atomicAdd(&st_area[slot], s.area);
            atomicAdd(&st_sr[slot], s.sr);
            atomicAdd(&st_sc[slot], s.sc);
            atomicMin(&st_minr[slot], s.minr);
            atomicMin(&st_minc[slot], s.minc);
            atomicMax(&st_maxr[slot], s.maxr);
            atomicMax(&st_maxc[slot], s.maxc);
        } else {   // more than MAXR components in this tile: one partial per run, straight to global
            const int idx = atomicAdd(pcount, 1);
            if (idx < cap_parts) emit_partial(parts, idx, f, rgid, s.area, s.minr, s.minc, s.maxr, s.maxc, s.sr,
                                              s.sc, x == r ? 2 : 1);
            else *overflow = 1;
        }
    }
    phase_barrier();

    // ---- phase E: one partial per tile-local component -> global list
    const int n = min(s_misc[8], MAXR);
    if (tid == 0) s_misc[9] = atomicAdd(pcount, n);
    phase_barrier();
    for (int s = tid; s < n; s += 256) {
        const int idx = s_misc[9] + s;
        if (idx < cap_parts) {
            emit_partial(parts, idx, f, run_gid((int)sroot[s]), st_area[s], st_minr[s], st_minc[s], st_maxr[s],
                         st_maxc[s], st_sr[s], st_sc[s], 0);
        } else {
            *overflow = 1;
        }
    }
    return true;
}

// grid = (tiles per frame, T): every tile; tiles with too many runs go on the list
template <int BX, int GPT, bool SINGLE>
__global__ void __launch_bounds__(256)
k_ccl_local(const uint32_t* __restrict__ fbits, Geom g, int* __restrict__ parent, Partial* __restrict__ parts,
            int* __restrict__ pcount, int cap_parts, int32_t* __restrict__ overflow, int* __restrict__ big_tiles,
            int* __restrict__ big_count, int32_t* __restrict__ nseg) {
    wait_for_previous_kernel();
    if (!ccl_local_tile<BX, GPT, LocalSmem::RUNS_FAST, true, SINGLE>(fbits, g, parent, parts, pcount, cap_parts, overflow,
                                                                      nseg, (int)blockIdx.y, (int)blockIdx.x)) {
        if (threadIdx.x == 0) big_tiles[atomicAdd(big_count, 1)] = (int)(blockIdx.y * gridDim.x + blockIdx.x);
    }
}

// the listed tiles, with room for the worst case (usually none: the CTAs find an empty list and leave)
template <int BX, int GPT, bool SINGLE>
__global__ void __launch_bounds__(256)
k_ccl_local_big(const uint32_t* __restrict__ fbits, Geom g, int* __restrict__ parent, Partial* __restrict__ parts,
                int* __restrict__ pcount, int cap_parts, int32_t* __restrict__ overflow,
                const int* __restrict__ big_tiles, const int* __restrict__ big_count, int tiles_per_frame,
                int32_t* __restrict__ nseg) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int n = *big_count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int t = big_tiles[i];
        ccl_local_tile<BX, GPT, LocalSmem::runs_max(GPT), false, SINGLE>(fbits, g, parent, parts, pcount, cap_parts, overflow,
                                                                         nseg, t / tiles_per_frame, t % tiles_per_frame);
        __syncthreads();                     // shared memory is reused by the next tile
    }
}

// ------------------------------------------------------------------------------------
// Ranking, labels, table
// ------------------------------------------------------------------------------------
// warp per (frame, block row): every root (a run start with parent[b] == b) gets
// parent[b] = -(1 + its rank among the roots of this block row); rowcount = roots in the row.
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
                    uint32_t rs = run_starts(gr.A[i] | gr.B[i]);
                    const int base = by * g.BW + 16 * (4 * q + i);
                    while (rs) {
                        const int b = __ffs((int)rs) - 1;
                        rs &= rs - 1;
                        if (par[base + (b >> 1)] == base + (b >> 1)) RB[i] |= 1u << b;
                    }
                }
            }
        }
        const uint32_t cnt = __popc(RB[0]) + __popc(RB[1]) + __popc(RB[2]) + __popc(RB[3]);
        const uint32_t anyroot = __ballot_sync(0xFFFFFFFFu, cnt != 0u);
        if (anyroot == 0u) continue;                      // warp-uniform
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
                    const int b = __ffs((int)r) - 1;
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
    wait_for_previous_kernel();
    let_next_kernel_launch();
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
k_ccl_offsets(int T, const int32_t* __restrict__ nseg, int32_t* segoff, const int32_t* base, int cap_rows,
              int32_t* __restrict__ overflow) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = base ? *base : 0;   // sub-batches of one submit chain their offsets (base == segoff of the previous one + its T)
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
k_seg_init(int T, int frame_base, const int32_t* __restrict__ nseg, const int32_t* __restrict__ segoff,
           swb_segment* __restrict__ rows, int cap_rows) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    // one warp per frame (a frame has tens to a few hundred rows): T / 8 CTAs instead of 4 T mostly idle ones
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= T) return;
    const int n = nseg[f];
    for (int i = threadIdx.x & 31; i < n; i += 32) {
        const long long r = (long long)segoff[f] + i;
        if (r >= cap_rows) return;
        swb_segment s;
        s.frame = frame_base + f;
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

__device__ __forceinline__ void add_to_row(swb_segment* s, uint32_t area, int minr, int minc, int maxr, int maxc,
                                           unsigned long long sr, unsigned long long sc) {
    atomicAdd(&s->area, (int)area);
    atomicMin(&s->bbox[0], minr);
    atomicMin(&s->bbox[1], minc);
    atomicMax(&s->bbox[2], maxr + 1);
    atomicMax(&s->bbox[3], maxc + 1);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_row), sr);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum_col), sc);
}

// label of an occupied block: walk parents to the tagged root.  Tiled path: the tag is
// -label.  Wide path: the tag is -(1 + rank in the block row) and rowbase is added.
__device__ __forceinline__ int label_of(const int* par, int x, const uint32_t* rbase, int BW) {
    int v = par[x];
    while (v >= 0) {
        x = v;
        v = par[x];
    }
    return rbase ? (int)rbase[x / BW] - v : -v;
}

// Wide path only: per run, add the run's regionprops to its component's table row
// with global atomics (roots are tagged -(1 + rank in their block row)).
__global__ void __launch_bounds__(256)
k_ccl_props_wide(const uint32_t* __restrict__ fbits, Geom g, const int* __restrict__ parent,
                 const uint32_t* __restrict__ rowbase, const int32_t* __restrict__ segoff, swb_segment* rows,
                 int cap_rows) {
    int f, by, q;
    if (!my_group(g, f, by, q)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    Group gr;
    load_group(gr, fb, g, by, q);
    if (!gr.any()) return;
    const int* par = parent + (long long)f * g.BH * g.BW;
    const uint32_t* rbase = rowbase + (long long)f * g.BH;
    const long long off = segoff[f];
    RunIter it;
    it.init(gr);
    int i, k0;
    uint32_t A, B;
    while (it.next(gr, i, k0, A, B)) {
        const int label = label_of(par, by * g.BW + 16 * (4 * q + i) + k0, rbase, g.BW);
        const uint32_t m = run_mask(link_bits(A | B), k0);
        const RunStats s = run_stats(A & m, B & m, 2 * by, 32 * (4 * q + i));
        const long long r = off + label - 1;
        if (r < cap_rows) add_to_row(rows + r, s.area, s.minr, s.minc, s.maxr, s.maxc, s.sr, s.sc);
    }
}

// Ranking from the partial list (tiled path).  Every component has exactly one
// "representative" partial: the slot partial of the tile that holds its root, or the
// overflow partial of the run that starts at the root (flags 0 / 2).  A representative
// whose root survived the boundary merges (parent[root] == root) is a component.
__device__ __forceinline__ bool is_root_partial(const Partial& p, const int* par) {
    return p.flags != 1 && par[p.root] == p.root;
}

// pass 1: roots per block row
__global__ void __launch_bounds__(256)
k_root_count(const Partial* __restrict__ parts, const int* __restrict__ pcount, int cap_parts, Geom g,
             const int* __restrict__ parent, uint32_t* __restrict__ rowcount) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int n = min(*pcount, cap_parts);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Partial p = parts[i];
        if (is_root_partial(p, parent + (long long)p.frame * g.BH * g.BW))
            atomicAdd(&rowcount[(long long)p.frame * g.BH + p.root / g.BW], 1u);
    }
}

// pass 2: roots of one block row side by side (arbitrary order) in rootlist
__global__ void __launch_bounds__(256)
k_root_place(const Partial* __restrict__ parts, const int* __restrict__ pcount, int cap_parts, Geom g,
             const int* __restrict__ parent, const uint32_t* __restrict__ rowbase,
             uint32_t* __restrict__ rowfill, const int32_t* __restrict__ segoff, int* __restrict__ rootlist,
             int cap_rows) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int n = min(*pcount, cap_parts);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Partial p = parts[i];
        if (!is_root_partial(p, parent + (long long)p.frame * g.BH * g.BW)) continue;
        const long long row = (long long)p.frame * g.BH + p.root / g.BW;
        const long long pos = (long long)segoff[p.frame] + rowbase[row] + atomicAdd(&rowfill[row], 1u);
        if (pos < cap_rows) rootlist[pos] = p.root;
    }
}

// pass 3: label = 1 + roots in earlier rows + smaller roots in the same row; parent[root] = -label
// ... and the root's thread initialises its row of the segment table (frame, label, empty sums)
__global__ void __launch_bounds__(256)
k_root_rank(const Partial* __restrict__ parts, const int* __restrict__ pcount, int cap_parts, Geom g,
            int* __restrict__ parent, const uint32_t* __restrict__ rowbase, const uint32_t* __restrict__ rowfill,
            const int32_t* __restrict__ segoff, const int* __restrict__ rootlist, int cap_rows, int frame_base,
            swb_segment* __restrict__ rows) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int n = min(*pcount, cap_parts);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Partial p = parts[i];
        int* par = parent + (long long)p.frame * g.BH * g.BW;
        if (p.flags == 1) continue;
        const int self = par[p.root];
        if (self != p.root) continue;           // merged away (self >= 0) -- tags are written below, only by us
        const long long row = (long long)p.frame * g.BH + p.root / g.BW;
        const long long lo = (long long)segoff[p.frame] + rowbase[row];
        const int cnt = (int)rowfill[row];
        int rank = 0;
        for (int j = 0; j < cnt; ++j)
            if (lo + j < cap_rows && rootlist[lo + j] < p.root) ++rank;
        const int label = (int)(rowbase[row] + rank + 1);
        par[p.root] = -label;
        const long long r = (long long)segoff[p.frame] + label - 1;
        if (r < cap_rows) {
            swb_segment s;
            s.frame = frame_base + p.frame;
            s.label = label;
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
}

// every partial sum -> its component's row of the segment table
__global__ void __launch_bounds__(256)
k_props_final(const Partial* __restrict__ parts, const int* __restrict__ pcount, int cap_parts, Geom g,
              const int* __restrict__ parent, const int32_t* __restrict__ segoff, swb_segment* rows,
              int cap_rows) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int n = min(*pcount, cap_parts);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Partial p = parts[i];
        const int label = label_of(parent + (long long)p.frame * g.BH * g.BW, p.root, nullptr, g.BW);
        const long long r = (long long)segoff[p.frame] + label - 1;
        if (r >= cap_rows) continue;
        add_to_row(rows + r, (uint32_t)p.area, p.minr, p.minc, p.maxr, p.maxc, p.sr, p.sc);
    }
}

// Dense int32 label image.  A warp owns a span of 128 pixels (4 labels per lane = one 16-byte
// store) over 8 rows.  The bit words of the span are loaded once, coalesced (one word per lane
// per round), and handed to the lanes that need them with shuffles, so the row loop has no
// dependent global load on the (dominant) background path; rows are taken in pairs because the
// two pixel rows of a block row share their 2x2 blocks (one walk to the tagged root for both).
__global__ void __launch_bounds__(256)
k_write_labels_i32(const uint32_t* __restrict__ fbits, Geom g, const int* __restrict__ parent,
                   const uint32_t* __restrict__ rowbase, int32_t* __restrict__ labels) {
    wait_for_previous_kernel();
    constexpr int PX = 4;                   // labels per lane and row = one 16-byte store
    constexpr int WPS = PX;                 // words per span row (32 * PX pixels / 32)
    constexpr int ROWS = 8;                 // = 32 / WPS: one load round covers the span
    constexpr int BR = ROWS / 2;            // block rows
    const int lane = threadIdx.x & 31;
    const int span = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int f = blockIdx.z;
    const int yb = blockIdx.y * ROWS;
    const int x = span * (32 * PX) + lane * PX;          // first pixel of this lane
    if (span * (32 * PX) >= g.mpitch) return;            // warp-uniform
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr4;
    uint32_t wreg;
    {
        const int y = yb + lane / WPS;
        const int j = span * WPS + lane % WPS;
        wreg = (y < g.h && j < g.wpr4) ? __ldg(fb + (long long)y * g.wpr4 + j) : 0u;
    }
    const int* par = parent + (long long)f * g.BH * g.BW;
    const int wsel = (lane * PX) >> 5, sh = (lane * PX) & 31;
    uint32_t bitsA[BR], bitsB[BR];
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) {
        const uint32_t wordA = __shfl_sync(0xFFFFFFFFu, wreg, (2 * rr) * WPS + wsel);
        const uint32_t wordB = __shfl_sync(0xFFFFFFFFu, wreg, (2 * rr + 1) * WPS + wsel);
        bitsA[rr] = (wordA >> sh) & ((1u << PX) - 1u);
        bitsB[rr] = (wordB >> sh) & ((1u << PX) - 1u);   // rows past the image were loaded as 0
    }
    if (x >= g.mpitch) return;
    // The lane's 2 blocks in each of the 4 block rows: all first-level parent loads are issued together
    // and the (two or three deep) chains to the tagged roots advance in lock-step, so a warp that
    // sees birds pays one chain of dependent loads, not one per block.
    int v[BR][2], xr[BR][2];
    const int b00 = (yb >> 1) * g.BW + (x >> 1);
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) {
        const uint32_t P = bitsA[rr] | bitsB[rr];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            xr[rr][i] = b00 + rr * g.BW + i;
            v[rr][i] = ((P >> (2 * i)) & 3u) ? __ldg(par + xr[rr][i]) : -1;
        }
    }
    bool more;
    do {
        more = false;
#pragma unroll
        for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (v[rr][i] >= 0) {
                    xr[rr][i] = v[rr][i];
                    v[rr][i] = par[xr[rr][i]];
                    more = true;
                }
    } while (more);
    const uint32_t* rbase = rowbase ? rowbase + (long long)f * g.BH : nullptr;
    int32_t* out = labels + (long long)f * g.h * g.mpitch + x;
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) {
        const int y = yb + 2 * rr;
        if (y >= g.h) break;
        int lab[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) lab[i] = rbase ? (int)rbase[xr[rr][i] / g.BW] - v[rr][i] : -v[rr][i];
        int4 a, b;
        a.x = (bitsA[rr] & 1u) ? lab[0] : 0; a.y = (bitsA[rr] & 2u) ? lab[0] : 0;
        a.z = (bitsA[rr] & 4u) ? lab[1] : 0; a.w = (bitsA[rr] & 8u) ? lab[1] : 0;
        b.x = (bitsB[rr] & 1u) ? lab[0] : 0; b.y = (bitsB[rr] & 2u) ? lab[0] : 0;
        b.z = (bitsB[rr] & 4u) ? lab[1] : 0; b.w = (bitsB[rr] & 8u) ? lab[1] : 0;
        __stcs(reinterpret_cast<int4*>(out + (long long)y * g.mpitch), a);
        if (y + 1 < g.h) __stcs(reinterpret_cast<int4*>(out + (long long)(y + 1) * g.mpitch), b);
    }
}

// Dense uint8 label image (SWB_LABELS_U8: labels.astype(np.uint8), image_filtering.py:329).  One byte
// per pixel is a quarter of the int32 traffic, so the span-per-warp layout above would be bound by its
// instruction stream (one lane with a bird in its 16 pixels stalls the warp in a long unrolled path).
// Here a lane owns one 32-pixel bit word = 32 output bytes: background words are zero stores, and a
// word with foreground runs one compact loop over its RUNS (the walk to the tagged root from the run's
// first block, then the run's mask bits expanded to bytes and ANDed with the replicated label).
__global__ void __launch_bounds__(256, 8)
k_write_labels_u8(const uint32_t* __restrict__ fbits, Geom g, int T, const int* __restrict__ parent,
                  const uint32_t* __restrict__ rowbase, uint8_t* __restrict__ labels) {
    wait_for_previous_kernel();
    // block = (32 words, 8 block rows); grid = (ceil(words per row / 32), ceil(BH / 8), T): no index divisions.
    // A lane owns one word of BOTH pixel rows of a block row: one label lookup per occupied 2x2 block.
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int by = blockIdx.y * 8 + threadIdx.y;
    const int f = blockIdx.z;
    if (j >= (g.mpitch >> 5) || by >= g.BH) return;
    const int y = 2 * by;
    const bool two = y + 1 < g.h;
    const long long row = (long long)f * g.h + y;
    const uint32_t wa = __ldg(fbits + row * g.wpr4 + j);
    const uint32_t wb = two ? __ldg(fbits + (row + 1) * g.wpr4 + j) : 0u;
    uint32_t oa[8], ob[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) oa[q] = ob[q] = 0u;
    if (wa | wb) {
        const int* par = parent + (long long)f * g.BH * g.BW;
        const uint32_t* rbase = rowbase ? rowbase + (long long)f * g.BH : nullptr;
        const int b0 = by * g.BW + 16 * j;
        // one lookup per RUN (a maximal chain of touching 2x2 blocks inside the word: all of it is one
        // component): the walk starts at the run's first block, whose parent chain ends at the tagged
        // root (the tag is -label).  A word usually holds one run, so the lanes of a warp that see a
        // bird do their two or three dependent loads side by side instead of once per block.
        const uint32_t P = wa | wb;
        const uint32_t H = link_bits(P);
        uint32_t rs = run_starts(P);
        while (rs) {
            const int k0 = (__ffs((int)rs) - 1) >> 1;
            rs &= rs - 1;
            const uint32_t m = run_mask(H, k0);
            int x = b0 + k0, v = par[x];
            while (v >= 0) {
                x = v;
                v = par[x];
            }
            const uint32_t lab = (uint32_t)(rbase ? (int)rbase[x / g.BW] - v : -v) & 0xFFu;
            const uint32_t rep = lab * 0x01010101u;
            const uint32_t ma = wa & m, mb = wb & m;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                oa[q] |= expand4((ma >> (4 * q)) & 0xFu) & rep;
                ob[q] |= expand4((mb >> (4 * q)) & 0xFu) & rep;
            }
        }
    }
    // one 32-byte store per lane and row (STG.256): whole sectors, a warp writes 1 KB of a label row
    uint8_t* o = labels + row * g.mpitch + 32 * j;
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(oa[0]), "r"(oa[1]), "r"(oa[2]),
                 "r"(oa[3]), "r"(oa[4]), "r"(oa[5]), "r"(oa[6]), "r"(oa[7])
                 : "memory");
    if (two)
        asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + g.mpitch), "r"(ob[0]), "r"(ob[1]),
                     "r"(ob[2]), "r"(ob[3]), "r"(ob[4]), "r"(ob[5]), "r"(ob[6]), "r"(ob[7])
                     : "memory");
}

// ------------------------------------------------------------------------------------
// uint8 compatibility (SWB_LABELS_U8): regionprops of labels.astype(np.uint8)
// (image_filtering.py:329,335).  Components whose int32 labels agree mod 256 are ONE region of
// the truncated image (area / sums add, bbox is the union) and labels that are 0 mod 256
// vanish into the background.  One CTA per frame folds the frame's rows into 255 shared-memory
// accumulators and leaves them compacted (ascending label) in the frame's slice of `stage`;
// k_ccl_offsets then scans the per-frame counts and k_u8_compact packs the table.
__global__ void __launch_bounds__(256)
k_u8_merge(const swb_segment* __restrict__ rows, const int32_t* __restrict__ segoff, int cap_rows,
           swb_segment* __restrict__ stage, int32_t* __restrict__ nseg_u8) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    __shared__ int s_area[256], s_b0[256], s_b1[256], s_b2[256], s_b3[256], s_frame;
    __shared__ unsigned long long s_sr[256], s_sc[256];
    __shared__ int s_wtot[8];
    const int f = blockIdx.x, v = threadIdx.x;
    s_area[v] = 0; s_b0[v] = 0x7FFFFFFF; s_b1[v] = 0x7FFFFFFF; s_b2[v] = 0; s_b3[v] = 0;
    s_sr[v] = 0ull; s_sc[v] = 0ull;
    __syncthreads();
    const int lo = min(segoff[f], cap_rows), hi = min(segoff[f + 1], cap_rows);
    for (int i = lo + v; i < hi; i += 256) {
        const swb_segment r = rows[i];
        if (i == lo) s_frame = r.frame;
        const int k = r.label & 0xFF;
        if (k == 0) continue;
        atomicAdd(&s_area[k], r.area);
        atomicMin(&s_b0[k], r.bbox[0]);
        atomicMin(&s_b1[k], r.bbox[1]);
        atomicMax(&s_b2[k], r.bbox[2]);
        atomicMax(&s_b3[k], r.bbox[3]);
        atomicAdd(&s_sr[k], (unsigned long long)r.sum_row);
        atomicAdd(&s_sc[k], (unsigned long long)r.sum_col);
    }
    __syncthreads();
    const bool used = s_area[v] > 0;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, used);
    const int lane = v & 31, warp = v >> 5;
    if (lane == 0) s_wtot[warp] = __popc(bal);
    __syncthreads();
    int pos = __popc(bal & ((1u << lane) - 1u)), total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) pos += s_wtot[w];
        total += s_wtot[w];
    }
    if (used) {
        swb_segment o;
        o.frame = s_frame; o.label = v; o.area = s_area[v];
        o.bbox[0] = s_b0[v]; o.bbox[1] = s_b1[v]; o.bbox[2] = s_b2[v]; o.bbox[3] = s_b3[v];
        o.reserved = 0;
        o.sum_row = (long long)s_sr[v]; o.sum_col = (long long)s_sc[v];
        stage[(long long)f * 255 + pos] = o;
    }
    if (v == 0) nseg_u8[f] = total;
}

__global__ void __launch_bounds__(256)
k_u8_compact(const swb_segment* __restrict__ stage, const int32_t* __restrict__ segoff_u8,
             swb_segment* __restrict__ out, int cap) {
    wait_for_previous_kernel();
    let_next_kernel_launch();
    const int f = blockIdx.x, i = threadIdx.x;
    const int lo = segoff_u8[f], n = segoff_u8[f + 1] - lo;
    if (i < n && lo + i < cap) out[lo + i] = stage[(long long)f * 255 + i];
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

// extract_segment_images (image_filtering.py:338-369) + the classifier's Resize((24, 24))
// (segment_classification.py:20) for every row of the segment table, one CTA per row.
//
// The rectangle is the reference's: bbox grown symmetrically to crop x crop where it is smaller
// (floor / ceil split), shifted by the ROI origin and cut from the FULL frame with numpy's slice
// semantics — a negative start wraps around (in practice: an empty image), an end past the frame
// is truncated, a bbox larger than the crop is kept whole.  A crop x crop rectangle is copied.
// Anything else goes through what transforms.Resize does to the PIL image: Pillow's two-pass
// BILINEAR resampling (src/libImaging/Resample.c: precompute_coeffs in double, coefficients in
// 22-bit fixed point, the first pass rounded to uint8, horizontal first — vertical first when the
// image is more than 100 times taller than wide and shrinks vertically, Image.resize).  An empty
// rectangle gives a zero tile; rects[] tells the caller which case a row was.
struct ResampleAxis {           // one output axis of `crop` samples over n_in input samples
    double scale, support, ss;
    int n_in;
};
__device__ __forceinline__ ResampleAxis make_axis(int n_in, int n_out) {
    ResampleAxis a;
    a.n_in = n_in;
    a.scale = __ddiv_rn((double)(float)n_in, (double)n_out);
    const double fs = a.scale < 1.0 ? 1.0 : a.scale;
    a.support = fs;               // bilinear: support 1.0 * filterscale
    a.ss = __ddiv_rn(1.0, fs);
    return a;
}
struct Taps { int lo, n; double center, ww; };
__device__ __forceinline__ double tap_weight(const ResampleAxis& a, const Taps& t, int j) {
    double x = __dmul_rn(__dadd_rn(__dsub_rn((double)(j + t.lo), t.center), 0.5), a.ss);
    if (x < 0.0) x = -x;
    return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}
__device__ __forceinline__ Taps make_taps(const ResampleAxis& a, int out_index) {
    Taps t;
    t.center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)out_index, 0.5), a.scale));
    int lo = (int)__dadd_rn(__dsub_rn(t.center, a.support), 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)__dadd_rn(__dadd_rn(t.center, a.support), 0.5);
    if (hi > a.n_in) hi = a.n_in;
    t.lo = lo;
    t.n = hi - lo;
    t.ww = 0.0;
    for (int j = 0; j < t.n; ++j) t.ww = __dadd_rn(t.ww, tap_weight(a, t, j));
    return t;
}
__device__ __forceinline__ int tap_fixed(const ResampleAxis& a, const Taps& t, int j) {
    double k = tap_weight(a, t, j);
    if (t.ww != 0.0) k = __ddiv_rn(k, t.ww);
    return (int)__dadd_rn(0.5, __dmul_rn(k, 4194304.0));        // 1 << PRECISION_BITS (22)
}
__device__ __forceinline__ int clip8(int acc) {
    const int v = acc >> 22;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

constexpr int MAX_CROP = 64;

__global__ void __launch_bounds__(256)
k_gather_crops(const uint8_t* __restrict__ frames, long long frame_stride, long long pitch, int channels,
               int frame_h, int frame_w, int roi_x0, int roi_y0, const swb_segment* __restrict__ rows,
               int n_rows, int crop, uint8_t* __restrict__ dst, int32_t* __restrict__ rects) {
    __shared__ Taps s_taps[2][MAX_CROP];          // [0]: along x (horizontal pass), [1]: along y
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    const swb_segment s = rows[r];
    int lo[2], hi[2];                              // [0] = rows (y), [1] = columns (x)
    const int lim[2] = {frame_h, frame_w};
    const int org[2] = {roi_y0, roi_x0};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        int b0 = s.bbox[a], b1 = s.bbox[a + 2];
        const int d = crop - (b1 - b0);
        if (d > 0) { b0 -= d / 2; b1 += (d + 1) / 2; }
        b0 += org[a];
        b1 += org[a];
        // frame[b0:b1]: numpy slice semantics
        lo[a] = b0 < 0 ? max(b0 + lim[a], 0) : min(b0, lim[a]);
        hi[a] = b1 < 0 ? max(b1 + lim[a], 0) : min(b1, lim[a]);
        if (hi[a] < lo[a]) hi[a] = lo[a];
    }
    const int ny = hi[0] - lo[0], nx = hi[1] - lo[1];
    if (rects && threadIdx.x < 4)
        rects[4 * r + threadIdx.x] = threadIdx.x == 0 ? lo[0] : (threadIdx.x == 1 ? lo[1] : (threadIdx.x == 2 ? hi[0] : hi[1]));
    const uint8_t* src = frames + (long long)s.frame * frame_stride + (long long)lo[0] * pitch + (long long)lo[1] * channels;
    uint8_t* out = dst + (long long)r * crop * crop * channels;
    const int n = crop * crop * channels;
    if (ny == 0 || nx == 0) {                      // block-uniform
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = 0;
        return;
    }
    const bool need_h = nx != crop, need_v = ny != crop;
    if (!need_h && !need_v) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int c = i % channels, px = (i / channels) % crop, py = i / (channels * crop);
            out[i] = src[(long long)py * pitch + px * channels + c];
        }
        return;
    }
    const ResampleAxis ax = make_axis(nx, crop), ay = make_axis(ny, crop);
    if (threadIdx.x < crop) {
        if (need_h) s_taps[0][threadIdx.x] = make_taps(ax, threadIdx.x);
    } else if (threadIdx.x >= 64 && threadIdx.x < 64 + crop) {
        if (need_v) s_taps[1][threadIdx.x - 64] = make_taps(ay, threadIdx.x - 64);
    }
    __syncthreads();
    const bool v_first = need_h && need_v && (long long)ny > 100ll * nx;   // ny > crop is implied (nx >= 1, crop <= 64)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i % channels, px = (i / channels) % crop, py = i / (channels * crop);
        int val;
        if (need_h && need_v) {
            // outer pass over the second axis, inner (first) pass evaluated on the fly and rounded to uint8
            const int oa = v_first ? 0 : 1;                      // outer axis: 0 = x taps, 1 = y taps
            const ResampleAxis& A_out = oa == 0 ? ax : ay;
            const ResampleAxis& A_in = oa == 0 ? ay : ax;
            const Taps& T_out = s_taps[oa][oa == 0 ? px : py];
            const Taps& T_in = s_taps[oa ^ 1][oa == 0 ? py : px];
            int acc = 1 << 21;
            for (int jo = 0; jo < T_out.n; ++jo) {
                int inner = 1 << 21;
                for (int ji = 0; ji < T_in.n; ++ji) {
                    const int y = oa == 0 ? T_in.lo + ji : T_out.lo + jo;
                    const int x = oa == 0 ? T_out.lo + jo : T_in.lo + ji;
                    inner += (int)src[(long long)y * pitch + x * channels + c] * tap_fixed(A_in, T_in, ji);
                }
                acc += clip8(inner) * tap_fixed(A_out, T_out, jo);
            }
            val = clip8(acc);
        } else if (need_h) {
            const Taps& T = s_taps[0][px];
            int acc = 1 << 21;
            for (int j = 0; j < T.n; ++j) acc += (int)src[(long long)py * pitch + (T.lo + j) * channels + c] * tap_fixed(ax, T, j);
            val = clip8(acc);
        } else {
            const Taps& T = s_taps[1][py];
            int acc = 1 << 21;
            for (int j = 0; j < T.n; ++j) acc += (int)src[(long long)(T.lo + j) * pitch + px * channels + c] * tap_fixed(ay, T, j);
            val = clip8(acc);
        }
        out[i] = (uint8_t)val;
    }
}

}  // namespace

template <int BX, int GPT, bool SINGLE>
static void launch_local(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b) {
    constexpr int BY = GPT * 256 / BX;
    const int tiles = (g.BH + BY - 1) / BY;
    dim3 grid(tiles, T);
    static PerDeviceOnce once;
    if (once.need()) {
        cudaFuncSetAttribute(k_ccl_local<BX, GPT, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)LocalSmem::bytes(LocalSmem::RUNS_FAST, BX, GPT));
        cudaFuncSetAttribute(k_ccl_local_big<BX, GPT, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)LocalSmem::bytes(LocalSmem::runs_max(GPT), BX, GPT));
    }
    int* big_count = b.pcount + 1;
    launch_dependent(k_ccl_local<BX, GPT, SINGLE>, grid, dim3(256), LocalSmem::bytes(LocalSmem::RUNS_FAST, BX, GPT), s, fbits, g,
                     b.parent, b.parts, b.pcount, b.cap_parts, b.overflow, b.big_tiles, big_count, b.nseg);
    const int big_grid = (int)std::min<long long>((long long)tiles * T, 148 * 4);
    launch_dependent(k_ccl_local_big<BX, GPT, SINGLE>, dim3(big_grid), dim3(256),
                     LocalSmem::bytes(LocalSmem::runs_max(GPT), BX, GPT), s, fbits, g, b.parent, b.parts, b.pcount,
                     b.cap_parts, b.overflow, b.big_tiles, big_count, tiles, b.nseg);
    const int n_boundaries = tiles - 1;
    if (n_boundaries > 0) {
        const int Q = g.wpr4 >> 2;
        dim3 bgrid((Q + 31) / 32, (n_boundaries + 7) / 8, T);
        launch_dependent(k_ccl_boundary, bgrid, dim3(32, 8), 0, s, fbits, g, BY, b.parent);
    }
}

// frames that fit one tile of up to 4096 words: groups per thread (1, 2 or 4), or 0 when the frame needs several tiles
static int single_tile_gpt(const Geom& g, int bx) {
    if (bx > 8) return 0;
    for (int gpt = 1; gpt <= 4; gpt *= 2)
        if (gpt * 256 / bx >= g.BH) return gpt;
    return 0;
}

template <int BX>
static void launch_local_single(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b, int gpt) {
    if constexpr (BX <= 8) {
        if (gpt == 1) launch_local<BX, 1, true>(s, fbits, T, g, b);
        else if (gpt == 2) launch_local<BX, 2, true>(s, fbits, T, g, b);
        else launch_local<BX, 4, true>(s, fbits, T, g, b);
    }
}

// grid of the kernels that walk the partial list (its length is only known on the device): enough CTAs
// that a thread sees one or two partials — they are chains of dependent loads, parallelism is what hides them
constexpr int PART_GRID = 148 * 16;

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SWB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

// The counters the labelling kernels start from.  Done before the filtering kernels of a submit so
// that nothing but kernels sits between K2 and the labelling chain (programmatic dependent launch).
void ccl_prepare(cudaStream_t s, int T, const Geom& g, const CclBuffers& b, bool chained) {
    cudaMemsetAsync(b.pcount, 0, 2 * sizeof(int), s);                  // partial counter, listed-tile counter
    if (!chained) cudaMemsetAsync(b.overflow, 0, sizeof(int32_t), s);  // a chained submit clears it once, before its first sub-batch
    if ((g.wpr4 >> 2) <= 32)
        cudaMemsetAsync(b.rowcount, 0, (size_t)2 * T * g.BH * sizeof(uint32_t), s);   // rowcount, rowfill (tiled path)
}

// The dense label image of T frames (after the ranking kernels have tagged the roots).
cudaError_t launch_write_labels(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b,
                                void* labels, int label_elem_size, int* n_launches) {
    const bool tiled = (g.wpr4 >> 2) <= 32;
    const uint32_t* rbase_for_labels = tiled ? nullptr : b.rowcount;
    // int32: one warp per span of 128 pixels x 8 rows, up to 8 warps per CTA (uint8: see k_write_labels_u8)
    const int px_per_span = 128;
    const int nspans = (g.mpitch + px_per_span - 1) / px_per_span;
    const int wpb = nspans < 8 ? nspans : 8;
    dim3 grid((nspans + wpb - 1) / wpb, (g.h + 7) / 8, T);
    if (label_elem_size == 4) {
        launch_dependent(k_write_labels_i32, grid, dim3(32 * wpb), 0, s, fbits, g, b.parent, rbase_for_labels,
                         (int32_t*)labels);
    } else {
        const dim3 ugrid(((g.mpitch >> 5) + 31) / 32, (g.BH + 7) / 8, T);
        launch_dependent(k_write_labels_u8, ugrid, dim3(32, 8), 0, s, fbits, g, T, b.parent, rbase_for_labels,
                         (uint8_t*)labels);
    }
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ccl(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b,
                       void* labels, int label_elem_size, int* n_launches, cudaEvent_t* ev, int n_ev,
                       const CclChain* chain, bool prepared) {
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
    if (!prepared) ccl_prepare(s, T, g, b, chain != nullptr);
    const int single_gpt = tiled ? single_tile_gpt(g, bx) : 0;    // > 0: one CTA labels and ranks a whole frame
    if (single_gpt > 0) {
        switch (bx) {
            case 1: launch_local_single<1>(s, fbits, T, g, b, single_gpt); break;
            case 2: launch_local_single<2>(s, fbits, T, g, b, single_gpt); break;
            case 4: launch_local_single<4>(s, fbits, T, g, b, single_gpt); break;
            default: launch_local_single<8>(s, fbits, T, g, b, single_gpt); break;
        }
        launches += 2;
        mark();
        launch_dependent(k_ccl_offsets, dim3(1), dim3(1024), 0, s, T, b.nseg, b.segoff,
                         chain ? chain->segoff_base : (const int32_t*)nullptr, b.cap_rows, b.overflow);
        launch_dependent(k_seg_init, dim3((T + 7) / 8), dim3(256), 0, s, T, chain ? chain->frame_base : 0, b.nseg, b.segoff, b.rows,
                         b.cap_rows);
        mark();
        launch_dependent(k_props_final, dim3(PART_GRID), dim3(256), 0, s, b.parts, b.pcount, b.cap_parts, g, b.parent,
                         b.segoff, b.rows, b.cap_rows);
        launches += 3;
        mark();
        if (labels != nullptr) {
            launch_write_labels(s, fbits, T, g, b, labels, label_elem_size, nullptr);
            launches += 1;
        }
        mark();
        if (n_launches) *n_launches += launches;
        return cudaGetLastError();
    }
    if (tiled) {
        switch (bx) {
            case 1: launch_local<1, 1, false>(s, fbits, T, g, b); break;
            case 2: launch_local<2, 1, false>(s, fbits, T, g, b); break;
            case 4: launch_local<4, 1, false>(s, fbits, T, g, b); break;
            case 8: launch_local<8, 1, false>(s, fbits, T, g, b); break;
            case 16: launch_local<16, 1, false>(s, fbits, T, g, b); break;
            default: launch_local<32, 1, false>(s, fbits, T, g, b); break;
        }
        launches += 3;
        mark();
        launch_dependent(k_root_count, dim3(PART_GRID), dim3(256), 0, s, b.parts, b.pcount, b.cap_parts, g, b.parent,
                         b.rowcount);
    } else {
        k_ccl_init<<<ggrid, gblock, 0, s>>>(fbits, g, b.parent);
        k_ccl_merge<<<ggrid, gblock, 0, s>>>(fbits, g, b.parent);
        launches += 2;
        mark();
        const long long n_rows_w = (long long)T * g.BH;
        k_ccl_roots<<<(int)((n_rows_w * 32 + 255) / 256), 256, 0, s>>>(fbits, T, g, b.parent, b.rowcount);
    }
    launch_dependent(k_ccl_scan, dim3((T * 32 + 255) / 256), dim3(256), 0, s, T, g, b.rowcount, b.nseg);
    launch_dependent(k_ccl_offsets, dim3(1), dim3(1024), 0, s, T, b.nseg, b.segoff,
                     chain ? chain->segoff_base : (const int32_t*)nullptr, b.cap_rows, b.overflow);
    if (!tiled) {
        dim3 grid((T + 7) / 8);
        k_seg_init<<<grid, 256, 0, s>>>(T, chain ? chain->frame_base : 0, b.nseg, b.segoff, b.rows, b.cap_rows);
        launches += 1;
    }
    launches += 3;
    if (tiled) {
        uint32_t* rowfill = b.rowcount + (size_t)T * g.BH;
        launch_dependent(k_root_place, dim3(PART_GRID), dim3(256), 0, s, b.parts, b.pcount, b.cap_parts, g, b.parent,
                         b.rowcount, rowfill, b.segoff, b.rootlist, b.cap_rows);
        launch_dependent(k_root_rank, dim3(PART_GRID), dim3(256), 0, s, b.parts, b.pcount, b.cap_parts, g, b.parent,
                         b.rowcount, rowfill, b.segoff, b.rootlist, b.cap_rows, chain ? chain->frame_base : 0, b.rows);
        mark();
        launch_dependent(k_props_final, dim3(PART_GRID), dim3(256), 0, s, b.parts, b.pcount, b.cap_parts, g, b.parent,
                         b.segoff, b.rows, b.cap_rows);
        launches += 3;
    } else {
        mark();
        k_ccl_props_wide<<<ggrid, gblock, 0, s>>>(fbits, g, b.parent, b.rowcount, b.segoff, b.rows, b.cap_rows);
        launches += 1;
    }
    mark();
    if (labels != nullptr) {
        launch_write_labels(s, fbits, T, g, b, labels, label_elem_size, nullptr);
        launches += 1;
    }
    mark();
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}

cudaError_t launch_u8_merge(cudaStream_t s, int T, const CclBuffers& b, const U8Table& u, const int32_t* base,
                            int* n_launches) {
    launch_dependent(k_u8_merge, dim3(T), dim3(256), 0, s, (const swb_segment*)b.rows, (const int32_t*)b.segoff,
                     b.cap_rows, u.stage, u.nseg);
    launch_dependent(k_ccl_offsets, dim3(1), dim3(1024), 0, s, T, (const int32_t*)u.nseg, u.segoff, base, u.cap,
                     b.overflow);
    launch_dependent(k_u8_compact, dim3(T), dim3(256), 0, s, (const swb_segment*)u.stage, (const int32_t*)u.segoff,
                     u.rows, u.cap);
    if (n_launches) *n_launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_pack_bits(cudaStream_t s, const uint8_t* img, int h, int w, uint32_t* fbits, int wpr4) {
    k_pack_bits<<<(h * wpr4 + 255) / 256, 256, 0, s>>>(img, h, w, fbits, wpr4);
    return cudaGetLastError();
}

cudaError_t launch_gather_crops_n(cudaStream_t s, const uint8_t* frames, long long frame_stride, long long pitch,
                                  int channels, int frame_h, int frame_w, int roi_x0, int roi_y0,
                                  const swb_segment* rows, int n_rows, int crop, uint8_t* dst, int32_t* rects) {
    if (n_rows <= 0) return cudaSuccess;
    if (crop > MAX_CROP) return cudaErrorInvalidValue;
    k_gather_crops<<<n_rows, 256, 0, s>>>(frames, frame_stride, pitch, channels, frame_h, frame_w, roi_x0, roi_y0,
                                          rows, n_rows, crop, dst, rects);
    return cudaGetLastError();
}

}  // namespace swb
