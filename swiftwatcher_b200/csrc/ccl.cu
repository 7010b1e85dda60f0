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
//   ccl_init    parent[b] = first block of b's horizontal run inside its word
//   ccl_merge   unions across word boundaries and with the block row above
//   ccl_roots   root flags, per-word prefix and per-row counts (warp per row)
//   ccl_scan    per-frame exclusive scan of row counts -> segments per frame
//   ccl_offsets exclusive scan over frames -> row offsets of the segment table
//   seg_init    initialise the table rows (frame, label, empty bbox)
//   ccl_label   per block: root -> label; regionprops atomics into the table
//   write_labels dense int32 / uint8 label image from bits + block labels
#include "swb_internal.cuh"

namespace swb {

namespace {

constexpr uint32_t EVEN = 0x55555555u;

struct RowWords {
    uint32_t A, B;  // rows 2by and 2by+1 of word j
};

__device__ __forceinline__ RowWords load_pair(const uint32_t* fb, const Geom& g, int by, int j) {
    RowWords r;
    const uint32_t* p = fb + (long long)(2 * by) * g.wpr + j;
    r.A = p[0];
    r.B = (2 * by + 1 < g.h) ? p[g.wpr] : 0u;
    return r;
}

__device__ __forceinline__ bool decode(long long idx, const Geom& g, int T, int& f, int& by, int& j) {
    const long long per_frame = (long long)g.BH * g.wpr;
    if (idx >= per_frame * T) return false;
    f = (int)(idx / per_frame);
    int rem = (int)(idx - (long long)f * per_frame);
    by = rem / g.wpr;
    j = rem - by * g.wpr;
    return true;
}

__global__ void __launch_bounds__(256)
k_ccl_init(const uint32_t* __restrict__ fbits, int T, Geom g, int* __restrict__ parent) {
    int f, by, j;
    if (!decode((long long)blockIdx.x * blockDim.x + threadIdx.x, g, T, f, by, j)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr;
    RowWords r = load_pair(fb, g, by, j);
    const uint32_t P = r.A | r.B;
    if (!P) return;
    uint32_t O = (P | (P >> 1)) & EVEN;   // bit 2k: block k occupied
    const uint32_t H = P & (P << 1) & EVEN;  // bit 2k: block k touches block k-1 (same word)
    const int base = by * g.BW + 16 * j;
    int* par = parent + (long long)f * g.BH * g.BW;
    int start = 0;
    while (O) {
        const int b = __ffs(O) - 1;
        O &= O - 1;
        const int k = b >> 1;
        if (!((H >> b) & 1u)) start = k;
        par[base + k] = base + start;
    }
}

__device__ __forceinline__ int find_root(const int* par, int x) {
    int p = par[x];
    while (p != x) {
        x = p;
        p = par[x];
    }
    return x;
}

__device__ __forceinline__ int find_root_volatile(int* par, int x) {
    int p = ((volatile int*)par)[x];
    while (p != x) {
        x = p;
        p = ((volatile int*)par)[x];
    }
    return x;
}

// Lock-free union; larger root is linked under the smaller one.
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

__global__ void __launch_bounds__(256)
k_ccl_merge(const uint32_t* __restrict__ fbits, int T, Geom g, int* parent) {
    int f, by, j;
    if (!decode((long long)blockIdx.x * blockDim.x + threadIdx.x, g, T, f, by, j)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr;
    RowWords r = load_pair(fb, g, by, j);
    const uint32_t P = r.A | r.B;
    if (!P) return;
    int* par = parent + (long long)f * g.BH * g.BW;
    const int base = by * g.BW + 16 * j;

    // horizontal link across the word boundary
    if ((P & 1u) && j > 0) {
        RowWords l = load_pair(fb, g, by, j - 1);
        if ((l.A | l.B) >> 31) unite(par, base, base - 1);
    }
    if (by == 0 || r.A == 0u) return;

    // links with the block row above: only its bottom pixel row matters
    const uint32_t* up = fb + (long long)(2 * by - 1) * g.wpr;
    const uint32_t Bp = up[j];
    const uint32_t BpL = j > 0 ? up[j - 1] : 0u;
    const uint32_t BpR = j + 1 < g.wpr ? up[j + 1] : 0u;
    const uint32_t Bp_l = __funnelshift_l(BpL, Bp, 1);  // bit i = Bp[i-1]
    const uint32_t Bp_r = __funnelshift_r(Bp, BpR, 1);  // bit i = Bp[i+1]
    const uint32_t A = r.A;
    uint32_t UP = (Bp | (Bp >> 1)) & (A | (A >> 1)) & EVEN;   // block k <-> up block k
    uint32_t UL = Bp_l & A & EVEN;                            // pixel (2k) <-> up pixel (2k-1)
    uint32_t UR = ((Bp_r & A) >> 1) & EVEN;                   // pixel (2k+1) <-> up pixel (2k+2)
    // drop links implied by others
    const uint32_t H = P & (P << 1) & EVEN;
    UL &= ~(UP & Bp);                 // up blocks k-1,k already joined through Bp[2k-1],Bp[2k]
    UR &= ~(UP & (Bp >> 1));          // up blocks k,k+1 already joined through Bp[2k+1],Bp[2k+2]
    UP &= ~(H & (UP << 2) & Bp_l & Bp);  // cur k-1~k, up k-1~k and cur k-1 ~ up k-1
    const int upbase = (by - 1) * g.BW + 16 * j;
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

// warp per (frame, block row): root flags + exclusive prefix over the row's words
__global__ void __launch_bounds__(256)
k_ccl_roots(const uint32_t* __restrict__ fbits, int T, Geom g, const int* __restrict__ parent,
            uint32_t* __restrict__ rootbits, uint32_t* __restrict__ wordbase,
            uint32_t* __restrict__ rowcount) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)T * g.BH) return;
    const int f = (int)(wid / g.BH);
    const int by = (int)(wid - (long long)f * g.BH);
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr;
    const int* par = parent + (long long)f * g.BH * g.BW;
    const long long rowoff = ((long long)f * g.BH + by) * g.wpr;
    uint32_t running = 0;
    for (int j0 = 0; j0 < g.wpr; j0 += 32) {
        const int j = j0 + lane;
        uint32_t RB = 0, O = 0;
        if (j < g.wpr) {
            RowWords r = load_pair(fb, g, by, j);
            const uint32_t P = r.A | r.B;
            O = (P | (P >> 1)) & EVEN;
            const int base = by * g.BW + 16 * j;
            uint32_t o = O;
            while (o) {
                const int b = __ffs(o) - 1;
                o &= o - 1;
                if (par[base + (b >> 1)] == base + (b >> 1)) RB |= 1u << b;
            }
        }
        const uint32_t cnt = __popc(RB);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (O) {
            rootbits[rowoff + j] = RB;
            wordbase[rowoff + j] = running + incl - cnt;
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

__global__ void __launch_bounds__(256)
k_ccl_label(const uint32_t* __restrict__ fbits, int T, Geom g, const int* __restrict__ parent,
            const uint32_t* __restrict__ rootbits, const uint32_t* __restrict__ wordbase,
            const uint32_t* __restrict__ rowbase, const int32_t* __restrict__ segoff,
            int* __restrict__ blocklabel, swb_segment* rows, int cap_rows) {
    int f, by, j;
    if (!decode((long long)blockIdx.x * blockDim.x + threadIdx.x, g, T, f, by, j)) return;
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr;
    RowWords r = load_pair(fb, g, by, j);
    const uint32_t P = r.A | r.B;
    if (!P) return;
    const int* par = parent + (long long)f * g.BH * g.BW;
    int* bl = blocklabel + (long long)f * g.BH * g.BW;
    const uint32_t* rb = rootbits + (long long)f * g.BH * g.wpr;
    const uint32_t* wb = wordbase + (long long)f * g.BH * g.wpr;
    const uint32_t* rbase = rowbase + (long long)f * g.BH;
    const long long off = segoff[f];
    const int base = by * g.BW + 16 * j;

    uint32_t O = (P | (P >> 1)) & EVEN;
    Acc acc;
    acc.label = 0;
    int last_parent = -1, last_label = 0;
    while (O) {
        const int b = __ffs(O) - 1;
        O &= O - 1;
        const int k = b >> 1;
        const int p0 = par[base + k];
        int label;
        if (p0 == last_parent) {
            label = last_label;
        } else {
            const int root = find_root(par, p0);
            const int rby = root / g.BW;
            const int rx = root - rby * g.BW;
            const int rj = rx >> 4, rk = rx & 15;
            const uint32_t bitsw = rb[rby * g.wpr + rj];
            label = 1 + (int)(rbase[rby] + wb[rby * g.wpr + rj] + __popc(bitsw & ((1u << (2 * rk)) - 1u)));
            last_parent = p0;
            last_label = label;
        }
        bl[base + k] = label;
        // the block's pixels
        const int a0 = (r.A >> b) & 1, a1 = (r.A >> (b + 1)) & 1;
        const int c0 = (r.B >> b) & 1, c1 = (r.B >> (b + 1)) & 1;
        const int y0 = 2 * by, x0 = 32 * j + b;
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
    flush(acc, rows, off, cap_rows);
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
    const uint32_t* fb = fbits + (long long)f * g.h * g.wpr;
    const int* bl = blocklabel + (long long)f * g.BH * g.BW;
    LT* out = labels + (long long)f * g.h * g.mpitch;
    const int j = x >> 5, sh = x & 31;
    for (int yy = 0; yy < rows_per_thread; ++yy) {
        const int y = yb + yy;
        if (y >= g.h) break;
        const uint32_t bits = (fb[(long long)y * g.wpr + j] >> sh) & ((PX == 32) ? 0xFFFFFFFFu : ((1u << PX) - 1u));
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
k_pack_bits(const uint8_t* __restrict__ img, int h, int w, uint32_t* __restrict__ fbits, int wpr) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= h * wpr) return;
    const int y = idx / wpr, j = idx - y * wpr;
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

cudaError_t launch_ccl(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b,
                       void* labels, int label_elem_size, int* n_launches, cudaEvent_t* ev, int n_ev) {
    const long long n_words = (long long)T * g.BH * g.wpr;
    const int nb_words = (int)((n_words + 255) / 256);
    int evi = 0;
    auto mark = [&]() {
        if (ev && evi < n_ev) cudaEventRecord(ev[evi++], s);
    };
    k_ccl_init<<<nb_words, 256, 0, s>>>(fbits, T, g, b.parent);
    k_ccl_merge<<<nb_words, 256, 0, s>>>(fbits, T, g, b.parent);
    mark();
    const long long n_rows_w = (long long)T * g.BH;
    k_ccl_roots<<<(int)((n_rows_w * 32 + 255) / 256), 256, 0, s>>>(fbits, T, g, b.parent, b.rootbits, b.wordbase,
                                                                  b.rowcount);
    k_ccl_scan<<<(T * 32 + 255) / 256, 256, 0, s>>>(T, g, b.rowcount, b.nseg);
    k_ccl_offsets<<<1, 1024, 0, s>>>(T, b.nseg, b.segoff, b.cap_rows, b.overflow);
    {
        dim3 grid(4, T);
        k_seg_init<<<grid, 256, 0, s>>>(T, b.nseg, b.segoff, b.rows, b.cap_rows);
    }
    mark();
    k_ccl_label<<<nb_words, 256, 0, s>>>(fbits, T, g, b.parent, b.rootbits, b.wordbase, b.rowcount, b.segoff,
                                         b.blocklabel, b.rows, b.cap_rows);
    mark();
    int launches = 7;
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

cudaError_t launch_pack_bits(cudaStream_t s, const uint8_t* img, int h, int w, uint32_t* fbits, int wpr) {
    k_pack_bits<<<(h * wpr + 255) / 256, 256, 0, s>>>(img, h, w, fbits, wpr);
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
