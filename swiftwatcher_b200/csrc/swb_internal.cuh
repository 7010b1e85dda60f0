// Internal declarations shared by the swb200 translation units (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/swb200.h"

namespace swb {

// Geometry of the processed region (the ROI of crop_frame, image_filtering.py:199-203).
//
// Two coordinate systems are used on the device:
//  * "raw" coordinates: the ROI widened to 32-pixel alignment in the source
//    frame, x in [X0a, X0a + wa), X0a = roi_x0 & ~31.  The temporal kernel works
//    here so that every thread reads 48 contiguous, 16-byte aligned BGR bytes.
//  * "roi" coordinates: x in [0, w) exactly as the reference sees the cropped
//    frame.  The morphology kernel realigns raw bits by dx = roi_x0 - X0a, and
//    every later buffer (final bits, mask, labels, 2x2 blocks) is ROI-local.
struct Geom {
    int h, w;        // ROI height / width
    int dx;          // roi_x0 - X0a, 0..31
    int wa;          // raw width in pixels, multiple of 32, >= dx + w
    int wpr_raw;     // wa / 32
    int wpr;         // ceil(w / 32): words per row of the final bit mask
    int wpr4;        // wpr rounded up to 4: row pitch (words) of the final bit mask (zero padded)
    int mpitch;      // wpr * 32: row pitch (elements) of mask / label images
    int BH;          // ceil(h / 2): 2x2-block rows
    int BW;          // 16 * wpr4:   2x2-block columns (padded)
};

// Where the temporal kernel finds frame j (j = 0 is the first output frame).
struct FrameSrc {
    const uint8_t* cur;        // frame 0, ROI row 0, raw column 0 (byte address)
    long long frame_stride;    // bytes between frames
    long long pitch;           // bytes between rows
    int n_inline_halo;         // frames -n_inline_halo..-1 are readable at cur + j*stride
    int hist_valid;            // frames j < 0 come from `hist` instead
    const uint8_t* hist;       // [N-1][h][wa] gray: carried history as compact gray frames; slot s = frame s-(N-1)
    uint8_t* hist_out;         // where to leave the last N-1 frames of this submit (or null)
    int avail_w;               // pixels readable from raw column 0 in a row (guarded path)
};

struct MorphCfg {
    int radius;      // 0 (no morphology), 1 (3x3) or 2 (5x5)
    int n_ops;       // 0, 2 (open) or 4 (open + close) / 2 (close only)
    int is_erode[4]; // op sequence
};

// ---- launchers (each returns the cudaError_t of the launch) ---------------
cudaError_t launch_fg_bits(cudaStream_t s, const FrameSrc& src, int channels, int median_n,
                           int T, const Geom& g, int thresh, uint16_t* raw_bits, bool aligned,
                           int* n_launches, int gpu_share = 1, int forced_ts = 0);
int last_temporal_subchunk();   // frames per temporal sub-chunk of this thread's last launch_fg_bits
cudaError_t launch_morph_mask(cudaStream_t s, const uint32_t* raw_bits, int T, const Geom& g,
                              const MorphCfg& m, uint32_t* fbits, uint8_t* mask, int* n_launches);

// Regionprops partial sums of one tile-local component (ccl.cu, tiled path).
struct Partial {
    int frame, root;               // frame index; global block id of the tile-local root
    int area;
    int minr, minc, maxr, maxc;    // inclusive
    unsigned sr, sc;               // sums of row / column coordinates within the tile
    int flags;                     // 0 = one per tile-local component; 1 / 2 = per-run overflow partial
                                   // (2: the run that starts at the root block)
    int pad[2];
};

struct CclBuffers {
    int* parent;          // [T][BH][BW]
    uint32_t* rowcount;   // [2][T][BH]: roots per block row (exclusive row base after the scan), then rowfill
    int32_t* nseg;        // [T] segments per frame
    int32_t* segoff;      // [T+1] exclusive prefix of nseg
    swb_segment* rows;    // [cap_rows]
    int cap_rows;
    int32_t* overflow;    // device flag: 1 when total rows > cap_rows (or partials > cap_parts)
    Partial* parts;       // [cap_parts] tile-local regionprops partial sums
    int* pcount;          // [2]: number of partials appended; number of listed (dense) tiles
    int* big_tiles;       // [T * tiles per frame] tiles with more runs than the common labelling kernel holds
    int* rootlist;        // [cap_rows] roots grouped by block row (tiled path ranking)
    int cap_parts;
};
// A submit from host memory is cut into sub-batches of frames (each starts as soon as its
// frames have landed on the device); their segment-table offsets are chained: sub-batch k
// starts where k-1 ended (*segoff_base).  Sub-batches run in order on one stream.
struct CclChain {
    int frame_base;                     // first frame of the sub-batch within the submit
    const int32_t* segoff_base;         // device: offset of the sub-batch's first table row (null = 0)
};
cudaError_t launch_ccl(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g,
                       const CclBuffers& b, void* labels, int label_elem_size,
                       int* n_launches, cudaEvent_t* stage_events, int n_stage_events,
                       const CclChain* chain = nullptr, bool prepared = false);
// uint8 compatibility table (SWB_LABELS_U8): the rows of components whose labels agree mod 256
// merged on the device (image_filtering.py:329,335).  The pointers are those of the (sub-)batch.
struct U8Table {
    swb_segment* stage;   // [T][255] per-frame merged rows before compaction
    swb_segment* rows;    // [cap] compacted table (whole submit)
    int32_t* nseg;        // [T]
    int32_t* segoff;      // [T+1]
    int cap;
};
cudaError_t launch_u8_merge(cudaStream_t s, int T, const CclBuffers& b, const U8Table& u, const int32_t* base,
                            int* n_launches);
cudaError_t launch_write_labels(cudaStream_t s, const uint32_t* fbits, int T, const Geom& g, const CclBuffers& b,
                                void* labels, int label_elem_size, int* n_launches);
void ccl_prepare(cudaStream_t s, int T, const Geom& g, const CclBuffers& b, bool chained);

cudaError_t launch_pack_bits(cudaStream_t s, const uint8_t* img, int h, int w, uint32_t* fbits,
                             int wpr4);
cudaError_t launch_gather_crops_n(cudaStream_t s, const uint8_t* frames, long long frame_stride,
                                  long long pitch, int channels, int frame_h, int frame_w,
                                  int roi_x0, int roi_y0, const swb_segment* rows, int n_rows,
                                  int crop, uint8_t* dst, int32_t* rects);

// cudaFuncSetAttribute is per device: remember which devices a kernel's dynamic shared-memory
// limit has been raised on (one process may drive several GPUs, one context each).
struct PerDeviceOnce {
    unsigned long long done = 0;   // bit d: configured on device d (d < 64)
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        if (done & bit) return false;
        done |= bit;
        return true;
    }
};

// ---- RPCA background model (rpca.cu): the reference's own localisation, SURVEY.md §8f #4 ----
// Device-resident state of one IALM run (rpca_run_graph): the iteration loop is a conditional WHILE node of a
// CUDA graph, so everything the host loop kept in local variables lives here and the kernels read it.
struct RpcaState {
    double mu, inv_mu, thr;          // this iteration: mu, 1 / mu, lambda / mu
    double inv_mu_next, thr_next;    // the next iteration's (the fused pass prepares its Gram matrix)
    double dnorm, dual_norm;         // |X|_F;  max(|X|_2-ish, |X|_inf / lambda)  (image_filtering.py:269-275)
    int itr, done, zero, sweeps;     // iterations taken; stop flag; all-black batch; Jacobi sweeps so far
    long long cyc_eigen, cyc_jacobi; // SM cycles spent in k_rpca_eigen21 / in its Jacobi sweeps (all iterations)
    int sweeps_hist[8];              // Jacobi sweeps of the first eight iterations
};

struct RpcaWork {
    double *A0, *A1, *Y;            // [n][P] low-rank iterate (ping-pong) and the Lagrange multiplier
    double *gpart, *G, *W, *zpart;  // per-CTA Gram partials, packed Gram matrix, W = V f(S) V^T, |Z|^2 partials
    unsigned long long* sumsq;      // [2]: sum of squares, maximum of the gray stack
    double* h_buf;                  // pinned host staging (G, W, |Z|^2 partials, norms)
    long long P;
    int nmax, nctas;
    // device-side iteration loop (n = 21): state, eigenbasis of the previous iteration, the graph and what it was built for
    void* state;
    double* Vprev;
    void *graph, *graph_exec;
    const uint8_t* g_X;
    uint8_t* g_out;
    long long g_P;
    int graph_failed, last_mode, host_iters;   // last_mode: 1 = graph loop, 0 = host loop
    int device_loop;                           // option "rpca_device_loop": 1 / 0 force the loop's place, -1 = automatic
};
cudaError_t rpca_alloc(RpcaWork& w, long long P, int nmax);
void rpca_free(RpcaWork& w);
cudaError_t rpca_run(cudaStream_t s, const uint8_t* X, int n, long long P, RpcaWork& w, uint8_t* out, int* iters,
                     int* n_launches);
cudaError_t rpca_read_state(cudaStream_t s, RpcaWork& w, int* iters, int* sweeps, int* mode);
cudaError_t launch_crop_gray(cudaStream_t s, const uint8_t* frames, long long frame_stride, long long pitch,
                             int channels, int x0, int y0, int h, int w, int n, int newest_first, uint8_t* out);
// bilateral_blur (image_filtering.py:304-307 = cv2.bilateralFilter, 8-bit, one channel) for a stack of
// frames; `out` (uint8 images) and / or `bits` (the thresholded result as 1 bit per pixel, rows of
// wpr_bits words: the input of the morphology kernel) may be null.  frame_map: frame f reads image
// (reverse ? n - 1 - f : f) of the stack.  radius: the LUT's radius when the caller knows it (3 selects the
// unrolled kernel), 0 = use the generic kernel.
struct BilateralLut {
    float color[256];
    float space[64];
    int dy[64], dx[64];
    int ntaps, radius;
};
void bilateral_lut(int d, double sigma_color, double sigma_space, BilateralLut& lut);
cudaError_t launch_bilateral(cudaStream_t s, const uint8_t* in, int n, int h, int w, const BilateralLut* d_lut,
                             int reverse, uint8_t* out, int thresh, uint32_t* bits, int wpr_bits, int radius = 0);

// ---- programmatic dependent launch (sm_90+): a kernel launched with launch_dependent() may be
// scheduled while its predecessor in the stream is still draining; it must call
// wait_for_previous_kernel() before it touches anything the predecessor wrote.  The chain of small
// labelling kernels is latency-bound, so hiding the launch gaps is a visible share of their time.
#ifdef __CUDACC__
__device__ __forceinline__ void wait_for_previous_kernel() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void let_next_kernel_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// 4 mask bits -> 4 bytes of 0x00 / 0xFF: put the bits on the byte sign positions
// (disjoint shifted copies, no carries) and let PRMT replicate the sign bits.
// (prmt.b32 directly: __byte_perm masks the replicate bit of the selector away.)
__device__ __forceinline__ uint32_t expand4(uint32_t nibble) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(nibble * 0x10204080u), "r"(0u), "r"(0x0000BA98u));
    return d;
}
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                    Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// single-stage kernels (stages.cu)
cudaError_t launch_stage_gray(cudaStream_t s, const uint8_t* bgr, int h, int w, uint8_t* out);
cudaError_t launch_stage_median(cudaStream_t s, const uint8_t* stack, int n, int h, int w, uint8_t* out);
cudaError_t launch_stage_absdiff(cudaStream_t s, const uint8_t* a, const uint8_t* b, long long n, uint8_t* out);
cudaError_t launch_stage_thresh(cudaStream_t s, const uint8_t* in, long long n, int thresh, uint8_t* out);
cudaError_t launch_stage_minmax(cudaStream_t s, const uint8_t* in, int h, int w, int se_h, int se_w,
                                int is_max, uint8_t* out);
cudaError_t launch_stage_props(cudaStream_t s, const void* labels, int elem_size, int h, int w,
                               swb_segment* acc, int cap);
cudaError_t launch_synth(cudaStream_t s, uint8_t* dst, uint32_t seed, uint32_t video, int t0, int n,
                         int h, int w, int n_birds);

}  // namespace swb
