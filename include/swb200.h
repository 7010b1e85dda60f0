/*
 * swb200.h — C ABI of the B200-native swiftwatcher filtering/segmentation path.
 *
 * The reference (joshuacwnewton/swiftwatcher) is pure Python and has no FFI on
 * this path; its seam is FrameQueue.preprocess_queue / FrameQueue.segment_queue
 * (swiftwatcher/data_structures.py:171-217), which call the free functions of
 * swiftwatcher/image_filtering.py once per 21-frame batch
 * (swiftwatcher/__main__.py:77-78).  Every entry point below names the
 * reference function(s) it replaces.  Plain pointers and sizes only; no
 * torch / C++ types cross this boundary.  All functions return SWB_OK (0) or a
 * negative error code; the message is available through swb_last_error().
 * Nothing here ever falls back to the CPU: without a CUDA device every compute
 * entry point fails with SWB_ERR_CUDA.
 */
#ifndef SWB200_H
#define SWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB_OK            0
#define SWB_ERR_INVALID  -1   /* bad argument / configuration            */
#define SWB_ERR_CUDA     -2   /* CUDA runtime error, or no device        */
#define SWB_ERR_CAPACITY -3   /* more frames / segments than configured  */
#define SWB_ERR_STATE    -4   /* call order (collect before submit, ...) */

#define SWB_MEM_HOST   0      /* pointer is host memory (pinned or pageable) */
#define SWB_MEM_DEVICE 1      /* pointer is device memory on cfg.device       */

#define SWB_LABELS_I32 0      /* int32 labels 1..n, OpenCV numbering                         */
#define SWB_LABELS_U8  1      /* reference behaviour: labels.astype(uint8)                   */
                              /* (image_filtering.py:329) and regionprops of THAT image      */

#define SWB_HALO_CARRY (-1)   /* n_halo value: use the history carried from the last submit  */

#define SWB_OUT_MASK    1     /* flags: materialise the uint8 {0,255} foreground mask  */
#define SWB_OUT_LABELS  2     /* flags: materialise the dense label image              */

/* Pipeline configuration: the literals the reference hard-codes at its call
 * sites (data_structures.py:194-206, __main__.py:78) made explicit. */
typedef struct swb_config {
    int32_t device;          /* CUDA device ordinal                                          */
    int32_t frame_h;         /* full frame height (rows)                                     */
    int32_t frame_w;         /* full frame width (pixels)                                    */
    int32_t channels;        /* 3 = BGR interleaved uint8, 1 = already grayscale             */
    int64_t frame_pitch;     /* bytes per frame row; 0 = frame_w * channels                  */
    int64_t frame_stride;    /* bytes per frame; 0 = frame_h * frame_pitch                   */
    int32_t roi_x0, roi_y0;  /* crop_region[0] = (x0, y0)   (image_filtering.py:199-203)     */
    int32_t roi_x1, roi_y1;  /* crop_region[1] = (x1, y1), exclusive                         */
    int32_t median_n;        /* temporal median window, odd, 1..9 (BASELINE: 5, 9)           */
    int32_t threshold;       /* thresh_to_zero threshold, 15 at data_structures.py:198       */
    int32_t morph_size;      /* square structuring element: 0 (none), 3 or 5                 */
    int32_t do_open;         /* grey_opening  (data_structures.py:202)                       */
    int32_t do_close;        /* grey_closing after the opening (BASELINE "open/close")       */
    int32_t label_mode;      /* SWB_LABELS_I32 | SWB_LABELS_U8                               */
    int32_t out_flags;       /* SWB_OUT_MASK | SWB_OUT_LABELS                                */
    int32_t max_frames;      /* max output frames per submit                                 */
    int32_t max_segments;    /* max segment rows per submit (0 = 1024 per frame).  Exceeding it is
                              * reported by swb_collect (SWB_ERR_CAPACITY), never silently.  A large
                              * submit from host memory is filtered in up to four sub-batches of
                              * frames, each with an equal share of the labelling scratch, so a
                              * submit whose segments all sit in a few frames may need a larger value.
                              * SWB_LABELS_U8: the limit applies to the components before they are
                              * merged mod 256 */
    int32_t bg_model;        /* SWB_BG_MEDIAN (rolling median, BASELINE.json) or SWB_BG_RPCA */
    int32_t gpu_share;       /* contexts expected to work side by side on this GPU (e.g. one per video,
                              * each on its own stream); 0 or 1 = the context has the GPU to itself.
                              * Only sizes grids (longer temporal sub-chunks, fewer CTAs per launch):
                              * results do not depend on it                                         */
    int32_t reserved;
} swb_config;

/* Background models.  SWB_BG_RPCA is the reference's own localisation (rpca + bilateral_blur,
 * image_filtering.py:220-307, data_structures.py:191-196): every submit is one batch of
 * n_frames <= 32 frames (the reference's queue holds 21) that is decomposed on its own — no
 * temporal history, n_halo is ignored, median_n is unused — then bilateralFilter(7, 15, 1),
 * threshold, opening and labelling as usual.  For the reference's batch of 21 frames the whole iteration
 * loop can run on the device (a CUDA-graph WHILE node: eigenproblem, stopping test and all; swb_submit is
 * then asynchronous as in the median mode): the default for that batch size, see the option
 * "rpca_device_loop".  Otherwise (other batch sizes, or the option set to 0) the stopping test and the n x n
 * eigenproblem run on the host and swb_submit BLOCKS until the decomposition has converged (it cannot be
 * used on a stream that is being captured). */
#define SWB_BG_MEDIAN 0
#define SWB_BG_RPCA   1

/* One row of the per-frame segment table: what skimage.measure.regionprops
 * (image_filtering.py:332-335) exposes and swiftwatcher consumes.
 * centroid = (sum_row / area, sum_col / area) in float64 reproduces
 * numpy's coords.mean(axis=0) bit-exactly (integer sums, one rounding). */
typedef struct swb_segment {
    int32_t frame;           /* output frame index within the submit (0-based)   */
    int32_t label;           /* label value in the label image                   */
    int32_t area;            /* pixel count                                      */
    int32_t bbox[4];         /* (min_row, min_col, max_row, max_col), half-open  */
    int32_t reserved;
    int64_t sum_row;         /* sum of row coordinates over the segment's pixels */
    int64_t sum_col;         /* sum of column coordinates                        */
} swb_segment;

typedef struct swb_ctx swb_ctx;

/* Library / error plumbing ------------------------------------------------- */
const char* swb_version(void);
/* Message of the last error on `ctx` (or, with ctx == NULL, of the last failed
 * call that had no context).  The reference has no error convention on this
 * path (cv2.error / IndexError); the Python wrapper raises RuntimeError. */
const char* swb_last_error(const swb_ctx* ctx);
int swb_device_count(int32_t* count);

/* Context ------------------------------------------------------------------ */
/* One context per (GPU, video); owns all device memory and a stream.  Not
 * thread-safe.  Replaces the FrameQueue batch state (data_structures.py:116-124). */
int swb_create(const swb_config* cfg, swb_ctx** out);
int swb_destroy(swb_ctx* ctx);
/* Forget the carried temporal history (start of a new video). */
int swb_reset(swb_ctx* ctx);
/* Launch on an external CUDA stream (cudaStream_t passed as void*) instead of
 * the context's own; NULL restores the own stream. */
int swb_set_stream(swb_ctx* ctx, void* cuda_stream);

/* The hot path ------------------------------------------------------------- */
/* Replaces FrameQueue.preprocess_queue + segment_queue
 * (data_structures.py:171-217): crop_frame, convert_grayscale, [rolling
 * median + absdiff per BASELINE.json in place of rpca + bilateral_blur],
 * thresh_to_zero, grayscale_opening[/closing], cc_labeling,
 * get_segment_properties for `n_frames` consecutive frames.
 *
 * `frames` points at n_halo + n_frames full frames (oldest first) laid out as
 * cfg.frame_stride / frame_pitch say.  The first n_halo frames are temporal
 * history only (0 <= n_halo <= median_n - 1); history that is not supplied is
 * the earliest supplied frame replicated.  n_halo == SWB_HALO_CARRY uses the
 * history the context carried over from the previous submit instead.
 * Asynchronous on the context's stream; the caller keeps `frames` alive until
 * swb_collect / swb_sync returns. */
int swb_submit(swb_ctx* ctx, const uint8_t* frames, int32_t n_frames,
               int32_t n_halo, int32_t mem_kind);

/* Waits for the last submit and copies its segment table to host memory.
 * rows: capacity `cap` rows, ordered by (frame, label); per_frame_counts:
 * n_frames entries (may be NULL).  Replaces get_segment_properties over the
 * queue (data_structures.py:210). */
int swb_collect(swb_ctx* ctx, swb_segment* rows, int64_t cap, int64_t* n_rows,
                int32_t* per_frame_counts);
int swb_sync(swb_ctx* ctx);
/* swb_collect plus the dense outputs of the whole submit in one call and (normally) one stream
 * synchronisation: masks [n_frames][roi_h][roi_w] uint8 and labels [n_frames][roi_h][roi_w] (int32 or
 * uint8 as configured), tightly packed, either may be NULL.  With page-locked destinations
 * (swb_host_alloc) all copies are DMA transfers queued behind the kernels.  This is what one
 * FrameQueue.segment_queue batch needs back (data_structures.py:202-217). */
int swb_collect_all(swb_ctx* ctx, swb_segment* rows, int64_t cap, int64_t* n_rows,
                    int32_t* per_frame_counts, uint8_t* masks, void* labels);
/* The same in two halves: swb_collect_begin queues every device -> host copy behind the kernels of the
 * last submit and returns at once; swb_collect_end waits for them and reports the counts.  Between the two
 * the caller is free to work (e.g. the tracker on the previous batch); the destination buffers must stay
 * valid and page-locked destinations are what makes the copies asynchronous.  No swb_submit in between. */
int swb_collect_begin(swb_ctx* ctx, swb_segment* rows, int64_t cap, uint8_t* masks, void* labels);
int swb_collect_end(swb_ctx* ctx, int64_t* n_rows, int32_t* per_frame_counts);
/* Tuning knobs that never change a result: "host_pipeline" (0/1: cut large host submits into
 * sub-batches that are filtered while later frames are still being copied; default 1),
 * "sub_batch_min_px" (least work per sub-batch in pixels; default 64 Mi), "rpca_device_loop" (SWB_BG_RPCA, 21-frame batches: 1 = the
 * iteration loop as a CUDA-graph WHILE node on the device, swb_submit asynchronous; 0 = the host loop, swb_submit
 * blocks; -1 = automatic, the default: the device loop), "temporal_subchunk" (frames per
 * temporal sub-chunk of the filtering kernel, rounded up to a multiple of 6; 0 = chosen from the grid size). */
int swb_set_option(swb_ctx* ctx, const char* name, int64_t value);
/* Frames per temporal sub-chunk the filtering kernel used for the last submit (each sub-chunk
 * re-reads its median_n - 1 predecessors); lets tests aim at the sub-chunk boundaries. */
int swb_last_subchunk(swb_ctx* ctx, int32_t* frames);

/* Dense per-frame outputs of the last submit, frames [t0, t0 + n):
 * mask: uint8 {0,255}, == (opened > 0) of data_structures.py:202-204;
 * labels: int32 (SWB_LABELS_I32) or uint8 (SWB_LABELS_U8), data_structures.py:206.
 * dst is tightly packed [n][roi_h][roi_w]. */
int swb_get_masks(swb_ctx* ctx, int32_t t0, int32_t n, uint8_t* dst, int32_t mem_kind);
int swb_get_labels(swb_ctx* ctx, int32_t t0, int32_t n, void* dst, int32_t mem_kind);
/* Bit-packed final mask (1 bit/pixel, little-endian bit order inside uint32
 * words, words_per_row = ceil(roi_w / 32)), tightly packed. */
int swb_get_mask_bits(swb_ctx* ctx, int32_t t0, int32_t n, uint32_t* dst, int32_t mem_kind);
/* Zero-copy views of the context-owned output buffers (device pointers). */
int swb_device_views(swb_ctx* ctx, uint8_t** mask, int64_t* mask_pitch,
                     void** labels, int64_t* labels_pitch_elems,
                     swb_segment** rows, int32_t** per_frame_counts);

/* Per-kernel device timing of the last submit (CUDA events on the launch
 * stream).  names: up to `cap` static strings; ms: milliseconds. */
int swb_enable_timing(swb_ctx* ctx, int32_t on);
int swb_get_timing(swb_ctx* ctx, const char** names, float* ms, int32_t cap, int32_t* n);
/* Number of kernels launched by this context since creation. */
int64_t swb_launch_count(const swb_ctx* ctx);

/* Batched segment crops for the classifier, on the device: extract_segment_images
 * (image_filtering.py:338-369) for every row of the last submit's table, followed by what the
 * classifier's transforms.Resize((24, 24)) (segment_classification.py:20) does to it.
 * rects [n_rows][4] (may be NULL) receives the rectangle (y0, x0, y1, x1) the reference slices from
 * the FULL frame: the bbox grown symmetrically to crop x crop where it is smaller, shifted by the ROI
 * origin, numpy slice semantics — a start left / above the frame wraps around (an empty image unless
 * the frame is tiny), an end past the frame is truncated, a bbox larger than crop is kept whole.
 * dst [n_rows][crop][crop][channels]: the rectangle itself when it is exactly crop x crop; otherwise
 * the rectangle resampled to crop x crop exactly as Pillow's Image.resize(BILINEAR) does it (two
 * passes, 22-bit fixed-point coefficients); all zeros when the rectangle is empty (the reference's
 * ToPILImage raises there — the caller sees it in rects).  crop <= 64.  SWB_LABELS_U8: the rows are
 * those of the merged table.  Needs the full frames of the last submit on the device. */
int swb_gather_crops(swb_ctx* ctx, int32_t crop, uint8_t* dst, int32_t* rects, int32_t mem_kind);

/* Single-stage entry points: one reference function each, host buffers in/out,
 * tightly packed.  They exist so that each reference function has a drop-in
 * with the same meaning; the fused swb_submit path is the fast one. -------- */
/* convert_grayscale, image_filtering.py:188-196 (cv2 BGR2GRAY, 15-bit fixed point). */
int swb_stage_gray(int32_t device, const uint8_t* bgr, int32_t h, int32_t w, uint8_t* out);
/* rolling temporal median of n (odd, <= 9) gray frames [n][h][w] (no reference fn). */
int swb_stage_median(int32_t device, const uint8_t* stack, int32_t n, int32_t h, int32_t w, uint8_t* out);
/* cv2.absdiff (no reference fn). */
int swb_stage_absdiff(int32_t device, const uint8_t* a, const uint8_t* b, int32_t h, int32_t w, uint8_t* out);
/* thresh_to_zero, image_filtering.py:310-316. */
int swb_stage_thresh_to_zero(int32_t device, const uint8_t* in, int32_t h, int32_t w, int32_t thresh, uint8_t* out);
/* grayscale_opening (closing=0), image_filtering.py:319-322, or its dual (closing=1);
 * flat se_h x se_w structuring element, scipy 'reflect' border. */
int swb_stage_grey_morph(int32_t device, const uint8_t* in, int32_t h, int32_t w,
                         int32_t se_h, int32_t se_w, int32_t closing, uint8_t* out);
/* cc_labeling, image_filtering.py:325-329: 8-connected, OpenCV numbering.
 * out_i32 and/or out_u8 may be NULL. */
int swb_stage_cc_label(int32_t device, const uint8_t* in, int32_t h, int32_t w,
                       int32_t* out_i32, uint8_t* out_u8, int32_t* n_labels);
/* get_segment_properties, image_filtering.py:332-335, on a uint8 or int32
 * label image (elem_size 1 or 4).  Rows ordered by label. */
int swb_stage_regionprops(int32_t device, const void* labels, int32_t elem_size,
                          int32_t h, int32_t w, swb_segment* rows, int32_t cap, int32_t* n_rows);

/* rpca, image_filtering.py:220-253, on a stack of n gray frames [n][h][w] (column k of the
 * reference's matrix = frame k, i.e. pass the frames in the order the reference would):
 * out[k] = clip(-E, 0, 255) as uint8; iters = IALM iterations taken (may be NULL). */
int swb_stage_rpca(int32_t device, const uint8_t* frames, int32_t n, int32_t h, int32_t w,
                   uint8_t* out, int32_t* iters);
/* bilateral_blur, image_filtering.py:304-307 = cv2.bilateralFilter(frame, d, sigma_color,
 * sigma_space) for an 8-bit single-channel image, d <= 7. */
int swb_stage_bilateral(int32_t device, const uint8_t* in, int32_t h, int32_t w, int32_t d,
                        double sigma_color, double sigma_space, uint8_t* out);
/* SWB_BG_RPCA only: IALM iterations the last submit took (image_filtering.py:281-298), the Jacobi sweeps its
 * eigenproblems needed in all, and whether the iteration loop ran on the device (1: a CUDA-graph WHILE node,
 * the 21-frame batch of the reference; 0: the host loop, other batch sizes).  Waits for the submit. */
int swb_rpca_stats(swb_ctx* ctx, int32_t* iterations, int32_t* jacobi_sweeps, int32_t* device_loop);
/* SWB_BG_RPCA only: the "RPCA" images (clip(-E, 0, 255), uint8, ROI-sized) of frames
 * [t0, t0 + n) of the last submit. */
int swb_get_rpca(swb_ctx* ctx, int32_t t0, int32_t n, uint8_t* dst, int32_t mem_kind);

/* Tracker cost matrix: formulate_cost_matrix, segment_tracking.py:46-102, for two
 * consecutive frames' segments on the GPU.  A workspace owns its device / page-locked buffers (sized for
 * max_segments = the most segments two consecutive frames may hold together); no allocation per call.
 * prev_yx / curr_yx: centroids (row, col) as float64 pairs; first_yx[i]: the centroid of
 * segment_history[0] of previous segment i, read only where has_history[i] != 0 (:217-222).
 * *matrix: the (n_prev + n_curr)^2 row-major float64 cost matrix in page-locked memory owned by the
 * workspace (valid until the next call): match block 0.5 * 2^(dist - 25) + 0.5 * angle cost (:190-243),
 * diagonal 1 (:246-250), everything else 1 + DBL_EPSILON (:179-187).  The Hungarian step
 * (scipy.optimize.linear_sum_assignment, :253-260) stays with the caller. */
typedef struct swb_tracker swb_tracker;
int swb_tracker_create(int32_t device, int32_t max_segments, swb_tracker** out);
int swb_tracker_destroy(swb_tracker* t);
int swb_tracker_costs(swb_tracker* t, const double* prev_yx, const double* first_yx,
                      const uint8_t* has_history, int32_t n_prev, const double* curr_yx, int32_t n_curr,
                      double** matrix);
const char* swb_tracker_last_error(const swb_tracker* t);
int64_t swb_tracker_launch_count(const swb_tracker* t);

/* Host-side helper of the drop-in (no device involved): copies n tiles of rows x row_bytes bytes, tile i starting at
 * host address src[i] with `pitch` bytes between its rows, into dst[n][rows][row_bytes].  FrameQueue.segment_queue
 * cuts the colour crops of all segments of a batch (extract_segment_images, image_filtering.py:338-369: views into the
 * host frames) with one call instead of one numpy slice-and-copy per segment. */
int swb_host_gather_tiles(const uint64_t* src, int64_t pitch, int32_t rows, int32_t row_bytes, int64_t n, uint8_t* dst);

/* Glue of the batched segment classifier (segment_classification.py:13-45: every segment image goes through the
 * reference's SqueezeNet; here all segments of a batch go through it at once, each layer evaluated only on the window
 * of positions the 24x24 crop can influence).  The convolutions are library kernels; these two kernels move a layer's
 * output into the next layer's input patch.  Device pointers, float32, NHWC (channels-last) layout, launched on
 * `stream` (a cudaStream_t, e.g. torch's current stream; NULL = the default stream), asynchronous.
 * swb_nhwc_paste:   dst[b][off_y + y][off_x + x][c] = src[b][y][x][c] (+ bias[c] when bias != NULL, then ReLU when
 *                   relu != 0); src is [batch][h][w][channels], dst is [batch][dst_h][dst_w][channels]; src == dst with
 *                   the same geometry and zero offsets is the in-place bias + ReLU of a convolution output.
 * swb_nhwc_maxpool: out[b][oy][ox][c] = max of the kernel x kernel window at (oy * stride, ox * stride) of
 *                   in[batch][in_h][in_w][channels]; out is [batch][(in_h - kernel) / stride + 1][(in_w - kernel) / stride + 1]
 *                   [channels]; channels must be a multiple of 4 and both pointers 16-byte aligned. */
int swb_nhwc_paste(const float* src, float* dst, int64_t batch, int32_t channels, int32_t h, int32_t w, int32_t dst_h,
                   int32_t dst_w, int32_t off_y, int32_t off_x, const float* bias, int32_t relu, void* stream);
int swb_nhwc_maxpool(const float* in, float* out, int64_t batch, int32_t channels, int32_t in_h, int32_t in_w,
                     int32_t kernel, int32_t stride, void* stream);

/* Page-locked host memory for frame ingest (io_video.py:11-165 decodes frames into host
 * arrays; frames decoded into these buffers reach the device by DMA at full PCIe speed
 * and asynchronously).  Portable across devices.  Needs a CUDA device like everything else. */
int swb_host_alloc(void** ptr, uint64_t bytes);
int swb_host_free(void* ptr);

/* Synthetic video (bench / tests): frames t0..t0+n-1 of the seeded generator,
 * bit-identical to oracle/synth.py.  dst: [n][h][w][3] uint8. */
int swb_synth_frames(int32_t device, uint8_t* dst, int32_t mem_kind, uint32_t seed, uint32_t video,
                     int32_t t0, int32_t n, int32_t h, int32_t w, int32_t n_birds);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H */
