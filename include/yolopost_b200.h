/*
 * yolopost_b200 - C-ABI of the B200-native YOLO detection post-processing path.
 *
 * Drop-in boundary.  The reference (Chriz122/ultralytics_pro, pure Python) has no
 * FFI layer for this path: its boundary is Python attribute binding of
 *     ultralytics/nn/modules/head.py:151   Detect._inference(self, x)            -> ypb_decode_dense
 *     ultralytics/nn/modules/head.py:184   Detect.decode_bboxes(...)             -> ypb_decode_dense
 *     ultralytics/nn/modules/head.py:1026  OBB.forward / :1040 OBB.decode_bboxes -> ypb_decode_dense (angle)
 *     ultralytics/utils/nms.py:13          non_max_suppression(prediction, ...)  -> ypb_nms_from_dense
 *     (head.py:151 + nms.py:13 back to back, the `postprocess` timer envelope)   -> ypb_nms_from_head
 *     ultralytics/utils/nms.py:239/187/299 TorchNMS.nms / fast_nms / batched_nms -> ypb_nms_boxes
 * The ctypes binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer except `ypb_*_desc*` / `ypb_nms_params*` (host structs) is a DEVICE pointer owned by the caller
 *     (torch caching allocator); the library allocates nothing and frees nothing;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises;
 *   - return value: 0 = OK, negative = error (see ypb_status); ypb_last_error_string() gives the text for the
 *     calling thread;
 *   - there is no CPU fallback: the library needs a CUDA device of compute capability 10.x.
 */
#ifndef YOLOPOST_B200_H_
#define YOLOPOST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define YPB_API __attribute__((visibility("default")))
#else
#define YPB_API
#endif

#define YPB_ABI_VERSION 11
#define YPB_MAX_LEVELS 8
#define YPB_MAX_PEERS 8 /* GPUs of one NVSwitch node */

typedef enum { YPB_F32 = 0, YPB_F16 = 1, YPB_BF16 = 2 } ypb_dtype;

typedef enum {
  YPB_OK = 0,
  YPB_ERR_INVALID_ARGUMENT = -1,
  YPB_ERR_UNSUPPORTED = -2,
  YPB_ERR_WORKSPACE_TOO_SMALL = -3,
  YPB_ERR_CUDA = -4
} ypb_status;

/* Suppression rule (utils/nms.py). */
typedef enum {
  YPB_NMS_GREEDY = 0,       /* nms.py:239-296 == torchvision.ops.nms: suppress iff inter/(a_i+a_j-inter) > thr, no eps */
  YPB_NMS_FAST_PROBIOU = 1, /* nms.py:187-236 with metrics.py:251 batch_probiou: drop j iff ANY higher-ranked i has piou >= thr */
  YPB_NMS_FAST_BOXIOU = 2   /* nms.py:187-236 with metrics.py:54 box_iou (eps 1e-7 in the denominator), >= thr */
} ypb_nms_rule;

/* Detect-head raw output: L levels of (B, 4*reg_max+nc, H_l, W_l), anchors (H_l*W_l) contiguous.
 * head.py:121-122 builds exactly these tensors; head.py:162 is the concat this library never materialises. */
typedef struct {
  int32_t num_levels;
  int32_t batch;
  int32_t nc;      /* classes */
  int32_t reg_max; /* DFL bins per side (block.py:232); 1 = no DFL (raw ltrb) */
  int32_t dtype;   /* ypb_dtype of the level tensors */
  int32_t reserved;
  const void* level_ptr[YPB_MAX_LEVELS];
  int32_t level_h[YPB_MAX_LEVELS];
  int32_t level_w[YPB_MAX_LEVELS];
  int64_t level_batch_stride[YPB_MAX_LEVELS];   /* elements */
  int64_t level_channel_stride[YPB_MAX_LEVELS]; /* elements; pixel stride must be 1 */
  float level_stride[YPB_MAX_LEVELS];           /* model stride of the level (head.py:79 self.stride) */
} ypb_head_desc;

/* Decoded prediction tensor (B, 4+nc+extra, A) as consumed by nms.py:13 - any strides (models/nas/predict.py:54
 * passes a permuted view). */
typedef struct {
  const void* ptr;
  int32_t dtype;
  int32_t batch;
  int32_t channels; /* 4 + nc + extra */
  int32_t anchors;
  int64_t stride_b, stride_c, stride_a; /* elements */
  /* anchor_subset: NULL, or device (B, subset_len) int64 anchor indices in [0, anchors): ONLY these anchors are candidates (the
   * second stage of the end2end top-k, head.py:209-211, ranks the pairs of the K anchors kept by the first stage - read where
   * they lie, no gathered copy).  Row ids, kept indices and the tie order stay those of the full tensor. */
  const int64_t* anchor_subset;
  int32_t subset_len;
  int32_t reserved;
} ypb_dense_desc;

/* Arguments of non_max_suppression (nms.py:13-29) that reach the device. */
typedef struct {
  float conf_thres;     /* already rounded to the prediction dtype by the caller (torch casts the scalar, nms.py:76) */
  float iou_thres_eff;  /* GREEDY: largest float <= iou_thres (double compare in torchvision); FAST_*: (float)iou_thres */
  int32_t nc;
  int32_t extra;        /* channels after the class scores (mask coeffs, keypoints, angle) */
  int32_t max_det;      /* nms.py:157 */
  int32_t max_nms;      /* nms.py:137 */
  float max_wh;         /* nms.py:143; 0 when agnostic */
  int32_t multi_label;  /* nms.py:82,114 (already AND-ed with nc > 1) */
  int32_t rule;         /* ypb_nms_rule; FAST_PROBIOU == rotated=True (boxes stay xywh + angle = last channel) */
  int32_t rows_cap;     /* candidate rows reserved per image in the workspace (A, or A*nc for multi_label) */
  const uint32_t* class_mask; /* device bitmask of allowed classes (nms.py:127-131) or NULL */
  /* The exporter's NMSModel flavour (engine/exporter.py:1389-1481; ypb_nms_from_dense only):
   *   nms_box_divisor > 0 : suppression runs on multiplier * (box / divisor) [+ cls * max_wh, with max_wh = multiplier]
   *                         (exporter.py:1437-1452: boxes normalised by the larger image side, class offset in units of 1/nc);
   *   boxes_xyxy          : columns 0..3 of the prediction are already corners (the export decode, head.py:189) - no nms.py:86;
   *   pad_output          : rows past the kept count are written as zeros (exporter.py:1478-1479 zero padding), their
   *                         indices (ypb_nms_out.idx) as -1. */
  float nms_box_divisor;
  float nms_box_multiplier;
  int32_t boxes_xyxy;
  int32_t pad_output;
  /* Optional device array of B per-image confidence thresholds replacing conf_thres (ypb_nms_from_dense only): the
   * second pass of the end2end top-k (head.py:193-214 Detect.postprocess) thresholds each image at its own K-th score. */
  const float* conf_per_image;
  /* Optional device int32[B + 1] (per-image row counters + the octet counter of the fused path), ZERO on entry: the calls
   * then use it instead of the workspace's own counters and skip their cudaMemset node - the suppression kernel, the last
   * reader, leaves it zeroed again ("clean on exit"), so a plan that owns such an array launches kernels only.
   * NULL = the library clears its counters inside the workspace itself (one memset node per call). */
  int32_t* clean_counters;
  /* Class-scan kernel of the fused path (ypb_nms_from_head*): YPB_SCAN_LDG = one-wave grid of small CTAs with register-staged,
   * software-pipelined 128-bit loads - shares every SM with the kernels of other streams, the form a multi-stream pipeline
   * wants; YPB_SCAN_TMA = persistent one-CTA-per-SM kernel fed by TMA tensor loads through a shared-memory ring - the
   * faster single kernel (16-bit heads especially) when nothing else needs the SMs; YPB_SCAN_AUTO = LDG unless the
   * environment variable YPB_SCAN_TMA=1 is set. */
  int32_t scan_kernel;
  int32_t reserved2;
} ypb_nms_params;

/* Per-image letterbox transform, values exactly as the reference computes them on the host:
 *   utils/ops.py:120-127 (scale_boxes): gain = min(h1/h0, w1/w0) or ratio_pad[0][0]; pad_x/pad_y = round((w1-w0*gain)/2-0.1)
 *                                       or ratio_pad[1]; img_w/img_h = the target image size used by clip_boxes (:163)
 *   utils/ops.py:580-587 (scale_coords): cpad_x/cpad_y = (w1-w0*gain)/2 un-rounded, or ratio_pad[1]
 * every member already rounded to fp32 (what torch does to a Python scalar operand of an fp32 tensor op). */
typedef struct {
  float gain, pad_x, pad_y, img_w, img_h, cpad_x, cpad_y, reserved;
} ypb_scale_xform;

typedef enum {
  YPB_BOXES_NONE = 0,            /* leave columns 0..3 alone */
  YPB_BOXES_XYXY = 1,            /* ops.py:102-135 scale_boxes(xywh=False): pad all four, / gain, clip_boxes */
  YPB_BOXES_XYWH = 2,            /* scale_boxes(xywh=True): pad x,y only, / gain, no clip (obb/predict.py:60, obb/val.py:223) */
  YPB_BOXES_XYWHR = 3,           /* ops.py:621-636 regularize_rboxes on (x,y,w,h,angle@angle_col), then XYWH (obb/predict.py:59-60) */
  YPB_BOXES_CLIP_ONLY = 4,       /* ops.py:152-177 clip_boxes */
  YPB_BOXES_REGULARIZE_ONLY = 5  /* ops.py:621-636 regularize_rboxes */
} ypb_box_mode;

#define YPB_SCALE_PADDING 1          /* ops.py:128 `padding=True` */
#define YPB_SCALE_NORMALIZE 2        /* ops.py:591-594 `normalize=True` (coords only) */
#define YPB_SCALE_COORDS_CLIP_ONLY 4 /* ops.py:598-618 clip_coords alone */

/* Result buffers (device).  rows: (B, max_det, 6+extra) fp32 = x1,y1,x2,y2,conf,cls,extra... (xywh + angle when
 * rotated); idx: (B, max_det) int64 anchor index of each kept row (nms.py:161 keepi) or NULL; count: (B) int32;
 * cand_count: (B) int32 rows that passed the confidence filter before any cap (diagnostic) or NULL.
 * scale_xforms: NULL, or a device array of B transforms: the kept rows are written already rescaled to the original
 * image - detect/predict.py:120 scale_boxes (greedy rule) or obb/predict.py:59-60 regularize_rboxes + scale_boxes(xywh=True)
 * (FAST_PROBIOU rule) fused into the gather; scale_padding = ops.py:128 `padding`. */
typedef struct {
  float* rows;
  int64_t* idx;
  int32_t* count;
  int32_t* cand_count;
  const ypb_scale_xform* scale_xforms;
  int32_t scale_padding;
  /* One-sided result gather over NVLink peer memory (multi-GPU): when num_peers > 0 the gather stage of
   * the suppression kernel also stores this rank's kept rows and counts straight into every peer's gather buffer
   * (peer-mapped device pointers, e.g. torch symmetric memory / CUDA IPC) - the analogue of dist.gather_object(stats),
   * detect/val.py:226-240, without a collective: no rank ever waits for another's KERNEL inside the step.
   * Every rank's gather buffer is a RING of peer_depth entries, each entry = num_peers slots of one packed result
   * ((B, max_det, cols) rows then (B) counts); launch number seq (1, 2, ...) of a rank goes to entry seq % peer_depth.
   *   peer_rows[p] / peer_count[p] : where THIS rank's rows / counts live in ENTRY 0 of peer p's ring (p = my_rank: the local ring)
   *   peer_entry_stride            : floats between consecutive ring entries
   *   peer_flag[p]                 : peer p's arrival flags, one int32 per rank; flag[my_rank] is set to the launch
   *                                  sequence number (system-scope release) once all images of the launch are stored
   *   peer_ack                     : LOCAL int32 per rank, written by the peers' ypb_peer_wait: ack[p] = last launch whose
   *                                  entry peer p has released.  The kernel does not overwrite entry seq % depth in peer p before
   *                                  ack[p] >= seq - peer_depth (back-pressure; NULL = none, single-consumer diagnostics only)
   *   peer_state                   : local device int32[4], zero-initialised once by the caller ([0] CTAs done,
   *                                  [1] launches sent so far = the sequence number ypb_peer_wait waits for,
   *                                  [2] last batch handed to the consumer, [3] protocol overrun marker: both spin loops
   *                                  are bounded (seconds) so a missing consumer / dead peer cannot hang the GPU) */
  int32_t num_peers;
  int32_t my_rank;
  int32_t peer_depth;
  float* peer_rows[YPB_MAX_PEERS];
  int32_t* peer_count[YPB_MAX_PEERS];
  int32_t* peer_flag[YPB_MAX_PEERS];
  int32_t* peer_state;
  const int32_t* peer_ack;
  int64_t peer_entry_stride;
  /* count_host: NULL, or the device-accessible address of (B) int32 in pinned (mapped) HOST memory.  The suppression kernel
   * then stores every per-image count there as well: the one quantity of the call a host needs before it can cut the rows
   * (nms.py:159-161 returns a list of (n_i, 6+extra) tensors) arrives without a copy node behind the kernels; it is valid
   * once the stream has been synchronised (or an event recorded behind the call has completed). */
  int32_t* count_host;
} ypb_nms_out;

YPB_API int ypb_abi_version(void);
YPB_API const char* ypb_last_error_string(void);

/* Bytes of scratch needed by ypb_nms_from_head / ypb_nms_from_dense for this geometry. */
YPB_API size_t ypb_nms_workspace_bytes(int32_t batch, int32_t anchors, int32_t rows_cap, int32_t max_det, int32_t max_nms,
                               int32_t rule);

/* Dense decode == Detect._inference (head.py:151-169): writes (B, 4+nc[+1], A) in `out_dtype`.
 *   angle            : NULL, or (B, A) contiguous rotation channel of the OBB head (head.py:1028)
 *   angle_is_logit   : 1 = raw cv4 output, the kernel applies (sigmoid-0.25)*pi (head.py:1031); 0 = already activated
 *   append_angle     : 1 = also write the activated angle as channel 4+nc (head.py:1038)
 *   xyxy             : 1 = corners instead of cx,cy,w,h (head.py:189 end2end / self.xyxy); ignored when angle != NULL */
YPB_API int ypb_decode_dense(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t append_angle,
                     int32_t xyxy, void* out, int32_t out_dtype, int64_t out_stride_b, int64_t out_stride_c,
                     void* stream);

/* The two pieces of the decode a head may call on its own (YOLOEDetect.forward_lrpc, head.py:1777-1813, does):
 *   DFL.forward (nn/modules/block.py:250-253): x (B, 4*reg_max, A) -> out (B, 4, A) = sum_k k * softmax_k, same dtype;
 *     channel index = side*reg_max + bin; anchors contiguous (stride 1), strides in elements.
 *   Detect.decode_bboxes (head.py:184-191) == dist2bbox(dim=1) (utils/tal.py:367-376): dist (B, 4, A) l,t,r,b and
 *     anchor_points (1 or B, 2, A; any strides - the module caches a transposed (A, 2) tensor,
 *     head.py:164; anchor_stride_b = 0 broadcasts) -> out (B, 4, A) cx,cy,w,h (xywh=1) or x1,y1,x2,y2;
 *     with angle != NULL ((B, 1, A), already activated) it is OBB.decode_bboxes (head.py:1040-1042) == dist2rbox(dim=1)
 *     (utils/tal.py:385-403) and xywh is ignored.  Every step is rounded to `dtype` like the reference's tensor ops. */
YPB_API int ypb_dfl_expectation(const void* x, int32_t dtype, int32_t batch, int32_t reg_max, int32_t anchors,
                                int64_t stride_b, int64_t stride_c, void* out, int64_t out_stride_b, int64_t out_stride_c,
                                void* stream);
YPB_API int ypb_dist2bbox(const void* dist, int64_t dist_stride_b, int64_t dist_stride_c, const void* anchor_points,
                          int64_t anchor_stride_b, int64_t anchor_stride_c, int64_t anchor_stride_a, const void* angle, int64_t angle_stride_b,
                          int32_t dtype, int32_t batch, int32_t anchors, int32_t xywh, void* out, int64_t out_stride_b,
                          int64_t out_stride_c, void* stream);

/* Fused decode -> confidence filter -> sort/top-k -> suppression -> gather, reading the head once and never writing
 * the dense tensor.  `value_dtype` is the dtype the dense tensor WOULD have had (scores and boxes are rounded to it
 * so results are bit-identical to ypb_decode_dense followed by ypb_nms_from_dense). */
YPB_API int ypb_nms_from_head(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t value_dtype,
                      const ypb_nms_params* p, const ypb_nms_out* out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Riders of the fused path: per-anchor channels produced by a sibling branch of the head that travel with the kept
 * rows as columns 6.. (nms.py:112 `mask`): Segment mask coefficients (head.py:831,837: (B, nm, A)) are copied, Pose
 * keypoints (head.py:1248: raw (B, nk*ndim, A)) are decoded on the way (head.py:1254-1273 kpts_decode) - for the kept
 * anchors only, the dense (B, nk*ndim, A) decode of the reference is never computed. */
typedef enum { YPB_RIDER_RAW = 0, YPB_RIDER_KEYPOINTS = 1 } ypb_rider_kind;
typedef enum { YPB_SCAN_AUTO = 0, YPB_SCAN_LDG = 1, YPB_SCAN_TMA = 2 } ypb_scan_kernel;
typedef struct {
  const void* ptr;     /* (B, channels, A), anchors contiguous, same dtype as the head */
  int32_t channels;    /* == ypb_nms_params.extra */
  int32_t kind;        /* ypb_rider_kind */
  int32_t kpt_ndim;    /* 2 | 3 (kpt_shape[1]) for YPB_RIDER_KEYPOINTS */
  int32_t reserved;
  int64_t stride_b, stride_c; /* elements */
} ypb_riders_desc;

YPB_API int ypb_nms_from_head_riders(const ypb_head_desc* head, const ypb_riders_desc* riders, int32_t value_dtype,
                                     const ypb_nms_params* p, const ypb_nms_out* out, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* Pose.kpts_decode (head.py:1254-1273) on the whole (B, nk*ndim, A) tensor: out has the same shape/dtype, contiguous.
 * Level geometry (grid sizes, strides) is taken from `head` (level pointers are not read). */
YPB_API int ypb_kpts_decode(const ypb_head_desc* head, const void* kpts, int64_t stride_b, int64_t stride_c,
                            int32_t channels, int32_t kpt_ndim, void* out, void* stream);

/* process_mask / process_mask_native (utils/ops.py:489-541), batched over the kept rows of a batch.
 *   protos        : (B, C, mh, mw) prototype masks of the Segment head (head.py:830), fp32/fp16/bf16, pixels contiguous
 *   coeffs, boxes : per detection C mask coefficients (columns 6.. of the NMS rows) and the xyxy box (columns 0..3), fp32;
 *                   detection r of image b at ptr + b*image_stride + r*row_stride (elements)
 *   offsets       : device (B+1) int32 exclusive prefix of the per-image detection counts, or NULL when batch == 1
 *   total         : detections in the batch (== offsets[B]); out is (total, out_h, out_w) uint8, packed in image order
 *   window        : rows [win_top, win_top+win_h) x cols [win_left, win_left+win_w) of the prototype grid are resized to
 *                   (out_h, out_w) - the whole grid for process_mask, the un-padded part for scale_masks (ops.py:544-559)
 *   crop_mode     : YPB_MASK_CROP_PROTO  = crop at prototype resolution with boxes*(ratio_w, ratio_h) BEFORE resizing
 *                                          (process_mask, ops.py:505-510; out == (mh, mw) gives upsample=False)
 *                   YPB_MASK_CROP_OUTPUT = crop at output resolution with the boxes as given, AFTER resizing
 *                                          (process_mask_native, ops.py:538-540) */
typedef enum { YPB_MASK_CROP_PROTO = 1, YPB_MASK_CROP_OUTPUT = 2 } ypb_mask_crop;
typedef struct {
  const void* ptr;
  int32_t dtype;
  int32_t channels, mh, mw;
  int64_t stride_b, stride_c; /* elements */
} ypb_protos_desc;
YPB_API int ypb_process_mask(const ypb_protos_desc* protos, const float* coeffs, int64_t coef_image_stride,
                             int64_t coef_row_stride, const float* boxes, int64_t box_image_stride, int64_t box_row_stride,
                             const int32_t* offsets, int32_t batch, int32_t total, int32_t out_h, int32_t out_w,
                             int32_t win_top, int32_t win_left, int32_t win_h, int32_t win_w, int32_t crop_mode,
                             float ratio_w, float ratio_h, uint8_t* out, void* workspace, size_t workspace_bytes,
                             void* stream);
/* workspace (optional): ypb_process_mask_workspace_bytes(total, out_h, out_w) bytes of device scratch.  With it the call
 * builds a work list of the (detection, 128x128 tile) pairs that can see their box (typically ~15 % of the tiles) and one kernel
 * computes only those while an extra warp of every CTA zero-fills all other tiles with streaming stores - the write of the
 * (total, out_h, out_w) result overlaps the arithmetic (out_w % 16 == 0; otherwise: cudaMemset of the result, then the listed
 * tiles); without it (NULL) one kernel zero-fills and computes every tile. */
YPB_API size_t ypb_process_mask_workspace_bytes(int32_t total, int32_t out_h, int32_t out_w);

/* Validator matching: engine/validator.py:267-307 match_predictions (non-scipy branch) with, optionally, the pairwise
 * metrics.py:54 box_iou of detect/val.py:287 computed on the fly.  correct: (B, rows_per_image, nthr) uint8.
 *   boxes mode : preds = result rows (x1,y1,x2,y2 in columns 0..3, class in column cls_col), labels = (sum M, 5) fp32
 *                rows cls,x1,y1,x2,y2 of all images, label_offsets = device (B+1) int32 prefix (NULL: one image, m labels)
 *   matrix mode: iou = (m, n) fp32 matrix iou[l*iou_stride + d] of ONE image (obb/val.py, segment/val.py, pose/val.py
 *                compute their own), preds = the n predicted classes (pred_row_stride 1, cls_col 0), true_cls = (m)
 *   thresholds : HOST array of nthr <= 16 IoU levels (validator iouv); count: device (B) kept rows or NULL
 *   workspace  : needed only when nthr * max_labels * 4 exceeds 200 KB of shared memory: nthr * sum M int32 */
YPB_API int ypb_match_predictions(const float* preds, int64_t pred_image_stride, int64_t pred_row_stride, int32_t cls_col,
                                  int32_t batch, int32_t rows_per_image, const int32_t* count, const float* labels,
                                  const int32_t* label_offsets, int32_t m, int32_t max_labels, const float* iou,
                                  int64_t iou_stride, const float* true_cls, const float* thresholds, int32_t nthr,
                                  uint8_t* correct, void* workspace, size_t workspace_bytes, void* stream);

/* Consumer side of the one-sided gather: enqueues a tiny kernel that (1) RELEASES the ring entry the previous ypb_peer_wait
 * of this lane handed out - every read the consumer enqueued since is stream-ordered before it - by writing its sequence
 * number into ack[my_rank] of every producer (peer_ack[p] = peer p's acknowledgement array, peer-mapped; NULL = no acks),
 * then (2) spins (system-scope acquire) until every one of the `world` arrival flags of THIS rank has reached the sequence
 * number of this rank's own latest launch minus `lag` (state[1]; all ranks run the same launch sequence), i.e. until the
 * results of that launch of every rank have landed here, and (3) writes that entry's index (seq % depth) to *slot_index
 * (device int64, may be NULL) for the consumer's kernels.  lag > 0 is a pipelined gather: the step never stalls on a slower
 * rank; it needs depth >= lag + 2.  lag < 0 = IN ORDER: the launch after the one handed out last (state[2] + 1), whatever this
 * rank has launched since - the form for a consumer that runs CONCURRENTLY with the next step's kernels (a forked branch of
 * the step's CUDA graph: the gather of step q-1 is then consumed beside, not behind, the kernels of step q).
 * The entry stays valid until the NEXT ypb_peer_wait on this lane executes. */
YPB_API int ypb_peer_wait(const int32_t* flags, int32_t world, int32_t* state, int32_t lag, int32_t depth,
                          int32_t* const* peer_ack, int32_t my_rank, int64_t* slot_index, void* stream);

/* ypb_peer_wait with the simplest complete consumer attached, in ONE kernel: after the wait, the returned ring entry
 * (`entry_floats` floats at ring + index * entry_floats; a multiple of 4, 16-byte aligned) is copied to `out`. */
YPB_API int ypb_peer_wait_copy(const int32_t* flags, int32_t world, int32_t* state, int32_t lag, int32_t depth,
                               int32_t* const* peer_ack, int32_t my_rank, int64_t* slot_index, const float* ring,
                               int64_t entry_floats, float* out, void* stream);

/* Same call restricted to some of its kernels, for per-kernel timing with CUDA events (bench.py roofline) and
 * profiling.  `stage` is a bit mask: 1 = clear counters + class-scan/filter/compaction kernel, 2 = survivor box-decode
 * kernel, 4 = sort + suppression + gather kernel (each on what the earlier stages left in `workspace`);
 * 0 or 7 = all three (== ypb_nms_from_head); 8 (with 1) = do not clear the counters first (kernel-only timing of the scan: the
 * caller clears them once per batch of launches and accepts that row slots accumulate). */
YPB_API int ypb_nms_from_head_stage(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit,
                                    int32_t value_dtype, const ypb_nms_params* p, const ypb_nms_out* out,
                                    void* workspace, size_t workspace_bytes, void* stream, int32_t stage);

/* non_max_suppression on an already decoded tensor (nms.py:13-166). */
YPB_API int ypb_nms_from_dense(const ypb_dense_desc* pred, const ypb_nms_params* p, const ypb_nms_out* out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* TorchNMS.nms / fast_nms on one box set (nms.py:187-296): boxes (n, 4|5) fp32 row-major, scores (n) fp32.
 * keep: (n) int64 indices in descending-score order, keep_count: (1) int32.  Workspace: ypb_nms_boxes_workspace_bytes(n). */
YPB_API size_t ypb_nms_boxes_workspace_bytes(int32_t n);
YPB_API int ypb_nms_boxes(const float* boxes, const float* scores, int32_t n, int32_t box_dim, int32_t rule,
                  float iou_thres_eff, int64_t* keep, int32_t* keep_count, void* workspace, size_t workspace_bytes,
                  void* stream);

/* Packs the kept rows of a batch back to back in image order (the list-of-tensors return type of nms.py:159-161 is then
 * ONE split on the host): rows (B, max_det, cols) fp32 / idx (B, max_det) int64 / count (B) as written by the NMS calls ->
 * out_rows (>= sum count, cols), out_idx (>= sum count), out_offsets (B+1) exclusive prefix of count; any output may be NULL. */
YPB_API int ypb_compact_results(const float* rows, const int64_t* idx, const int32_t* count, int32_t batch, int32_t max_det,
                                int32_t cols, float* out_rows, int64_t* out_idx, int32_t* out_offsets, void* stream);

/* utils/metrics.py:54-75 box_iou (box_dim 4: x1,y1,x2,y2; eps 1e-7 in the denominator) and metrics.py:251-284
 * batch_probiou (box_dim 5: x,y,w,h,r) as free functions: boxes1 (n, box_dim), boxes2 (m, box_dim) fp32 row-major ->
 * out (n, m) fp32.  The suppression kernels evaluate the same device functions pair by pair without this matrix. */
YPB_API int ypb_pairwise_iou(const float* boxes1, int32_t n, const float* boxes2, int32_t m, int32_t box_dim, float* out,
                             void* stream);

/* ---- result-side steps that follow NMS in the reference's predictors / validators -------- */

/* In-place rescale of result rows from the letterboxed network input to the original image.
 *   rows            : fp32, row r of image b at rows + b*image_stride + r*row_stride (elements); columns 0..3 = box
 *   count           : device (B) int32 kept rows per image (ypb_nms_out.count) or NULL = all rows_per_image rows
 *   xforms / xform  : device array of B transforms, or NULL to use the single host `xform` for every image
 *   coords,nk,ndim  : optional keypoints riding on each row (pose/predict.py:73-75): nk points of ndim (2|3) floats, the first
 *                     at coords + b*coord_image_stride + r*coord_row_stride; scaled like ops.py:562-595 scale_coords */
YPB_API int ypb_scale_rows(float* rows, int64_t image_stride, int64_t row_stride, int32_t batch, int32_t rows_per_image,
                           const int32_t* count, const ypb_scale_xform* xforms, const ypb_scale_xform* xform,
                           int32_t box_mode, int32_t flags, int32_t angle_col, float* coords,
                           int64_t coord_image_stride, int64_t coord_row_stride, int32_t nk, int32_t ndim, void* stream);

/* Diagnostic: when set to a device buffer of batch*32 int64, the sort+suppress kernel stores clock64() marks per CTA
 * (0 start, 1 ranked, 2+i after chunk i, 30 before gather, 31 end); NULL disables (default). */
YPB_API void ypb_debug_set_phase_buffer(void* device_buffer);

/* Diagnostic: exhaustively checks that the device sigmoid used by the decode kernels is monotone non-decreasing
 * over all finite fp32 inputs after rounding to `dtype`; writes the number of violations to *violations (device). */
YPB_API int ypb_selftest_sigmoid_monotone(int32_t dtype, unsigned long long* violations, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLOPOST_B200_H_ */
