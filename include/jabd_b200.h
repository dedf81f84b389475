/*
 * jabd_b200.h -- C-ABI of libjabd_b200.so: the JABD box-geometry hot path on B200 (sm_100a).
 *
 * The reference (R/ = JABD2080ti/, pure Python/PyTorch) has no FFI of its own; its boundary for this
 * path is a set of Python functions.  Each entry point below names the reference function(s) it
 * replaces (R/file:line); the Python package `jabd_b200` re-exports them under the reference's names.
 *
 * Conventions
 *   - Every pointer is DEVICE memory on the current CUDA device unless the name ends in `_host`.
 *   - fp32 data, C-contiguous.  `conf_t` is int64 (torch.LongTensor, R/nets/retinaface_training.py:199).
 *   - The caller owns every buffer, including the workspace; the library never allocates or frees
 *     device memory (one explicit exception: jabd_p2p_alloc / jabd_p2p_free, IPC-exportable buffers of the
 *     peer-memory exchange), keeps no pointer after return and has no global mutable state: every option is an
 *     argument of the call it affects.  (Process-wide memo tables of *device properties* -- shared-memory opt-in
 *     done, resident clusters per width -- are atomics and never depend on a call's arguments.)
 *   - Test and bench hooks (division self-test, FP32 probe) are NOT part of this ABI: they live in
 *     libjabd_b200_selftest.so / include/jabd_b200_selftest.h.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises unless
 *     its comment says so.  Calls are re-entrant and may run concurrently on distinct streams.
 *   - Return 0 on success or a negative JABD_E* code; `jabd_last_error()` gives a thread-local message.
 *   - There is no CPU implementation behind any of these symbols.
 */
#ifndef JABD_B200_H
#define JABD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define JABD_API __attribute__((visibility("default")))
#else
#define JABD_API
#endif

typedef void *jabd_stream_t; /* cudaStream_t */

enum {
    JABD_OK = 0,
    JABD_EINVAL = -1,    /* bad argument (null pointer, negative size, unsupported option) */
    JABD_EALIGN = -2,    /* pointer not aligned as documented */
    JABD_EWORKSPACE = -3,/* workspace missing or too small */
    JABD_ECUDA = -4,     /* a CUDA runtime call failed; see jabd_last_error() */
    JABD_ENODEVICE = -5  /* no sm_100 device */
};

/* GT row layout of the reference data loader (R/utils/dataloader.py:37-58):
 * x1 y1 x2 y2 | 5 x (lx, ly) | label, normalised to [0,1]. */
#define JABD_GT_ROW 15
#define JABD_DET_ROW 15 /* x1 y1 x2 y2 score | 5 x (lx, ly)   (R/predict.py:180) */

/* label_mode: 0 = live training match, conf = label          (R/nets/retinaface_training.py:137)
 *             1 = SSD-legacy match,    conf = label + 1      (R/utils/box_utils.py:315)            */
/* encode_mode: 1 = SSD encode of the matched box             (R/nets/retinaface_training.py:148)
 *              0 = raw matched x1y1x2y2 (match_iou/_ious)    (R/nets/retinaface_training_DIOU.py:229,
 *                                                             R/utils/box_utils.py:229-273)         */
/* flags for jabd_assign */
#define JABD_ASSIGN_DENSE 1 /* evaluate every prior x GT pair (no spatial culling); same results */
#define JABD_ASSIGN_PREP_ONLY 2 /* jabd_assign_match only: launch the staging kernel alone (per-kernel timing) */
#define JABD_ASSIGN_ASYNC 4 /* jabd_assign_host only: do not synchronise; host buffers must be pinned and the caller
                               synchronises `stream` (or an event recorded on it) before reading the outputs */
#define JABD_ASSIGN_DEVICE_OUT 16 /* jabd_assign_host only: loc_t / conf_t / landm_t are DEVICE buffers -- the targets stay in
                                     HBM for the loss (what MultiBoxLoss.forward does with them, R/nets/retinaface_training.py
                                     :220-227); only gt / gt_off cross the bus */

/* Shape of the matching kernel's work list, a per-call option like the detect kernels' cluster width (results never depend
 * on it): the first coarse_pct % of the prior tiles (coarse pyramid levels first) are paired with GT segments of seg_a rows,
 * the remaining tiles with segments of seg_b rows; 16 <= seg <= 192.  No JABD_ASSIGN_TUNE bits: the library's choice -- 64 for
 * a call that runs alone (short items: short tail), 192 for batches that overlap on the lanes of jabd_assign_batches (fewer
 * items: less per-item work). */
#define JABD_ASSIGN_TUNE(seg_a, seg_b, coarse_pct) (((seg_a) << 8) | ((seg_b) << 16) | ((coarse_pct) << 24))

JABD_API int jabd_version(void);
JABD_API const char *jabd_last_error(void);
/* Fills SM count and compute capability of the current device.  Synchronous, host only. */
JABD_API int jabd_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- P1: prior boxes.  Replaces Anchors.get_anchors / Anchors_eval.get_anchors (R/utils/anchors.py:9-42,
 * :43-79).  steps_host[n_levels]; min_sizes_host[sizes_off_host[n_levels]] grouped per level by
 * sizes_off_host[n_levels+1].  Output [P,4] (cx,cy,w,h), float64 arithmetic rounded once to fp32, optional
 * clamp to [0,1]. */
JABD_API int64_t jabd_priors_count(const int *steps_host, const int *sizes_off_host, int n_levels, int H, int W);
JABD_API int jabd_priors(const int *steps_host, const double *min_sizes_host, const int *sizes_off_host, int n_levels,
                         int H, int W, int clip, float *out, int64_t P, jabd_stream_t stream);

/* ---- M1/M2/E1/E2/D1/D2 as stand-alone operators (the reference's public helpers) -------------------- */
/* point_form (R/nets/retinaface_training.py:8-10): [n,4] cxcywh -> [n,4] x1y1x2y2. */
JABD_API int jabd_point_form(const float *boxes, int64_t n, float *out, jabd_stream_t stream);
/* jaccard (R/nets/retinaface_training.py:41-59): dense IoU [A,B]; both inputs point-form. */
JABD_API int jabd_jaccard(const float *box_a, int64_t A, const float *box_b, int64_t B, float *out, jabd_stream_t stream);
/* intersect (R/nets/retinaface_training.py:22-39): dense intersection areas [A,B]. */
JABD_API int jabd_intersect(const float *box_a, int64_t A, const float *box_b, int64_t B, float *out, jabd_stream_t stream);
/* encode (R/nets/retinaface_training.py:61-70) and encode_landm (:72-84), n rows each. */
JABD_API int jabd_encode(const float *matched, const float *priors, int64_t n, float var0, float var1, float *out,
                         jabd_stream_t stream);
JABD_API int jabd_encode_landm(const float *matched, const float *priors, int64_t n, float var0, float *out,
                               jabd_stream_t stream);
/* decode (R/utils/utils_bbox.py:29-34) and decode_landm (:39-46).  `batch` images share `priors` [P,4];
 * loc is [batch,P,4] / pre is [batch,P,10]. */
JABD_API int jabd_decode(const float *loc, const float *priors, int64_t P, int batch, float var0, float var1, float *out,
                         jabd_stream_t stream);
JABD_API int jabd_decode_landm(const float *pre, const float *priors, int64_t P, int batch, float var0, float *out,
                               jabd_stream_t stream);

/* ---- T1/T2: batched target assignment.  Replaces the per-image loop of MultiBoxLoss.forward
 * (R/nets/retinaface_training.py:197-214) and match() (:93-162; SSD form R/utils/box_utils.py:276-320;
 * match_iou R/nets/retinaface_training_DIOU.py:176-246).
 *   priors [P,4] cxcywh, shared by the batch.
 *   gt [sumG,15] packed rows, gt_off [B+1] int32 image offsets (device).  An image with no GT gets
 *   all-zero targets (the reference's data loader never emits one, R/utils/dataloader.py:181-182).
 *   loc_t [B,P,4] f32, conf_t [B,P] i64, landm_t [B,P,10] f32 (NULL: skip, 8-arg SSD forms).
 *   Optional (NULL to skip): best_truth_idx [B,P] i32 and best_truth_overlap [B,P] f32 after the
 *   force-match override (:127-130), best_prior_idx [sumG] i32 / best_prior_overlap [sumG] f32 (:111).
 * Workspace: jabd_assign_workspace_bytes(), 256-byte aligned. */
JABD_API size_t jabd_assign_workspace_bytes(int B, int64_t P, int64_t sumG);
JABD_API int jabd_assign(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG,
                         float threshold, float var0, float var1, int label_mode, int encode_mode, int flags,
                         float *loc_t, int64_t *conf_t, float *landm_t, int *best_truth_idx, float *best_truth_overlap,
                         int *best_prior_idx, float *best_prior_overlap, void *workspace, size_t workspace_bytes,
                         jabd_stream_t stream);
/* Several independent batches in one call (a data loader's prefetch queue, gradient-accumulation micro-batches: the
 * reference calls MultiBoxLoss.forward once per batch, R/nets/retinaface_training.py:197-227, and nothing couples two
 * batches).  Batch i is enqueued on lanes[i % n_lanes] -- caller-owned streams, distinct from `stream` and from each other --
 * so that one batch's staging and encode kernels fill the SMs the persistent matching kernel of another leaves idle on its
 * ramp and tail (cfg2: 27 us per batch instead of 33).  Every lane is ordered after `stream` at entry and `stream` after
 * every lane at exit (events; nothing synchronises the host), so to the caller this behaves like n_batches jabd_assign calls
 * on `stream`; it may be captured into a CUDA graph through `stream`.  n_lanes == 0: the batches run back to back on
 * `stream`.  Batches on different lanes need their own outputs and workspaces (one lane is stream-ordered: a batch may reuse
 * the buffers of an earlier one there); `priors` and the scalar options are shared. */
typedef struct {
    const float *gt;      /* [sumG,15] */
    const int *gt_off;    /* [B+1] */
    int B;
    int64_t sumG;
    float *loc_t;         /* [B,P,4] */
    int64_t *conf_t;      /* [B,P] */
    float *landm_t;       /* [B,P,10] or NULL */
    void *workspace;      /* jabd_assign_workspace_bytes(B, P, sumG) */
    size_t workspace_bytes;
} jabd_assign_batch_t;
JABD_API int jabd_assign_batches(const float *priors, int64_t P, const jabd_assign_batch_t *batches_host, int n_batches,
                                 float threshold, float var0, float var1, int label_mode, int encode_mode, int flags,
                                 const jabd_stream_t *lanes_host, int n_lanes, jabd_stream_t stream);
/* The two phases of jabd_assign, exposed for per-kernel timing (bench.py roofline) and tests:
 * _match  = GT staging + IoU + both argmaxes into the workspace; _encode = force-match + gather + encode. */
JABD_API int jabd_assign_match(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG,
                               int flags, void *workspace, size_t workspace_bytes, jabd_stream_t stream);
JABD_API int jabd_assign_encode(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG,
                                float threshold, float var0, float var1, int label_mode, int encode_mode, float *loc_t,
                                int64_t *conf_t, float *landm_t, int *best_truth_idx, float *best_truth_overlap,
                                int *best_prior_idx, float *best_prior_overlap, void *workspace, size_t workspace_bytes,
                                jabd_stream_t stream);
/* Same call with HOST buffers (pageable or pinned): copies gt/gt_off in, runs jabd_assign, copies the
 * three target tensors out.  priors and the staging area stay on the device:
 * dev_scratch must hold jabd_assign_host_scratch_bytes().  Synchronises `stream` before returning unless
 * JABD_ASSIGN_ASYNC is set (two calls on two streams with two scratch areas then overlap copies and kernels).
 * With JABD_ASSIGN_DEVICE_OUT the three outputs are device pointers and nothing is copied back. */
/* Host-only helper for the list-of-arrays form the reference's data loader yields (R/utils/dataloader.py:37-58, one [G_i,15]
 * float32 array per image): copies the rows of B host arrays into one packed buffer and writes the B+1 offsets.
 * Returns sum(G) or a negative code (capacity_rows too small, null pointer, negative count). */
JABD_API int64_t jabd_pack_gt_rows(const float *const *rows, const int *counts, int B, float *gt_packed, int64_t capacity_rows,
                                   int *gt_off);
JABD_API size_t jabd_assign_host_scratch_bytes(int B, int64_t P, int64_t sumG, int with_landm);
/* offsets4 = {loc_t, conf_t, landm_t, total bytes}: the layout of the targets inside the staging area.  Host outputs that
 * are three views of one block with these offsets are copied back with ONE D2H transfer instead of three. */
JABD_API int jabd_assign_host_out_offsets(int B, int64_t P, int with_landm, size_t *offsets4);
JABD_API int jabd_assign_host(const float *priors_dev, int64_t P, const float *gt_host, const int *gt_off_host, int B,
                              float threshold, float var0, float var1, int label_mode, int encode_mode, int flags,
                              float *loc_t_host, int64_t *conf_t_host, float *landm_t_host, void *dev_scratch,
                              size_t dev_scratch_bytes, jabd_stream_t stream);

/* ---- box post-processing of Retinaface.detect_image on detection rows [B,K,15], in place:
 * letterbox != 0: (v - offset) * scale per x / y column (retinaface_correct_boxes, R/utils/utils_bbox.py:9-24);
 * to_pixels != 0: v * width / height (R/predict.py:194-195).  The score column is untouched.  fp64 arithmetic, one
 * rounding to fp32 per step, like the reference's numpy code.  post [B,6] doubles (device) =
 * {offset_x, offset_y, scale_x, scale_y, width, height}; counts [B] (NULL: all K rows of every image). */
JABD_API int jabd_correct_boxes(float *dets, const int *counts, const double *post, int B, int K, int letterbox,
                                int to_pixels, jabd_stream_t stream);

/* ---- the rest of MultiBoxLoss.forward (R/nets/retinaface_training.py:229-303) on the assigned
 * targets: positives, hard-negative mining (per-image radix select of the min(negpos_ratio*num_pos, P-1) largest
 * rank values instead of the reference's two full sorts, :270-281), smooth-L1 sums for boxes and landmarks,
 * cross-entropy over positives + mined negatives, normalisation.  num_classes == 2.
 *   loc_data [B,P,4], conf_data [B,P,2] (logits), landm_data [B,P,10]: the network outputs;
 *   loc_t / conf_t / landm_t: outputs of jabd_assign;
 *   losses[3] = loss_l, loss_c, loss_landm; norms[2] = N, N1 (:293, :299); sel_mask [B,P] u8: bit0 pos (conf_t != 0),
 *   bit1 pos1 (conf_t > 0), bit2 mined negative -- kept for the backward pass.
 * _backward writes d(sum_k grad_losses[k] * losses[k]) / d(loc_data, conf_data, landm_data) (dense, zeros elsewhere);
 * grad_losses[3] is device memory. */
JABD_API size_t jabd_multibox_loss_workspace_bytes(int B);
JABD_API int jabd_multibox_loss_forward(const float *loc_data, const float *conf_data, const float *landm_data,
                                        const float *loc_t, const int64_t *conf_t, const float *landm_t, int B, int64_t P,
                                        int negpos_ratio, float *losses, float *norms, unsigned char *sel_mask,
                                        void *workspace, size_t workspace_bytes, jabd_stream_t stream);
JABD_API int jabd_multibox_loss_backward(const float *loc_data, const float *conf_data, const float *landm_data,
                                         const float *loc_t, const float *landm_t, const unsigned char *sel_mask,
                                         const float *norms, const float *grad_losses, int B, int64_t P, float *g_loc,
                                         float *g_conf, float *g_landm, jabd_stream_t stream);

/* ---- IoU-family overlaps, IouLoss and the DIoU variant of the MultiBox loss.
 * kind: 1 = IoU, 2 = GIoU, 3 = DIoU, 4 = CIoU.
 * jabd_bbox_overlaps_family: bbox_overlaps_{iou,giou,diou,ciou}(bboxes1, bboxes2) of R/utils/box_utils.py:5-158 -- the
 *   reference pairs row i of bboxes1 with row i of bboxes2 -> out [N] (clamped like the reference).
 * jabd_iou_loss_forward / _backward: IouLoss.forward (R/nets/retinaface_training_DIOU.py:491-525) and its gradient with
 *   respect to loc_p.  priors != NULL: pred_mode 'Center' (loc_p is decoded against priors [N,4] first), NULL: loc_p are
 *   boxes.  size_sum != 0: sum, else mean.  per_row [N] (optional) receives 1 - overlap; loss [1], grad_loss [1] device.
 * jabd_multibox_loss_forward_ex / _backward_ex: jabd_multibox_loss_* with the box term selected by loc_loss: 0 =
 *   smooth-L1 on encoded targets, 1..4 = IouLoss(kind) on decode(loc_data, priors [P,4]) against raw matched boxes in
 *   loc_t (jabd_assign with encode_mode 0, i.e. match_iou) -- MultiBoxLoss of R/nets/retinaface_training_DIOU.py:527-665. */
JABD_API int jabd_bbox_overlaps_family(const float *boxes1, const float *boxes2, int64_t N, int kind, float *out,
                                       jabd_stream_t stream);
JABD_API int jabd_iou_loss_forward(const float *loc_p, const float *loc_t, const float *priors, int64_t N, float var0,
                                   float var1, int kind, int size_sum, float *per_row, float *loss, jabd_stream_t stream);
JABD_API int jabd_iou_loss_backward(const float *loc_p, const float *loc_t, const float *priors, int64_t N, float var0,
                                    float var1, int kind, int size_sum, const float *grad_loss, float *g_loc,
                                    jabd_stream_t stream);
JABD_API int jabd_multibox_loss_forward_ex(const float *loc_data, const float *conf_data, const float *landm_data,
                                           const float *loc_t, const int64_t *conf_t, const float *landm_t, int B, int64_t P,
                                           int negpos_ratio, int loc_loss, const float *priors, float var0, float var1,
                                           float *losses, float *norms, unsigned char *sel_mask, void *workspace,
                                           size_t workspace_bytes, jabd_stream_t stream);
JABD_API int jabd_multibox_loss_backward_ex(const float *loc_data, const float *conf_data, const float *landm_data,
                                            const float *loc_t, const float *landm_t, const unsigned char *sel_mask,
                                            const float *norms, const float *grad_losses, int B, int64_t P, int loc_loss,
                                            const float *priors, float var0, float var1, float *g_loc, float *g_conf,
                                            float *g_landm, jabd_stream_t stream);

/* ---- WIDER-FACE AP evaluation (R/utils/utils_map.py:75-223), fp64 like the reference.
 * Rows: pred [sumN,5] = x y w h score per detection in file order (read_pred_file, :45-58), gt [sumG,4] = x y w h,
 * packed per image with pred_off / gt_off [I+1]; keep [sumG] u8 = 1 for the faces listed in the subset's gt_list
 * (`ignore[keep_index-1] = 1`, :198-200).
 * jabd_norm_score: in-place min-max normalisation of all scores (norm_score, :75-98; min starts at 1, max at 0).
 * jabd_wider_eval: per image image_eval (:100-132) + img_pr_info (:135-148), added into pr_curve [thresh_num,2]
 *   (the `pr_curve += _img_pr_info` of evaluation, :203); images with no predictions or no GT are skipped (:196-197).
 *   Optional per-prediction outputs of image_eval: pred_recall [sumN] i32, proposal_list [sumN] i32 (+1 / -1); rows of
 *   skipped images are left untouched (pred_recall) / set to 1 (proposal_list).
 * dataset_pr_info and voc_ap (:151-170) are 1000-element host arithmetic and stay in the Python drop-in. */
/* bbox_overlaps (:16-27) on point-form fp64 boxes [A,4] x [B,4] -> out [A,B]; img_pr_info (:135-148) for one image from
 * image_eval's outputs: pred [N,5], proposal_list [N] i32 (+1/-1), pred_recall [N] i32 -> pr_info [thresh_num,2] (overwritten);
 * workspace: 256-byte aligned, >= round_up(4*thresh_num,256) + round_up(4*N,256) bytes. */
JABD_API int jabd_bbox_overlaps_f64(const double *box_a, int64_t A, const double *box_b, int64_t B, double *out,
                                    jabd_stream_t stream);
JABD_API int jabd_img_pr_info(const double *pred, int64_t N, const int *proposal_list, const int *pred_recall,
                              int thresh_num, double *pr_info, void *workspace, size_t workspace_bytes, jabd_stream_t stream);
JABD_API int jabd_norm_score(double *pred, int64_t sumN, void *workspace, size_t workspace_bytes, jabd_stream_t stream);
JABD_API size_t jabd_wider_eval_workspace_bytes(int I, int64_t sumN, int64_t sumG, int thresh_num);
JABD_API int jabd_wider_eval(const double *pred, const int *pred_off, const double *gt, const int *gt_off,
                             const unsigned char *keep, int I, int64_t sumN, int64_t sumG, double iou_thresh,
                             int thresh_num, double *pr_curve, int *pred_recall, int *proposal_list, void *workspace,
                             size_t workspace_bytes, jabd_stream_t stream);

/* ---- S1/K1/N1/N2: score threshold, top-k, greedy NMS ---------------------------------------------- */
/* thresh_mode: 0 = none, 1 = score >= conf_thres (R/utils/utils_bbox.py:266), 2 = score > conf_thres. */
/* nms_mode: 0 = torchvision.ops.nms semantics (call site R/utils/utils_bbox.py:275-279): stable descending
 *               order, suppress iff (double)(inter/((area_i+area_j)-inter)) > nms_thres;
 *           1 = SSD-legacy nms / nms_r (R/utils/box_utils.py:384-448, R/utils/utils_bbox.py:116-180): ascending
 *               order picked from the end (ties: higher index first), union = (area_j-inter)+area_i,
 *               survive iff IoU <= (float)nms_thres.
 *           | JABD_NMS_EXACT_DIV: evaluate inter/union for every pair.  By default a pair whose IoU is more
 *               than 2^-20 (relative) away from the threshold is decided by comparing inter with thr*union,
 *               which provably gives the same decision (detect.cu, suppresses()); tests compare the two. */
#define JABD_NMS_EXACT_DIV 256
/* jabd_nms / jabd_diounms / jabd_detect run each image (segment) on a thread-block cluster of 1 to 8 CTAs -- one SM
 * each -- picked per call as the widest cluster for which the whole batch is still co-resident.  A call may pin the width
 * (tests, measurements; results do not depend on it): `| JABD_NMS_CLUSTER(c)` in nms_mode, `JABD_DET_CLUSTER(c)` in the
 * flags of jabd_detect*; c = 0 (automatic, the default) or 1..8 (any count, not only powers of two). */
#define JABD_NMS_CLUSTER(c) ((c) << 12)
#define JABD_DET_CLUSTER(c) (c)
/* Every jabd_nms / jabd_diounms / jabd_detect call leaves JABD_SEL_STATS ints per segment in its workspace at byte offset
 * jabd_nms_stats_offset(S, keep_cap): [0] selection rounds (<= 6144 candidates each), [1] rounds that ran the exact
 * three-pass radix select + ordered compaction (more than 8192 candidates at or above the cut bin, i.e. massive score ties),
 * [2] candidates handed to NMS, [3] 32-wide NMS chunks processed.  jabd_topk writes the same row per segment to the start of
 * its (optional) workspace.  Diagnostics only. */
#define JABD_SEL_STATS 4

/* Segmented top-k (K1; stable descending order, ties -> lower index).  scores[s*seg_stride + i*elem_stride],
 * i < N, for s < S segments; out_idx [S,K] i32 (padding -1), out_count [S].  workspace: NULL, or
 * jabd_topk_workspace_bytes() to receive the selection statistics (JABD_SEL_STATS ints per segment). */
JABD_API size_t jabd_topk_workspace_bytes(int S, int64_t N, int K);
JABD_API int jabd_topk(const float *scores, int64_t seg_stride, int64_t elem_stride, int S, int64_t N, float conf_thres,
                       int thresh_mode, int K, int *out_idx, int *out_count, void *workspace, size_t workspace_bytes,
                       jabd_stream_t stream);

/* Greedy NMS over S segments of N pre-decoded boxes.  boxes[s*box_seg_stride + i*box_stride + 0..3] x1y1x2y2,
 * scores[s*score_seg_stride + i*score_stride].  pre_nms_topk <= 0: uncapped.  keep_idx [S,keep_cap] i32 in
 * the reference's output order (padding -1), keep_count [S].  Only the first keep_cap keeps are produced. */
JABD_API size_t jabd_nms_workspace_bytes(int S, int64_t N, int keep_cap);
JABD_API size_t jabd_nms_stats_offset(int S, int keep_cap);
JABD_API int jabd_nms(const float *boxes, int64_t box_seg_stride, int64_t box_stride, const float *scores,
                      int64_t score_seg_stride, int64_t score_stride, int S, int64_t N, float conf_thres, int thresh_mode,
                      int pre_nms_topk, double nms_thres, int nms_mode, int keep_cap, int *keep_idx, int *keep_count,
                      void *workspace, size_t workspace_bytes, jabd_stream_t stream);

/* diounms (R/utils/utils_bbox.py:182-258): the SSD-legacy loop (ascending order picked from the end, top_k = pre_nms_topk,
 * union = (area_j-inter)+area_i) with the criterion IoU - (centre distance^2 / enclosing diagonal^2)^beta1 <= overlap. */
JABD_API int jabd_diounms(const float *boxes, int64_t box_seg_stride, int64_t box_stride, const float *scores,
                          int64_t score_seg_stride, int64_t score_stride, int S, int64_t N, int pre_nms_topk, double overlap,
                          float beta1, int keep_cap, int *keep_idx, int *keep_count, void *workspace, size_t workspace_bytes,
                          jabd_stream_t stream);

/* Fused inference post-processing for a batch (R/predict.py:167-181 composed per SURVEY D4):
 * class-1 score (conf[:,1]) -> threshold -> top-k -> decode of the candidates -> NMS -> first keep_cap rows
 * [x1 y1 x2 y2 score | decode_landm] (zero padded), prior indices (padding -1) and counts.
 * loc [B,P,4], conf [B,P,2], landm [B,P,10] (NULL: landmark columns are zero), priors [P,4]. */
JABD_API size_t jabd_detect_workspace_bytes(int B, int64_t P, int keep_cap);
JABD_API int jabd_detect(const float *loc, const float *conf, const float *landm, const float *priors, int B, int64_t P,
                         float var0, float var1, float conf_thres, int thresh_mode, int pre_nms_topk, double nms_thres,
                         int keep_cap, int flags, float *dets, int *counts, int *keep_idx, void *workspace,
                         size_t workspace_bytes, jabd_stream_t stream);
/* Same with HOST buffers: loc/conf/landm in, dets/counts/keep_idx out; priors stay on the device.
 * Synchronises `stream` before returning.  Pinned (page-locked, hence device-mapped) loc_host / landm_host are not
 * uploaded: the kernel reads the loc rows of the <= pre_nms_topk candidates and the landmark rows of the <= keep_cap kept
 * detections directly from host memory (conf is always copied: every score is scanned); pageable memory is copied. */
/* Several independent batches in one call.  n_lanes == 0 (recommended): ONE launch on `stream` whose grid covers the images of
 * all batches (16 batches per launch; each cluster finds its batch in a table carried by the kernel parameters), so the block
 * scheduler hands the next image to whichever SM falls free.  n_lanes > 0: batch i is its own launch on lanes[i % n_lanes]
 * (the lane protocol of jabd_assign_batches: caller-owned side streams, forked from and joined into `stream` by events).
 * Either way the automatic cluster width is chosen for all the images that are in flight together, i.e. narrower than a lone
 * call's -- one SM per image does the least redundant work, and the other images keep the remaining SMs busy.  Every batch
 * needs its own outputs and workspace (two batches on the same lane may share); `priors`, thresholds and `keep_cap` are
 * shared.  Rows, counts and keep lists are those of jabd_detect.  Capturable. */
typedef struct {
    const float *loc;     /* [B,P,4] */
    const float *conf;    /* [B,P,2] */
    const float *landm;   /* [B,P,10] or NULL */
    int B;
    float *dets;          /* [B,keep_cap,15] */
    int *counts;          /* [B] */
    int *keep_idx;        /* [B,keep_cap] */
    void *workspace;      /* jabd_detect_workspace_bytes(B, P, keep_cap) */
    size_t workspace_bytes;
} jabd_detect_batch_t;
JABD_API int jabd_detect_batches(const float *priors, int64_t P, const jabd_detect_batch_t *batches_host, int n_batches, float var0,
                                 float var1, float conf_thres, int thresh_mode, int pre_nms_topk, double nms_thres, int keep_cap,
                                 int flags, const jabd_stream_t *lanes_host, int n_lanes, jabd_stream_t stream);
JABD_API size_t jabd_detect_host_scratch_bytes(int B, int64_t P, int keep_cap, int with_landm);
JABD_API int jabd_detect_host(const float *loc_host, const float *conf_host, const float *landm_host,
                              const float *priors_dev, int B, int64_t P, float var0, float var1, float conf_thres,
                              int thresh_mode, int pre_nms_topk, double nms_thres, int keep_cap, int flags,
                              float *dets_host, int *counts_host, int *keep_idx_host, void *dev_scratch,
                              size_t dev_scratch_bytes, jabd_stream_t stream);
/* Same without the final synchronisation: host buffers must be pinned; the caller waits on `stream` (or an event
 * recorded on it) before reading the outputs.  Two calls on two streams with two scratch areas overlap. */
JABD_API int jabd_detect_host_async(const float *loc_host, const float *conf_host, const float *landm_host,
                              const float *priors_dev, int B, int64_t P, float var0, float var1, float conf_thres,
                              int thresh_mode, int pre_nms_topk, double nms_thres, int keep_cap, int flags,
                              float *dets_host, int *counts_host, int *keep_idx_host, void *dev_scratch,
                              size_t dev_scratch_bytes, jabd_stream_t stream);

/* ---- peer-memory exchange of the validation flow over NVLink / NVSwitch (csrc/p2p.cu) ---------------------------------
 * Replaces the reference's pickle -> pad -> two all_gather pattern for evaluation results (R/utils.py:62-90) and the NCCL
 * all-gather of sharding.DetectionGather: every rank copies its block of padded detections straight into slot `rank` of every
 * rank's receive buffer (16-byte peer stores, one NVSwitch hop) and raises a flag there; consumers wait on their own flags.
 * One process per GPU: buffers are exported / imported as CUDA IPC handles, which the caller ships over its host channel.
 * These four calls are the one place the library allocates device memory, explicitly and at the caller's request: IPC handles
 * need an allocation of their own (a framework's caching allocator hands out interior pointers).  alloc zero-fills and
 * synchronises the device once; open / close / free are the matching calls on the peers / the owner. */
#define JABD_P2P_HANDLE_BYTES 64
#define JABD_P2P_MAX_PEERS 16
JABD_API int jabd_p2p_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64);
JABD_API int jabd_p2p_open(const unsigned char *handle64, void **dev_ptr);
JABD_API int jabd_p2p_close(void *dev_ptr);
JABD_API int jabd_p2p_free(void *dev_ptr);
/* ONE kernel: copies `bytes` from src (device, 16-byte aligned) to peer_bufs[j] + dst_offset for every j < n_ranks (HOST arrays
 * of device pointers as mapped in THIS process; entry `rank` is the own buffer) and then stores `seq` into peer_flags[j][rank]
 * with system-scope release.  counters: n_ranks zeroed unsigned ints in device memory, private to the stream (left zero again).
 * seq must increase from one exchange to the next on the same flags.  bytes == 0 (src may be null) only raises the flags.
 * Handshake, ack_seq != 0: before anything is stored into peer j's buffer the kernel (a) stores ack_seq into peer_acks[j][rank]
 * -- "this rank has read what j sent into this slot last time"; the call must therefore be stream-ordered behind those reads --
 * and (b) waits until own_acks[j] >= ack_seq, i.e. until j has said the same (time-out as in jabd_p2p_wait, *status = 101 + j). */
JABD_API int jabd_p2p_allgather(const void *src, size_t bytes, void *const *peer_bufs, size_t dst_offset,
                                unsigned long long *const *peer_flags, unsigned long long *const *peer_acks,
                                const unsigned long long *own_acks, int n_ranks, int rank, unsigned long long seq,
                                unsigned long long ack_seq, unsigned int *counters, double timeout_s, int *status,
                                jabd_stream_t stream);
/* Enqueues a one-CTA kernel that returns once flags[j] >= seq for every j < n_ranks (system-scope acquire): work enqueued after
 * it on `stream` sees all n_ranks blocks.  After timeout_s (<= 0: 2 s) of spinning it gives up and stores 1 + j into *status
 * (device int, may be null) instead of hanging the device. */
JABD_API int jabd_p2p_wait(const unsigned long long *flags, int n_ranks, unsigned long long seq, double timeout_s, int *status,
                           jabd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* JABD_B200_H */
