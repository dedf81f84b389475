/*
 * jabd_oracle.c -- CPU restatement of the JABD box-geometry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker*: it may be imported,
 * linked or executed only from tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg.  Nothing under the product package
 * calls it and there is no CPU fallback in the product.
 *
 * Every function restates, in scalar fp32 C with the reference's operation
 * order, one function of the reference tree R/ = JABD2080ti/ (Python/PyTorch,
 * eager CPU semantics).  Citations are R/file:line.
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or fixtures
 * (SURVEY.md section 4), so this oracle is pinned against outputs of the
 * reference itself, imported and run in the dev container by
 * tests/golden/make_golden.py (torch 2.11.0 CPU, torchvision 0.26.0 CPU) and
 * committed under tests/golden/.  tests/test_oracle_golden.py holds the pins.
 * NMS arithmetic lives in a third-party dependency that is not vendored and
 * not version-pinned by the reference (torchvision.ops.nms, call site
 * R/utils/utils_bbox.py:275-279); the oracle of record is the torchvision
 * 0.26.0 CPU kernel whose published algorithm is restated in orc_nms_tv().
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: no FMA contraction,
 * so every +,-,*,/ rounds once exactly like eager torch CPU).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ priors */

/* R/utils/anchors.py:9-42 (Anchors.__init__/get_anchors; Anchors_eval :43-79
 * is the same loop).  Level-major, row i, col j, min_size; Python float64
 * arithmetic ((j+0.5)*step/W), rounded to fp32 by torch.Tensor(list); optional
 * clamp to [0,1] (:39-40).  sizes_off has n_levels+1 entries into min_sizes. */
ORC_API int64_t orc_priors(const int *steps, const double *min_sizes, const int *sizes_off,
                           int n_levels, int H, int W, int clip, float *out)
{
    int64_t n = 0;
    for (int k = 0; k < n_levels; ++k) {
        int step = steps[k];
        int fh = (int)ceil((double)H / (double)step); /* :21 */
        int fw = (int)ceil((double)W / (double)step);
        for (int i = 0; i < fh; ++i)
            for (int j = 0; j < fw; ++j)
                for (int s = sizes_off[k]; s < sizes_off[k + 1]; ++s) {
                    double ms = min_sizes[s];
                    double v[4];
                    v[0] = ((double)j + 0.5) * (double)step / (double)W; /* :33 */
                    v[1] = ((double)i + 0.5) * (double)step / (double)H; /* :34 */
                    v[2] = ms / (double)W;                               /* :31 */
                    v[3] = ms / (double)H;                               /* :32 */
                    if (out) {
                        for (int c = 0; c < 4; ++c) {
                            float f = (float)v[c];
                            if (clip) { if (f > 1.0f) f = 1.0f; if (f < 0.0f) f = 0.0f; }
                            out[4 * n + c] = f;
                        }
                    }
                    ++n;
                }
    }
    return n;
}

/* ---------------------------------------------------------------- geometry */

/* R/nets/retinaface_training.py:8-10 == R/utils/box_utils.py:160-170 */
ORC_API void orc_point_form(const float *boxes, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float *b = boxes + 4 * i;
        float hw = b[2] / 2.0f, hh = b[3] / 2.0f;
        out[4 * i + 0] = b[0] - hw;
        out[4 * i + 1] = b[1] - hh;
        out[4 * i + 2] = b[0] + hw;
        out[4 * i + 3] = b[1] + hh;
    }
}

/* One IoU exactly as intersect()+jaccard() evaluate it:
 * R/nets/retinaface_training.py:22-59 (== R/utils/box_utils.py:185-226).
 * inter = clamp(min(a2,b2)-max(a1,b1), 0).x * (...).y ; union = (area_a +
 * area_b) - inter ; iou = inter / union.  No epsilon, no clamp on union. */
static inline float iou_pair(const float *a, float area_a, const float *b, float area_b)
{
    float mx = (a[2] < b[2] ? a[2] : b[2]) - (a[0] > b[0] ? a[0] : b[0]);
    float my = (a[3] < b[3] ? a[3] : b[3]) - (a[1] > b[1] ? a[1] : b[1]);
    if (mx < 0.0f) mx = 0.0f;
    if (my < 0.0f) my = 0.0f;
    float inter = mx * my;
    float uni = (area_a + area_b) - inter;
    return inter / uni;
}

/* jaccard(box_a[A,4], box_b[B,4]) -> [A,B]; both in point form. */
ORC_API void orc_jaccard(const float *a, int64_t A, const float *b, int64_t B, float *out)
{
    float *area_b = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    for (int64_t j = 0; j < B; ++j)
        area_b[j] = (b[4 * j + 2] - b[4 * j + 0]) * (b[4 * j + 3] - b[4 * j + 1]);
    for (int64_t i = 0; i < A; ++i) {
        float area_a = (a[4 * i + 2] - a[4 * i + 0]) * (a[4 * i + 3] - a[4 * i + 1]);
        for (int64_t j = 0; j < B; ++j)
            out[i * B + j] = iou_pair(a + 4 * i, area_a, b + 4 * j, area_b[j]);
    }
    free(area_b);
}

/* R/nets/retinaface_training.py:61-70 (== R/utils/box_utils.py:323-344).
 * var0*wh is formed first in fp32 ((float)var0 * w), then an IEEE divide. */
static inline void encode_one(const float *m, const float *p, float var0, float var1, float *o)
{
    o[0] = ((m[0] + m[2]) / 2.0f - p[0]) / (var0 * p[2]);
    o[1] = ((m[1] + m[3]) / 2.0f - p[1]) / (var0 * p[3]);
    o[2] = logf((m[2] - m[0]) / p[2]) / var1;
    o[3] = logf((m[3] - m[1]) / p[3]) / var1;
}

ORC_API void orc_encode(const float *matched, const float *priors, int64_t n, float var0, float var1,
                        float *out)
{
    for (int64_t i = 0; i < n; ++i) encode_one(matched + 4 * i, priors + 4 * i, var0, var1, out + 4 * i);
}

/* R/nets/retinaface_training.py:72-84 */
static inline void encode_landm_one(const float *lm, const float *p, float var0, float *o)
{
    float dx = var0 * p[2], dy = var0 * p[3];
    for (int k = 0; k < 5; ++k) {
        o[2 * k + 0] = (lm[2 * k + 0] - p[0]) / dx;
        o[2 * k + 1] = (lm[2 * k + 1] - p[1]) / dy;
    }
}

ORC_API void orc_encode_landm(const float *matched, const float *priors, int64_t n, float var0, float *out)
{
    for (int64_t i = 0; i < n; ++i) encode_landm_one(matched + 10 * i, priors + 4 * i, var0, out + 10 * i);
}

/* ------------------------------------------------------------------- match */

/*
 * match() for one image.
 *   live 10-arg form   R/nets/retinaface_training.py:93-162   (label_mode 0)
 *   SSD 8-arg form     R/utils/box_utils.py:276-320           (label_mode 1: conf = label + 1)
 *   match_iou/_ious    R/nets/retinaface_training_DIOU.py:176-246,
 *                      R/utils/box_utils.py:229-273           (encode_mode 0: loc_t = matched box)
 * Steps, with the reference's semantics:
 *   overlaps = jaccard(truths, point_form(priors))                  :100-103
 *   best prior per GT  = overlaps.max(1)  -> first index on ties    :111
 *   best GT per prior  = overlaps.max(0)  -> first index on ties    :120
 *   best_truth_overlap[best_prior_idx] = 2                          :127
 *   for j in 0..G-1: best_truth_idx[best_prior_idx[j]] = j          :129-130 (last j wins)
 *   conf = labels[best_truth_idx]; conf[overlap < (float)thr] = 0   :137,143
 *   loc = encode(...), landm = encode_landm(...)                    :148-149
 * torch.max propagates NaN (a NaN wins, first NaN index); restated for
 * completeness although valid GT never produce one.
 * labels are float (+1/-1); conf_t is int64 (torch.LongTensor, :199), the
 * float->int64 store truncates.  landms / landm_t may be NULL (8-arg forms).
 * Returns 0, or -1 if G == 0 (torch raises on max over an empty dim).
 */
ORC_API int orc_match(float threshold, const float *truths, const float *priors, float var0, float var1,
                      const float *labels, const float *landms, int64_t G, int64_t P,
                      int label_mode, int encode_mode,
                      float *loc_t, int64_t *conf_t, float *landm_t,
                      int64_t *best_truth_idx, float *best_truth_overlap,
                      int64_t *best_prior_idx, float *best_prior_overlap)
{
    if (G <= 0) return -1;
    float *pf = (float *)malloc(sizeof(float) * 4 * (size_t)P);
    float *area_b = (float *)malloc(sizeof(float) * (size_t)P);
    int64_t *bti = best_truth_idx ? best_truth_idx : (int64_t *)malloc(sizeof(int64_t) * (size_t)P);
    float *bto = best_truth_overlap ? best_truth_overlap : (float *)malloc(sizeof(float) * (size_t)P);
    int64_t *bpi = best_prior_idx ? best_prior_idx : (int64_t *)malloc(sizeof(int64_t) * (size_t)G);
    float *bpo = best_prior_overlap ? best_prior_overlap : (float *)malloc(sizeof(float) * (size_t)G);
    orc_point_form(priors, P, pf);
    for (int64_t p = 0; p < P; ++p)
        area_b[p] = (pf[4 * p + 2] - pf[4 * p + 0]) * (pf[4 * p + 3] - pf[4 * p + 1]);
    for (int64_t g = 0; g < G; ++g) {
        const float *a = truths + 4 * g;
        float area_a = (a[2] - a[0]) * (a[3] - a[1]);
        float rbest = 0.0f; int64_t ridx = 0; int rnan = 0;
        for (int64_t p = 0; p < P; ++p) {
            float v = iou_pair(a, area_a, pf + 4 * p, area_b[p]);
            /* row max (dim 1): first max, NaN wins */
            if (p == 0) { rbest = v; ridx = 0; rnan = isnan(v); }
            else if (!rnan && (isnan(v) || v > rbest)) { rbest = v; ridx = p; rnan = isnan(v); }
            /* column max (dim 0) */
            if (g == 0) { bto[p] = v; bti[p] = 0; }
            else if (!isnan(bto[p]) && (isnan(v) || v > bto[p])) { bto[p] = v; bti[p] = g; }
        }
        bpi[g] = ridx; bpo[g] = rbest;
    }
    for (int64_t g = 0; g < G; ++g) bto[bpi[g]] = 2.0f;           /* :127 */
    for (int64_t g = 0; g < G; ++g) bti[bpi[g]] = g;              /* :129-130 */
    for (int64_t p = 0; p < P; ++p) {
        int64_t t = bti[p];
        const float *m = truths + 4 * t;
        float c = labels[t] + (label_mode ? 1.0f : 0.0f);        /* box_utils.py:315 */
        if (bto[p] < threshold) c = 0.0f;                         /* :143 */
        conf_t[p] = (int64_t)c;
        if (encode_mode) encode_one(m, priors + 4 * p, var0, var1, loc_t + 4 * p);
        else memcpy(loc_t + 4 * p, m, 4 * sizeof(float));        /* DIOU.py:229 */
        if (landm_t && landms) encode_landm_one(landms + 10 * t, priors + 4 * p, var0, landm_t + 10 * p);
    }
    free(pf); free(area_b);
    if (!best_truth_idx) free(bti);
    if (!best_truth_overlap) free(bto);
    if (!best_prior_idx) free(bpi);
    if (!best_prior_overlap) free(bpo);
    return 0;
}

/* ------------------------------------------------------------------ decode */

/* R/utils/utils_bbox.py:29-34 (== box_utils.py:348-367):
 * cxcy = p.cxcy + (loc.xy*var0)*p.wh ; wh = p.wh*exp(loc.wh*var1) ;
 * x1y1 = cxcy - wh/2 ; x2y2 = wh + x1y1. */
ORC_API void orc_decode(const float *loc, const float *priors, int64_t n, float var0, float var1, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float *l = loc + 4 * i, *p = priors + 4 * i;
        float cx = p[0] + (l[0] * var0) * p[2];
        float cy = p[1] + (l[1] * var0) * p[3];
        float w = p[2] * expf(l[2] * var1);
        float h = p[3] * expf(l[3] * var1);
        float x1 = cx - w / 2.0f, y1 = cy - h / 2.0f;
        out[4 * i + 0] = x1;
        out[4 * i + 1] = y1;
        out[4 * i + 2] = w + x1;
        out[4 * i + 3] = h + y1;
    }
}

/* R/utils/utils_bbox.py:39-46 */
ORC_API void orc_decode_landm(const float *pre, const float *priors, int64_t n, float var0, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float *l = pre + 10 * i, *p = priors + 4 * i;
        for (int k = 0; k < 5; ++k) {
            out[10 * i + 2 * k + 0] = p[0] + (l[2 * k + 0] * var0) * p[2];
            out[10 * i + 2 * k + 1] = p[1] + (l[2 * k + 1] * var0) * p[3];
        }
    }
}

/* --------------------------------------------------------------------- NMS */

typedef struct { float s; int64_t i; } orc_si;

static int cmp_desc_stable(const void *a, const void *b)
{
    const orc_si *x = (const orc_si *)a, *y = (const orc_si *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i); /* ties: lower index first */
}

/* Stable descending argsort: scores.sort(stable=True, descending=True). */
ORC_API void orc_argsort_desc(const float *scores, int64_t n, int64_t *order)
{
    orc_si *v = (orc_si *)malloc(sizeof(orc_si) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { v[i].s = scores[i]; v[i].i = i; }
    qsort(v, (size_t)n, sizeof(orc_si), cmp_desc_stable);
    for (int64_t i = 0; i < n; ++i) order[i] = v[i].i;
    free(v);
}

/* torchvision.ops.nms, CPU kernel of torchvision 0.26.0 (called from
 * R/utils/utils_bbox.py:275-279).  Stable descending sort; greedy; for a kept
 * box i every later, still-alive j is suppressed iff
 *   inter/((area_i+area_j)-inter) > iou_threshold      (float promoted to double)
 * with inter = max(0,xx2-xx1)*max(0,yy2-yy1).  Keeps come out in descending
 * score order as int64.  Returns the number kept. */
ORC_API int64_t orc_nms_tv(const float *boxes, const float *scores, int64_t n, double iou_threshold,
                           int64_t *keep)
{
    if (n <= 0) return 0;
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    unsigned char *sup = (unsigned char *)calloc((size_t)n, 1);
    orc_argsort_desc(scores, n, order);
    for (int64_t i = 0; i < n; ++i)
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    int64_t nk = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = order[_i];
        if (sup[i]) continue;
        keep[nk++] = i;
        const float *bi = boxes + 4 * i;
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j];
            if (sup[j]) continue;
            const float *bj = boxes + 4 * j;
            float xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
            float yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
            float xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
            float yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / ((area[i] + area[j]) - inter);
            if ((double)ovr > iou_threshold) sup[j] = 1;
        }
    }
    free(order); free(area); free(sup);
    return nk;
}

/* SSD-legacy greedy NMS: R/utils/box_utils.py:384-448 == nms_r
 * R/utils/utils_bbox.py:116-180.  Ascending sort, candidates = last top_k,
 * pick from the end; union = (area_j - inter) + area_i (note the association);
 * a candidate survives iff IoU <= (float)overlap.  `order_asc` is the
 * ascending argsort the caller obtained (torch's scores.sort(0) is not stable,
 * so the tie order is an input, not restated).  keep has n entries
 * (zero-filled), returns count. */
ORC_API int64_t orc_nms_ssd(const float *boxes, const int64_t *order_asc, int64_t n, float overlap,
                            int64_t top_k, int64_t *keep)
{
    for (int64_t i = 0; i < n; ++i) keep[i] = 0;
    if (n <= 0) return 0;
    int64_t m = n < top_k ? n : top_k;
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)m);
    memcpy(idx, order_asc + (n - m), sizeof(int64_t) * (size_t)m);
    int64_t count = 0;
    while (m > 0) {
        int64_t i = idx[m - 1];
        keep[count++] = i;
        if (m == 1) break;
        --m;
        const float *bi = boxes + 4 * i;
        float area_i = (bi[2] - bi[0]) * (bi[3] - bi[1]);
        int64_t w_ = 0;
        for (int64_t k = 0; k < m; ++k) {
            int64_t j = idx[k];
            const float *bj = boxes + 4 * j;
            float xx1 = bj[0] < bi[0] ? bi[0] : bj[0];   /* clamp(min=x1[i]) */
            float yy1 = bj[1] < bi[1] ? bi[1] : bj[1];
            float xx2 = bj[2] > bi[2] ? bi[2] : bj[2];   /* clamp(max=x2[i]) */
            float yy2 = bj[3] > bi[3] ? bi[3] : bj[3];
            float w = xx2 - xx1, h = yy2 - yy1;
            if (w < 0.0f) w = 0.0f;
            if (h < 0.0f) h = 0.0f;
            float inter = w * h;
            float area_j = (bj[2] - bj[0]) * (bj[3] - bj[1]);
            float uni = (area_j - inter) + area_i;
            float iou = inter / uni;
            if (iou <= overlap) idx[w_++] = j;
        }
        m = w_;
    }
    free(idx);
    return count;
}

/* diounms, R/utils/utils_bbox.py:182-258: the loop of orc_nms_ssd with the criterion
 * inter/union - (d/c)^beta1 <= overlap; d = squared centre distance, c = squared diagonal of the enclosing box.
 * torch.pow(u, 1.0) returns u itself. */
ORC_API int64_t orc_diounms(const float *boxes, const int64_t *order_asc, int64_t n, float overlap, int64_t top_k,
                            float beta1, int64_t *keep)
{
    for (int64_t i = 0; i < n; ++i) keep[i] = 0;
    if (n <= 0) return 0;
    int64_t m = n < top_k ? n : top_k;
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)m);
    memcpy(idx, order_asc + (n - m), sizeof(int64_t) * (size_t)m);
    int64_t count = 0;
    while (m > 0) {
        int64_t i = idx[m - 1];
        keep[count++] = i;
        if (m == 1) break;
        --m;
        const float *bi = boxes + 4 * i;
        float area_i = (bi[2] - bi[0]) * (bi[3] - bi[1]);
        float cxi = (bi[0] + bi[2]) / 2, cyi = (bi[1] + bi[3]) / 2;
        int64_t w_ = 0;
        for (int64_t k = 0; k < m; ++k) {
            int64_t j = idx[k];
            const float *bj = boxes + 4 * j;
            float inx1 = bj[0] < bi[0] ? bi[0] : bj[0], iny1 = bj[1] < bi[1] ? bi[1] : bj[1];
            float inx2 = bj[2] > bi[2] ? bi[2] : bj[2], iny2 = bj[3] > bi[3] ? bi[3] : bj[3];
            float cxj = (bj[0] + bj[2]) / 2, cyj = (bj[1] + bj[3]) / 2;
            float ddx = cxi - cxj, ddy = cyi - cyj;
            float d = ddx * ddx + ddy * ddy;
            float ox1 = bj[0] > bi[0] ? bi[0] : bj[0], oy1 = bj[1] > bi[1] ? bi[1] : bj[1];
            float ox2 = bj[2] < bi[2] ? bi[2] : bj[2], oy2 = bj[3] < bi[3] ? bi[3] : bj[3];
            float ex = ox2 - ox1, ey = oy2 - oy1;
            float c = ex * ex + ey * ey;
            float u = d / c;
            float w = inx2 - inx1, h = iny2 - iny1;
            if (w < 0.0f) w = 0.0f;
            if (h < 0.0f) h = 0.0f;
            float inter = w * h;
            float area_j = (bj[2] - bj[0]) * (bj[3] - bj[1]);
            float uni = (area_j - inter) + area_i;
            float pen = beta1 == 1.0f ? u : powf(u, beta1);
            float metric = inter / uni - pen;
            if (metric <= overlap) idx[w_++] = j;
        }
        m = w_;
    }
    free(idx);
    return count;
}

/* ------------------------------------------------------- composed pipeline */

/*
 * Inference pipeline for one image, as a composition of reference functions
 * (SURVEY.md D4): decode (R/utils/utils_bbox.py:29-34) -> class-1 score
 * (R/predict.py:171) -> threshold (`>=` at R/utils/utils_bbox.py:266, or the
 * strict `>` of the cfg3 pipeline) -> stable descending sort, first
 * pre_nms_topk (<=0: uncapped) -> torchvision nms -> first keep_topk (<=0:
 * uncapped) -> rows [box4, score, landm10] (R/predict.py:180, decode_landm
 * R/utils/utils_bbox.py:39-46).
 * boxes_override (optional, [P,4]) replaces the decoded boxes so that the NMS
 * stage can be checked on bit-identical box inputs.
 * Outputs: dets[keep_cap,15], keep_idx[keep_cap] (prior indices); returns count.
 */
ORC_API int64_t orc_detect(const float *loc, const float *conf, const float *landm, const float *priors,
                           int64_t P, float var0, float var1, float conf_thres, int strict,
                           int64_t pre_nms_topk, double nms_thres, int64_t keep_topk,
                           const float *boxes_override, float *dets, int64_t *keep_idx, int64_t keep_cap)
{
    float *boxes = (float *)malloc(sizeof(float) * 4 * (size_t)(P > 0 ? P : 1));
    if (boxes_override) memcpy(boxes, boxes_override, sizeof(float) * 4 * (size_t)P);
    else orc_decode(loc, priors, P, var0, var1, boxes);
    int64_t *sel = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P > 0 ? P : 1));
    float *ssc = (float *)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1));
    int64_t n = 0;
    for (int64_t p = 0; p < P; ++p) {
        float s = conf[2 * p + 1];
        if (strict ? (s > conf_thres) : (s >= conf_thres)) { sel[n] = p; ssc[n] = s; ++n; }
    }
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    orc_argsort_desc(ssc, n, order);
    int64_t m = (pre_nms_topk > 0 && n > pre_nms_topk) ? pre_nms_topk : n;
    float *cb = (float *)malloc(sizeof(float) * 4 * (size_t)(m > 0 ? m : 1));
    float *cs = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
    for (int64_t k = 0; k < m; ++k) {
        memcpy(cb + 4 * k, boxes + 4 * sel[order[k]], 4 * sizeof(float));
        cs[k] = ssc[order[k]];
    }
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
    int64_t nk = orc_nms_tv(cb, cs, m, nms_thres, keep);
    if (keep_topk > 0 && nk > keep_topk) nk = keep_topk;
    if (nk > keep_cap) nk = keep_cap;
    for (int64_t k = 0; k < nk; ++k) {
        int64_t p = sel[order[keep[k]]];
        keep_idx[k] = p;
        memcpy(dets + 15 * k, cb + 4 * keep[k], 4 * sizeof(float));
        dets[15 * k + 4] = cs[keep[k]];
        orc_decode_landm(landm + 10 * p, priors + 4 * p, 1, var0, dets + 15 * k + 5);
    }
    free(boxes); free(sel); free(ssc); free(order); free(cb); free(cs); free(keep);
    return nk;
}

ORC_API int orc_version(void) { return 1; }
