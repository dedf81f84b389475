// wider_eval.cu -- WIDER-FACE AP evaluation of a batch of images.
//
// Replaces, per image, image_eval + img_pr_info of R/utils/utils_map.py:100-148 (and the bbox_overlaps / intersect
// they call, :7-27) and accumulates the dataset PR counters that `evaluation` (:173-214) sums over images.  All
// arithmetic is fp64 like the reference's numpy code (boxes are parsed with float(), :45-58); the file is compiled with
// --fmad=false, so x+w, the intersection, the union and inter/union round exactly like numpy's separate passes.
//
// The reference walks the predictions of an image in order and keeps a per-GT state (recall_list).  That loop is
// restated as data-parallel steps that give the same integers:
//   1. per prediction h: best GT by IoU -- np.argmax's first maximum, a NaN (0/0) wins like np.max -- and whether
//      max_overlap >= iou_thresh (:121-123);
//   2. a GT that is "kept" (ignore[g] == 1, i.e. listed in the subset's gt_list) is recalled by the FIRST prediction
//      that matches it (recall_list[g]: 0 -> 1 once, :127-128): atomicMin over h.  A prediction whose best GT is not
//      kept gets proposal_list[h] = -1 (:124-126);
//   3. pred_recall[h] = number of kept GT recalled by predictions 0..h (:130-131) = inclusive scan of "h is the first
//      match of its GT"; the proposal count of img_pr_info (:143-144) = inclusive scan of proposal_list[h] == 1;
//   4. img_pr_info: for threshold t, r_index is the LAST h with score[h] >= 1 - (t+1)/thresh_num (:138-142).  Each h
//      finds the first t it satisfies (thresholds decrease with t), atomicMax of h per t, then a running maximum
//      over t; pr[t] += (proposal count, pred_recall)[r_index] into the dataset counters.  The counters are integers
//      held in fp64, so the order of the atomic additions cannot change them.
#include "common.cuh"

namespace jabd {

constexpr int kEvalThreads = 256;

struct EvalArgs {
    const double *pred;       // [sumN,5] x y w h score
    const int *pred_off;      // [I+1]
    const double *gt;         // [sumG,4] x y w h
    const int *gt_off;        // [I+1]
    const unsigned char *keep;// [sumG] 1: GT counted in this subset (ignore[g] == 1 in the reference)
    double iou_thresh;
    int thresh_num;
    double *pr_curve;         // [thresh_num,2] += (proposals, recalled)
    int *pred_recall;         // [sumN] optional outputs (image_eval's return values)
    int *proposal;            // [sumN]
    // workspace
    int *best_gt;             // [sumN] best GT of each prediction, -1: no match at the threshold
    int *first_h;             // [sumG]
    int *rmax;                // [I,thresh_num]
};

// intersect() / bbox_overlaps(), R/utils/utils_map.py:7-27, for one pair in point form
__device__ __forceinline__ double overlap_f64(double ax1, double ay1, double ax2, double ay2, double area_a, double bx1,
                                              double by1, double bx2, double by2, double area_b)
{
    const double w = fmax(__dsub_rn(fmin(ax2, bx2), fmax(ax1, bx1)), 0.0);
    const double h = fmax(__dsub_rn(fmin(ay2, by2), fmax(ay1, by1)), 0.0);
    const double inter = __dmul_rn(w, h);
    const double uni = __dsub_rn(__dadd_rn(area_a, area_b), inter);
    return __ddiv_rn(inter, uni);
}

// inclusive scan of two small ints across the CTA; running totals carried by the caller
__device__ __forceinline__ void block_scan2(int a, int b, int &ia, int &ib, int &ta, int &tb, int *s_a, int *s_b)
{
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    int xa = a, xb = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ya = __shfl_up_sync(kFull, xa, o), yb = __shfl_up_sync(kFull, xb, o);
        if (lane >= (unsigned)o) { xa += ya; xb += yb; }
    }
    if (lane == 31) { s_a[warp] = xa; s_b[warp] = xb; }
    __syncthreads();
    int ba = 0, bb = 0;
    ta = tb = 0;
#pragma unroll
    for (int w = 0; w < kEvalThreads / 32; ++w) {
        if (w < (int)warp) { ba += s_a[w]; bb += s_b[w]; }
        ta += s_a[w];
        tb += s_b[w];
    }
    ia = xa + ba;
    ib = xb + bb;
    __syncthreads();
}

// r_index bookkeeping of img_pr_info: rmax[t] = last prediction whose score first satisfies threshold t
__device__ __forceinline__ void eval_rmax(const double *pred, int N, int thresh_num, int *rmax)
{
    for (int h = threadIdx.x; h < N; h += kEvalThreads) {
        const double sc = pred[h * 5 + 4];
        int lo = 0, hi = thresh_num; // first t with score >= 1 - (t+1)/thresh_num (thresholds decrease with t); hi: none
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double th = __dsub_rn(1.0, __ddiv_rn((double)(mid + 1), (double)thresh_num));
            if (sc >= th) hi = mid; else lo = mid + 1;
        }
        if (lo < thresh_num) atomicMax(&rmax[lo], h);
    }
}

// running maximum of rmax over t = r_index of every threshold (:138-142), then pr[t] += (proposals, recalled)[r_index]
__device__ __forceinline__ void eval_thresholds(const int *rmax, int thresh_num, const int *prop_cum, const int *pred_recall,
                                                double *pr, int *s_a, int *s_carry)
{
    const int tid = threadIdx.x;
    if (tid == 0) *s_carry = -1;
    __syncthreads();
    for (int base = 0; base < thresh_num; base += kEvalThreads) {
        const int t = base + tid;
        int v = t < thresh_num ? rmax[t] : -1;
        const unsigned lane = lane_id(), warp = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, v, o);
            if (lane >= (unsigned)o) v = max(v, y);
        }
        if (lane == 31) s_a[warp] = v;
        __syncthreads();
        int pre = *s_carry, all = *s_carry;
        for (int w = 0; w < kEvalThreads / 32; ++w) {
            if (w < (int)warp) pre = max(pre, s_a[w]);
            all = max(all, s_a[w]);
        }
        v = max(v, pre);
        __syncthreads();
        if (tid == 0) *s_carry = all;
        if (t < thresh_num && v >= 0) {
            atomicAdd(&pr[2 * t], (double)prop_cum[v]);
            atomicAdd(&pr[2 * t + 1], (double)pred_recall[v]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kEvalThreads) wider_eval_kernel(EvalArgs a)
{
    __shared__ int s_a[kEvalThreads / 32], s_b[kEvalThreads / 32];
    __shared__ int s_carry;
    const int img = blockIdx.x;
    const int tid = threadIdx.x;
    const int n0 = a.pred_off[img], N = a.pred_off[img + 1] - n0;
    const int g0 = a.gt_off[img], G = a.gt_off[img + 1] - g0;
    // `if len(gt_boxes) == 0 or len(pred_info) == 0: continue` (R/utils/utils_map.py:196-197)
    if (N <= 0 || G <= 0) return;
    const double *pred = a.pred + (size_t)n0 * 5;
    const double *gt = a.gt + (size_t)g0 * 4;
    int *rmax = a.rmax + (size_t)img * a.thresh_num;

    for (int g = tid; g < G; g += kEvalThreads) a.first_h[g0 + g] = 0x7fffffff;
    for (int t = tid; t < a.thresh_num; t += kEvalThreads) rmax[t] = -1;
    __syncthreads();

    // ---- 1+2: best GT per prediction, first match per kept GT
    for (int h = tid; h < N; h += kEvalThreads) {
        const double px1 = pred[h * 5], py1 = pred[h * 5 + 1];
        const double px2 = __dadd_rn(pred[h * 5 + 2], px1), py2 = __dadd_rn(pred[h * 5 + 3], py1); // :112-113
        const double parea = __dmul_rn(__dsub_rn(px2, px1), __dsub_rn(py2, py1));
        double best = 0.0;
        int bi = 0;
        bool best_nan = false;
        for (int g = 0; g < G; ++g) {
            const double gx1 = gt[g * 4], gy1 = gt[g * 4 + 1];
            const double gx2 = __dadd_rn(gt[g * 4 + 2], gx1), gy2 = __dadd_rn(gt[g * 4 + 3], gy1);   // :114-115
            const double garea = __dmul_rn(__dsub_rn(gx2, gx1), __dsub_rn(gy2, gy1));
            const double v = overlap_f64(px1, py1, px2, py2, parea, gx1, gy1, gx2, gy2, garea);
            // np.max / np.argmax: the first NaN wins; otherwise the first maximum
            if (g == 0) { best = v; bi = 0; best_nan = v != v; }
            else if (!best_nan && (v != v || v > best)) { best = v; bi = g; best_nan = v != v; }
        }
        const bool match = best >= a.iou_thresh; // false for NaN
        a.best_gt[n0 + h] = match ? bi : -1;
        if (match && a.keep[g0 + bi]) atomicMin(&a.first_h[g0 + bi], h);
    }
    eval_rmax(pred, N, a.thresh_num, rmax);
    __syncthreads();

    // ---- 3: pred_recall[h] and the count of proposal_list[:h+1] == 1, as inclusive scans over h
    int run_a = 0, run_b = 0;
    for (int base = 0; base < N; base += kEvalThreads) {
        const int h = base + tid;
        int fa = 0, fb = 0;
        if (h < N) {
            const int bg = a.best_gt[n0 + h];
            const bool kept_gt = bg >= 0 && a.keep[g0 + bg];
            fa = (kept_gt && a.first_h[g0 + bg] == h) ? 1 : 0; // recall_list[bg]: 0 -> 1 at this prediction
            fb = (bg >= 0 && !a.keep[g0 + bg]) ? 0 : 1;        // proposal_list[h] == 1
        }
        int ia, ib, ta, tb;
        block_scan2(fa, fb, ia, ib, ta, tb, s_a, s_b);
        if (h < N) {
            a.pred_recall[n0 + h] = run_a + ia;
            a.proposal[n0 + h] = run_b + ib;
        }
        run_a += ta;
        run_b += tb;
    }
    __syncthreads();

    // ---- 4
    eval_thresholds(rmax, a.thresh_num, a.proposal + n0, a.pred_recall + n0, a.pr_curve, s_a, &s_carry);
}

// img_pr_info (R/utils/utils_map.py:135-148) on its own: one image, image_eval's outputs given.
__global__ void __launch_bounds__(kEvalThreads) img_pr_info_kernel(const double *pred, int N, const int *proposal_list,
                                                                   const int *pred_recall, int thresh_num, double *pr_info,
                                                                   int *rmax, int *prop_cum)
{
    __shared__ int s_a[kEvalThreads / 32], s_b[kEvalThreads / 32];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    for (int t = tid; t < thresh_num; t += kEvalThreads) { rmax[t] = -1; pr_info[2 * t] = 0.0; pr_info[2 * t + 1] = 0.0; }
    __syncthreads();
    eval_rmax(pred, N, thresh_num, rmax);
    int run = 0;
    for (int base = 0; base < N; base += kEvalThreads) {
        const int h = base + tid;
        const int f = (h < N && proposal_list[h] == 1) ? 1 : 0;
        int ia, ib, ta, tb;
        block_scan2(f, 0, ia, ib, ta, tb, s_a, s_b);
        if (h < N) prop_cum[h] = run + ia;
        run += ta;
    }
    __syncthreads();
    eval_thresholds(rmax, thresh_num, prop_cum, pred_recall, pr_info, s_a, &s_carry);
}

// bbox_overlaps (R/utils/utils_map.py:16-27): [A,4] x [B,4] point-form fp64 boxes -> [A,B]
__global__ void __launch_bounds__(256) bbox_overlaps_f64_kernel(const double *box_a, long long A, const double *box_b, long long B,
                                                                double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A * B) return;
    const double *p = box_a + (i / B) * 4, *q = box_b + (i % B) * 4;
    const double pa = __dmul_rn(__dsub_rn(p[2], p[0]), __dsub_rn(p[3], p[1]));
    const double qa = __dmul_rn(__dsub_rn(q[2], q[0]), __dsub_rn(q[3], q[1]));
    out[i] = overlap_f64(p[0], p[1], p[2], p[3], pa, q[0], q[1], q[2], q[3], qa);
}

// proposal[] holds the inclusive count while the kernel runs; image_eval's proposal_list is +-1 per prediction
__global__ void __launch_bounds__(kEvalThreads) wider_proposal_kernel(const int *pred_off, const int *gt_off, const int *best_gt,
                                                                      const unsigned char *keep, int *proposal, int I)
{
    const int img = blockIdx.x;
    const int n0 = pred_off[img], N = pred_off[img + 1] - n0;
    const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
    for (int h = threadIdx.x; h < N; h += kEvalThreads) {
        int v = 1;
        if (G > 0) {
            const int bg = best_gt[n0 + h];
            if (bg >= 0 && !keep[g0 + bg]) v = -1;
        }
        proposal[n0 + h] = v;
    }
}

// ---- norm_score (R/utils/utils_map.py:75-98): min-max normalisation of every score of the dataset --------------
// order-preserving double <-> uint64 so that min / max are integer atomics
__device__ __forceinline__ unsigned long long ord64(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord64_inv(unsigned long long u)
{
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

// mm[0] = ord64(min), mm[1] = ord64(max); initialised to min_score = 1 and max_score = 0 like the reference (:80-81)
__global__ void norm_init_kernel(unsigned long long *mm)
{
    mm[0] = ord64(1.0);
    mm[1] = ord64(0.0);
}

__global__ void __launch_bounds__(256) norm_minmax_kernel(const double *pred, long long n, unsigned long long *mm)
{
    unsigned long long lo = 0xffffffffffffffffull, hi = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long u = ord64(pred[i * 5 + 4]);
        lo = u < lo ? u : lo;
        hi = u > hi ? u : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(kFull, lo, o), h2 = __shfl_xor_sync(kFull, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if (lane_id() == 0) {
        atomicMin(mm, lo);
        atomicMax(mm + 1, hi);
    }
}

__global__ void __launch_bounds__(256) norm_apply_kernel(double *pred, long long n, const unsigned long long *mm)
{
    const double mn = ord64_inv(mm[0]), mx = ord64_inv(mm[1]);
    const double diff = __dsub_rn(mx, mn);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        pred[i * 5 + 4] = __ddiv_rn(__dsub_rn(pred[i * 5 + 4], mn), diff); // v[:, -1] = (v[:, -1] - min_score)/diff
}

static size_t eval_ws_layout(int I, int64_t sumN, int64_t sumG, int thresh_num, size_t *o_best, size_t *o_first, size_t *o_rmax,
                             size_t *o_rec, size_t *o_prop)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes ? bytes : 1, 256); return o; };
    const size_t b = take(sizeof(int) * (size_t)sumN), f = take(sizeof(int) * (size_t)sumG);
    const size_t r = take(sizeof(int) * (size_t)I * (size_t)thresh_num);
    const size_t rec = take(sizeof(int) * (size_t)sumN), prop = take(sizeof(int) * (size_t)sumN);
    if (o_best) { *o_best = b; *o_first = f; *o_rmax = r; *o_rec = rec; *o_prop = prop; }
    return off;
}

} // namespace jabd

using namespace jabd;

extern "C" {

size_t jabd_wider_eval_workspace_bytes(int I, int64_t sumN, int64_t sumG, int thresh_num)
{
    if (I < 0 || sumN < 0 || sumG < 0 || thresh_num < 0) return 0;
    return eval_ws_layout(I, sumN, sumG, thresh_num, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int jabd_bbox_overlaps_f64(const double *box_a, int64_t A, const double *box_b, int64_t B, double *out, jabd_stream_t stream)
{
    JABD_REQUIRE(A >= 0 && B >= 0, JABD_EINVAL, "bbox_overlaps_f64: negative size");
    if (A == 0 || B == 0) return JABD_OK;
    JABD_REQUIRE(box_a && box_b && out, JABD_EINVAL, "bbox_overlaps_f64: null pointer");
    JABD_REQUIRE(aligned_to(box_a, 8) && aligned_to(box_b, 8) && aligned_to(out, 8), JABD_EALIGN, "bbox_overlaps_f64: misaligned");
    JABD_REQUIRE(A * B < (1ll << 40), JABD_EINVAL, "bbox_overlaps_f64: A*B too large");
    bbox_overlaps_f64_kernel<<<(unsigned)((A * B + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(box_a, A, box_b, B, out);
    JABD_LAUNCH_CHECK("bbox_overlaps_f64_kernel");
    return JABD_OK;
}

int jabd_img_pr_info(const double *pred, int64_t N, const int *proposal_list, const int *pred_recall, int thresh_num,
                     double *pr_info, void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    JABD_REQUIRE(N >= 0 && thresh_num >= 0 && N < (1ll << 31), JABD_EINVAL, "img_pr_info: bad size");
    if (thresh_num == 0) return JABD_OK;
    JABD_REQUIRE(pr_info && ((pred && proposal_list && pred_recall) || N == 0), JABD_EINVAL, "img_pr_info: null pointer");
    const size_t need = round_up(sizeof(int) * (size_t)thresh_num, 256) + round_up(sizeof(int) * (size_t)(N ? N : 1), 256);
    JABD_REQUIRE(workspace && aligned_to(workspace, 256) && workspace_bytes >= need, JABD_EWORKSPACE,
                 "img_pr_info: needs %zu bytes of 256-byte aligned workspace", need);
    int *rmax = static_cast<int *>(workspace);
    int *prop_cum = reinterpret_cast<int *>(static_cast<char *>(workspace) + round_up(sizeof(int) * (size_t)thresh_num, 256));
    img_pr_info_kernel<<<1, kEvalThreads, 0, static_cast<cudaStream_t>(stream)>>>(pred, (int)N, proposal_list, pred_recall, thresh_num,
                                                                                  pr_info, rmax, prop_cum);
    JABD_LAUNCH_CHECK("img_pr_info_kernel");
    return JABD_OK;
}

int jabd_norm_score(double *pred, int64_t sumN, void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    JABD_REQUIRE(sumN >= 0, JABD_EINVAL, "norm_score: negative size");
    if (sumN == 0) return JABD_OK;
    JABD_REQUIRE(pred && aligned_to(pred, 8), JABD_EINVAL, "norm_score: pred null or misaligned");
    JABD_REQUIRE(workspace && aligned_to(workspace, 16) && workspace_bytes >= 16, JABD_EWORKSPACE,
                 "norm_score: needs 16 bytes of 16-byte aligned workspace");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *mm = static_cast<unsigned long long *>(workspace);
    const long long want = (sumN + 255) / 256;
    const unsigned blocks = (unsigned)(want < 1184 ? want : 1184);
    norm_init_kernel<<<1, 1, 0, st>>>(mm);
    norm_minmax_kernel<<<blocks, 256, 0, st>>>(pred, (long long)sumN, mm);
    norm_apply_kernel<<<blocks, 256, 0, st>>>(pred, (long long)sumN, mm);
    JABD_LAUNCH_CHECK("norm_score kernels");
    return JABD_OK;
}

int jabd_wider_eval(const double *pred, const int *pred_off, const double *gt, const int *gt_off, const unsigned char *keep, int I,
                    int64_t sumN, int64_t sumG, double iou_thresh, int thresh_num, double *pr_curve, int *pred_recall,
                    int *proposal_list, void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    JABD_REQUIRE(I >= 0 && sumN >= 0 && sumG >= 0 && thresh_num >= 0, JABD_EINVAL, "wider_eval: negative size");
    JABD_REQUIRE(sumN < (1ll << 31) && sumG < (1ll << 31), JABD_EINVAL, "wider_eval: more than 2^31 rows");
    if (I == 0 || thresh_num == 0) return JABD_OK;
    JABD_REQUIRE(pred_off && gt_off && pr_curve && (pred || sumN == 0) && ((gt && keep) || sumG == 0), JABD_EINVAL,
                 "wider_eval: null pointer");
    JABD_REQUIRE(aligned_to(pred, 8) && aligned_to(gt, 8) && aligned_to(pr_curve, 8), JABD_EALIGN,
                 "wider_eval: fp64 arrays must be 8-byte aligned");
    JABD_REQUIRE(workspace && aligned_to(workspace, 256), JABD_EWORKSPACE, "wider_eval: workspace null or not 256-byte aligned");
    size_t ob, of, orx, orec, oprop;
    const size_t need = eval_ws_layout(I, sumN, sumG, thresh_num, &ob, &of, &orx, &orec, &oprop);
    JABD_REQUIRE(workspace_bytes >= need, JABD_EWORKSPACE, "wider_eval: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    char *base = static_cast<char *>(workspace);
    EvalArgs a;
    a.pred = pred; a.pred_off = pred_off; a.gt = gt; a.gt_off = gt_off; a.keep = keep;
    a.iou_thresh = iou_thresh; a.thresh_num = thresh_num; a.pr_curve = pr_curve;
    a.best_gt = reinterpret_cast<int *>(base + ob);
    a.first_h = reinterpret_cast<int *>(base + of);
    a.rmax = reinterpret_cast<int *>(base + orx);
    a.pred_recall = pred_recall ? pred_recall : reinterpret_cast<int *>(base + orec);
    a.proposal = reinterpret_cast<int *>(base + oprop);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    wider_eval_kernel<<<(unsigned)I, kEvalThreads, 0, st>>>(a);
    JABD_LAUNCH_CHECK("wider_eval_kernel");
    if (proposal_list) {
        wider_proposal_kernel<<<(unsigned)I, kEvalThreads, 0, st>>>(pred_off, gt_off, a.best_gt, keep, proposal_list, I);
        JABD_LAUNCH_CHECK("wider_proposal_kernel");
    }
    return JABD_OK;
}

} // extern "C"
