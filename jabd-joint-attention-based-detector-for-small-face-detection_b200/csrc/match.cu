// match.cu -- batched target assignment: priors x GT IoU, both argmaxes, force-match, SSD encode.
//
// Replaces the per-image loop of MultiBoxLoss.forward (R/nets/retinaface_training.py:197-214) and match()
// (:93-162).  Three launches per batch, all on the caller's stream, nothing allocated, no host sync:
//
//   stage_gt_kernel     one CTA per image: GT rows -> packed 16-byte boxes in the workspace, argmax keys
//                       cleared, a per-image "well formed" flag (all areas finite and >= 0).
//   match_argmax_kernel one CTA per (image, 256-prior tile), one prior per thread.  The image's GT boxes are
//                       staged into shared memory by the TMA engine (cp.async.bulk + mbarrier, 512 boxes per
//                       copy, next copy in flight while the current chunk is consumed).  Best GT per prior
//                       (overlaps.max(0), :120) stays in registers -- GT are visited in ascending order and
//                       only a strictly larger IoU replaces the best, so ties keep the lowest index.  Best
//                       prior per GT (overlaps.max(1), :111) is a cross-CTA argmax: warp REDUX on the IoU
//                       bits, lowest lane on ties, then a 64-bit (ordered IoU bits | ~index) atomicMax, first
//                       in shared memory, once per CTA in global memory.  The [G,P] matrix is never stored.
//   match_encode_kernel one CTA per (image, tile): force-match (:127-130, largest j wins), gather of the
//                       matched GT row, threshold (:143), encode (:61-84), coalesced stores.
//
// Spatial culling (default): a GT whose box does not intersect the bounding box of the CTA's prior tile has
// IoU == +0 with every prior of the tile (all inputs well formed), which can neither replace a per-prior best
// (strict >) nor beat the (0, index 0) row default, so it is dropped from the CTA's list; inside the loop a
// warp skips the divide when no lane has a positive intersection.  JABD_ASSIGN_DENSE disables both and
// evaluates all P*G pairs; results are identical (tests compare the two).  If any GT of the image or any prior
// of the tile is malformed (negative/non-finite area) the CTA takes the generic dense path.
#include "common.cuh"

namespace jabd {

constexpr int kTile = 256;  // priors per CTA, one per thread
constexpr int kChunk = 512; // GT boxes per bulk copy (8 KB)

struct AssignWorkspace {
    unsigned long long *keys; // [sumG] best-prior argmax keys
    float4 *gtbox;            // [sumG] x1 y1 x2 y2
    int *bti;                 // [B,P] best GT per prior (before the force-match override)
    float *bto;               // [B,P] its IoU
    int *img_ok;              // [B]
};

static size_t assign_ws_layout(int B, int64_t P, int64_t sumG, AssignWorkspace *w, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += round_up(bytes, 256);
        return o;
    };
    size_t o_keys = take(sizeof(unsigned long long) * (size_t)(sumG > 0 ? sumG : 1));
    size_t o_box = take(sizeof(float4) * (size_t)(sumG > 0 ? sumG : 1));
    size_t o_bti = take(sizeof(int) * (size_t)B * (size_t)P);
    size_t o_bto = take(sizeof(float) * (size_t)B * (size_t)P);
    size_t o_ok = take(sizeof(int) * (size_t)(B > 0 ? B : 1));
    if (w) {
        w->keys = reinterpret_cast<unsigned long long *>(base + o_keys);
        w->gtbox = reinterpret_cast<float4 *>(base + o_box);
        w->bti = reinterpret_cast<int *>(base + o_bti);
        w->bto = reinterpret_cast<float *>(base + o_bto);
        w->img_ok = reinterpret_cast<int *>(base + o_ok);
    }
    return off;
}

// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stage_gt_kernel(const float *__restrict__ gt, const int *__restrict__ gt_off,
                                                       AssignWorkspace ws)
{
    const int b = blockIdx.x;
    const int g0 = gt_off[b];
    const int G = gt_off[b + 1] - g0;
    int ok = 1;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float *r = gt + (size_t)(g0 + g) * JABD_GT_ROW;
        const float4 a = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
        ws.gtbox[g0 + g] = a;
        ws.keys[g0 + g] = 0ull;
        const float aa = box_area(a);
        ok &= (aa >= 0.0f && aa < CUDART_INF_F) ? 1 : 0;
    }
    ok = __syncthreads_and(ok);
    if (threadIdx.x == 0) ws.img_ok[b] = ok;
}

// -------------------------------------------------------------------------------------------------------
struct TileSmem {
    float4 raw[kChunk];            // TMA destination
    float4 box[kChunk];            // compacted GT list of this CTA
    unsigned long long key[kChunk];// per-list-entry best-prior key (0 = no proposal)
    float area[kChunk];
    int gidx[kChunk];              // image-local GT index of the list entry
    float red[4][kTile / 32];
    int wcnt[kTile / 32];
    uint64_t mbar;
};

// propose (ord, p) for list entry j; plain read first so that most losers never issue the atomic
__device__ __forceinline__ void propose(unsigned long long *slot, unsigned long long k)
{
    if (k > *reinterpret_cast<volatile unsigned long long *>(slot)) atomicMax(slot, k);
}

// MODE 0: culled list + zero-intersection skip (well-formed inputs)
// MODE 1: dense, well-formed inputs (every pair evaluated, no NaN bookkeeping)
// MODE 2: dense, generic (first-element initialisation, NaN wins, negative values ordered)
template <int MODE>
__device__ __forceinline__ void consume_list(TileSmem &s, int n_list, float4 pb, float area_p, bool valid,
                                             uint32_t p, float &best, int &bidx, bool &have_best)
{
    const unsigned lane = lane_id();
#pragma unroll 2
    for (int j = 0; j < n_list; ++j) {
        const float4 a = s.box[j];
        const float w = fsub(fminf(a.z, pb.z), fmaxf(a.x, pb.x));
        const float h = fsub(fminf(a.w, pb.w), fmaxf(a.y, pb.y));
        if (MODE == 0) {
            const bool pos = (w > 0.0f) && (h > 0.0f);
            if (!__any_sync(kFull, pos)) continue; // exact: every lane's IoU is +0
            float v = 0.0f;
            if (pos) v = fdiv(fmul(w, h), fsub(fadd(s.area[j], area_p), fmul(w, h)));
            if (v > best) { best = v; bidx = s.gidx[j]; }
            const uint32_t bits = __float_as_uint(v); // v >= 0: raw bits are ordered
            const uint32_t wmax = __reduce_max_sync(kFull, bits);
            if (wmax != 0u) {
                const unsigned eq = __ballot_sync(kFull, bits == wmax);
                if (lane == (unsigned)(__ffs(eq) - 1)) propose(&s.key[j], make_key(wmax | 0x80000000u, p));
            }
        } else {
            const float inter = fmul(clamp0(w), clamp0(h));
            const float v = fdiv(inter, fsub(fadd(s.area[j], area_p), inter));
            uint32_t u;
            if (MODE == 1) {
                if (v > best) { best = v; bidx = s.gidx[j]; }
                u = __float_as_uint(v) | 0x80000000u;
            } else {
                if (valid) {
                    if (!have_best) { best = v; bidx = s.gidx[j]; have_best = true; }
                    else if (!(best != best) && ((v != v) || v > best)) { best = v; bidx = s.gidx[j]; }
                }
                u = valid ? ord_of(v) : 0u;
            }
            const uint32_t wmax = __reduce_max_sync(kFull, u);
            const bool worth = (MODE == 1) ? (wmax != 0x80000000u) : (wmax != 0u);
            if (worth) {
                const unsigned eq = __ballot_sync(kFull, u == wmax);
                if (lane == (unsigned)(__ffs(eq) - 1)) propose(&s.key[j], make_key(wmax, p));
            }
        }
    }
}

__global__ void __launch_bounds__(kTile) match_argmax_kernel(const float4 *__restrict__ priors, int P,
                                                             const int *__restrict__ gt_off, AssignWorkspace ws, int dense)
{
    __shared__ __align__(16) TileSmem s;
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = tid >> 5;
    const int g0 = gt_off[b];
    const int G = gt_off[b + 1] - g0;
    if (G <= 0) return; // encode kernel writes zeros for this image

    const int p = blockIdx.x * kTile + tid;
    const bool valid = p < P;
    // out-of-range threads carry a box that never has a positive intersection
    float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
    float area_p = 1.0f;
    if (valid) {
        pb = to_point_form(__ldg(priors + p));
        area_p = box_area(pb);
    }

    if (tid == 0) mbar_init(&s.mbar, 1);
    __syncthreads();
    uint32_t phase = 0;
    if (tid == 0) { // first GT chunk: in flight while the tile's bounding box is reduced
        const int n0 = G < kChunk ? G : kChunk;
        mbar_arrive_expect_tx(&s.mbar, (uint32_t)n0 * 16u);
        bulk_g2s(s.raw, ws.gtbox + g0, (uint32_t)n0 * 16u, &s.mbar);
    }

    // tile bounding box + prior sanity (block reduction)
    const bool prior_ok = !valid || (area_p > 0.0f && area_p < CUDART_INF_F);
    float bx1 = valid ? pb.x : CUDART_INF_F, by1 = valid ? pb.y : CUDART_INF_F;
    float bx2 = valid ? pb.z : -CUDART_INF_F, by2 = valid ? pb.w : -CUDART_INF_F;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx1 = fminf(bx1, __shfl_xor_sync(kFull, bx1, o));
        by1 = fminf(by1, __shfl_xor_sync(kFull, by1, o));
        bx2 = fmaxf(bx2, __shfl_xor_sync(kFull, bx2, o));
        by2 = fmaxf(by2, __shfl_xor_sync(kFull, by2, o));
    }
    if (lane == 0) { s.red[0][warp] = bx1; s.red[1][warp] = by1; s.red[2][warp] = bx2; s.red[3][warp] = by2; }
    const int tile_ok = __syncthreads_and(prior_ok ? 1 : 0);
#pragma unroll
    for (int w = 0; w < kTile / 32; ++w) {
        bx1 = fminf(bx1, s.red[0][w]); by1 = fminf(by1, s.red[1][w]);
        bx2 = fmaxf(bx2, s.red[2][w]); by2 = fmaxf(by2, s.red[3][w]);
    }
    const int mode = (tile_ok && ws.img_ok[b]) ? (dense ? 1 : 0) : 2;

    // The (IoU 0, prior 0) default of every row: proposed once by the tile that owns prior 0.  In mode 2
    // prior 0 proposes its real value inside the loop instead.
    if (blockIdx.x == 0 && mode != 2) {
        const unsigned long long k0 = make_key(0x80000000u, 0u);
        for (int g = tid; g < G; g += kTile) atomicMax(ws.keys + g0 + g, k0);
    }

    float best = 0.0f;
    int bidx = 0;
    bool have_best = false;

    for (int c0 = 0; c0 < G; c0 += kChunk) {
        const int n = (G - c0) < kChunk ? (G - c0) : kChunk;
        mbar_wait(&s.mbar, phase);
        phase ^= 1u;
        // build this CTA's list (order preserving, so ties keep the lowest GT index)
        int n_list = 0;
        for (int base = 0; base < n; base += kTile) {
            const int i = base + tid;
            const bool have = i < n;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have) a = s.raw[i];
            bool keep = have;
            if (mode == 0) {
                const float w = fsub(fminf(a.z, bx2), fmaxf(a.x, bx1));
                const float h = fsub(fminf(a.w, by2), fmaxf(a.y, by1));
                keep = have && (w > 0.0f) && (h > 0.0f);
            }
            const unsigned bal = __ballot_sync(kFull, keep);
            if (lane == 0) s.wcnt[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kTile / 32; ++w) {
                const int c = s.wcnt[w];
                woff += (w < warp) ? c : 0;
                tot += c;
            }
            if (keep) {
                const int pos = n_list + woff + __popc(bal & lanemask_lt());
                s.box[pos] = a;
                s.area[pos] = box_area(a);
                s.gidx[pos] = c0 + i;
                s.key[pos] = 0ull;
            }
            n_list += tot;
            __syncthreads();
        }
        // raw[] is dead: prefetch the next chunk behind the main loop
        if (tid == 0 && c0 + kChunk < G) {
            const int nn = (G - c0 - kChunk) < kChunk ? (G - c0 - kChunk) : kChunk;
            mbar_arrive_expect_tx(&s.mbar, (uint32_t)nn * 16u);
            bulk_g2s(s.raw, ws.gtbox + g0 + c0 + kChunk, (uint32_t)nn * 16u, &s.mbar);
        }
        if (mode == 0) consume_list<0>(s, n_list, pb, area_p, valid, (uint32_t)p, best, bidx, have_best);
        else if (mode == 1) consume_list<1>(s, n_list, pb, area_p, valid, (uint32_t)p, best, bidx, have_best);
        else consume_list<2>(s, n_list, pb, area_p, valid, (uint32_t)p, best, bidx, have_best);
        __syncthreads();
        for (int j = tid; j < n_list; j += kTile) {
            const unsigned long long k = s.key[j];
            if (k != 0ull) atomicMax(ws.keys + g0 + s.gidx[j], k);
        }
        __syncthreads();
    }
    if (valid) {
        ws.bti[(size_t)b * P + p] = bidx;
        ws.bto[(size_t)b * P + p] = best;
    }
}

// -------------------------------------------------------------------------------------------------------
struct EncodeArgs {
    const float4 *priors;
    const float *gt;
    const int *gt_off;
    int P;
    float threshold, var0, var1;
    int label_mode, encode_mode;
    float4 *loc_t;
    long long *conf_t;
    float *landm_t;
    int *out_bti;
    float *out_bto;
    int *out_bpi;
    float *out_bpo;
};

__global__ void __launch_bounds__(kTile) match_encode_kernel(EncodeArgs a, AssignWorkspace ws)
{
    __shared__ int s_forced[kTile];
    __shared__ __align__(16) float s_lm[kTile * 10];
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int p0 = blockIdx.x * kTile;
    const int p = p0 + tid;
    const int P = a.P;
    const bool valid = p < P;
    const int n_valid = (P - p0) < kTile ? (P - p0) : kTile;
    const int g0 = a.gt_off[b];
    const int G = a.gt_off[b + 1] - g0;
    const size_t row = (size_t)b * P + p;

    s_forced[tid] = -1;
    __syncthreads();
    // force-match: best_truth_idx[best_prior_idx[j]] = j for j ascending -> the largest j wins (:129-130)
    for (int g = tid; g < G; g += kTile) {
        const unsigned long long k = ws.keys[g0 + g];
        const uint32_t bp = key_idx(k);
        if (bp >= (uint32_t)p0 && bp < (uint32_t)(p0 + kTile)) atomicMax(&s_forced[bp - p0], g);
        if (blockIdx.x == 0) {
            if (a.out_bpi) a.out_bpi[g0 + g] = (int)bp;
            if (a.out_bpo) a.out_bpo[g0 + g] = ord_inv(key_ord(k));
        }
    }
    __syncthreads();

    float4 loc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long conf = 0;
    float lm[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) lm[k] = 0.0f;
    int idx = 0;
    float ov = 0.0f;
    if (valid && G > 0) {
        idx = ws.bti[row];
        ov = ws.bto[row];
        const int f = s_forced[tid];
        if (f >= 0) { idx = f; ov = 2.0f; } // :127
        const float *r = a.gt + (size_t)(g0 + idx) * JABD_GT_ROW;
        const float4 m = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
        const float4 pr = __ldg(a.priors + p);
        float c = __ldg(r + 14);
        if (a.label_mode) c = fadd(c, 1.0f);       // R/utils/box_utils.py:315
        if (ov < a.threshold) c = 0.0f;            // :143
        conf = (long long)c;                       // float -> int64 store truncates
        loc = a.encode_mode ? encode_box(m, pr, a.var0, a.var1) : m;
        if (a.landm_t) {
            const float dx = fmul(a.var0, pr.z), dy = fmul(a.var0, pr.w);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                lm[2 * k] = fdiv(fsub(__ldg(r + 4 + 2 * k), pr.x), dx);
                lm[2 * k + 1] = fdiv(fsub(__ldg(r + 5 + 2 * k), pr.y), dy);
            }
        }
    }
    if (valid) {
        a.loc_t[row] = loc;
        a.conf_t[row] = conf;
        if (a.out_bti) a.out_bti[row] = idx;
        if (a.out_bto) a.out_bto[row] = ov;
    }
    if (a.landm_t) {
        // [tile,10] rows through shared memory so the global stores are contiguous 16-byte vectors
#pragma unroll
        for (int k = 0; k < 10; ++k) s_lm[tid * 10 + k] = lm[k];
        __syncthreads();
        float *dst = a.landm_t + ((size_t)b * P + p0) * 10;
        const int nf = n_valid * 10;
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            const int n4 = nf >> 2;
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            const float4 *s4 = reinterpret_cast<const float4 *>(s_lm);
            for (int i = tid; i < n4; i += kTile) d4[i] = s4[i];
            for (int i = (n4 << 2) + tid; i < nf; i += kTile) dst[i] = s_lm[i];
        } else {
            for (int i = tid; i < nf; i += kTile) dst[i] = s_lm[i];
        }
    }
}

// -------------------------------------------------------------------------------------------------------
static int check_assign_common(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG,
                               void *workspace, size_t workspace_bytes)
{
    JABD_REQUIRE(B >= 0 && P >= 0 && sumG >= 0, JABD_EINVAL, "assign: negative size (B=%d P=%lld sumG=%lld)", B,
                 (long long)P, (long long)sumG);
    JABD_REQUIRE(B <= 65535, JABD_EINVAL, "assign: B=%d exceeds 65535 images per call", B);
    JABD_REQUIRE(P < (1ll << 31) - kTile && sumG < (1ll << 31), JABD_EINVAL, "assign: P or sumG exceeds int32 range");
    JABD_REQUIRE((int64_t)B * P < (1ll << 40), JABD_EINVAL, "assign: B*P too large");
    if (B == 0 || P == 0) return JABD_OK;
    JABD_REQUIRE(priors && gt_off && (gt || sumG == 0), JABD_EINVAL, "assign: null input pointer");
    JABD_REQUIRE(aligned_to(priors, 16), JABD_EALIGN, "assign: priors must be 16-byte aligned");
    JABD_REQUIRE(aligned_to(gt, 4) && aligned_to(gt_off, 4), JABD_EALIGN, "assign: gt/gt_off must be 4-byte aligned");
    JABD_REQUIRE(workspace != nullptr, JABD_EWORKSPACE, "assign: workspace is null");
    JABD_REQUIRE(aligned_to(workspace, 256), JABD_EALIGN, "assign: workspace must be 256-byte aligned");
    const size_t need = assign_ws_layout(B, P, sumG, nullptr, nullptr);
    JABD_REQUIRE(workspace_bytes >= need, JABD_EWORKSPACE, "assign: workspace too small (%zu < %zu bytes)", workspace_bytes,
                 need);
    return JABD_OK;
}

} // namespace jabd

using namespace jabd;

extern "C" {

size_t jabd_assign_workspace_bytes(int B, int64_t P, int64_t sumG)
{
    if (B < 0 || P < 0 || sumG < 0) return 0;
    return assign_ws_layout(B, P, sumG, nullptr, nullptr);
}

int jabd_assign_match(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG, int flags,
                      void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    int rc = check_assign_common(priors, P, gt, gt_off, B, sumG, workspace, workspace_bytes);
    if (rc != JABD_OK || B == 0 || P == 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AssignWorkspace ws;
    assign_ws_layout(B, P, sumG, &ws, static_cast<char *>(workspace));
    stage_gt_kernel<<<B, 256, 0, st>>>(gt, gt_off, ws);
    JABD_LAUNCH_CHECK("stage_gt_kernel");
    const dim3 grid((unsigned)((P + kTile - 1) / kTile), (unsigned)B);
    match_argmax_kernel<<<grid, kTile, 0, st>>>(reinterpret_cast<const float4 *>(priors), (int)P, gt_off, ws,
                                                (flags & JABD_ASSIGN_DENSE) ? 1 : 0);
    JABD_LAUNCH_CHECK("match_argmax_kernel");
    return JABD_OK;
}

int jabd_assign_encode(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG,
                       float threshold, float var0, float var1, int label_mode, int encode_mode, float *loc_t,
                       int64_t *conf_t, float *landm_t, int *best_truth_idx, float *best_truth_overlap, int *best_prior_idx,
                       float *best_prior_overlap, void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    int rc = check_assign_common(priors, P, gt, gt_off, B, sumG, workspace, workspace_bytes);
    if (rc != JABD_OK || B == 0 || P == 0) return rc;
    JABD_REQUIRE(loc_t && conf_t, JABD_EINVAL, "assign: loc_t/conf_t must not be null");
    JABD_REQUIRE(aligned_to(loc_t, 16) && aligned_to(conf_t, 8) && aligned_to(landm_t, 4), JABD_EALIGN,
                 "assign: loc_t needs 16-byte, conf_t 8-byte, landm_t 4-byte alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AssignWorkspace ws;
    assign_ws_layout(B, P, sumG, &ws, static_cast<char *>(workspace));
    EncodeArgs a;
    a.priors = reinterpret_cast<const float4 *>(priors);
    a.gt = gt;
    a.gt_off = gt_off;
    a.P = (int)P;
    a.threshold = threshold;
    a.var0 = var0;
    a.var1 = var1;
    a.label_mode = label_mode;
    a.encode_mode = encode_mode;
    a.loc_t = reinterpret_cast<float4 *>(loc_t);
    a.conf_t = reinterpret_cast<long long *>(conf_t);
    a.landm_t = landm_t;
    a.out_bti = best_truth_idx;
    a.out_bto = best_truth_overlap;
    a.out_bpi = best_prior_idx;
    a.out_bpo = best_prior_overlap;
    const dim3 grid((unsigned)((P + kTile - 1) / kTile), (unsigned)B);
    match_encode_kernel<<<grid, kTile, 0, st>>>(a, ws);
    JABD_LAUNCH_CHECK("match_encode_kernel");
    return JABD_OK;
}

int jabd_assign(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG, float threshold,
                float var0, float var1, int label_mode, int encode_mode, int flags, float *loc_t, int64_t *conf_t,
                float *landm_t, int *best_truth_idx, float *best_truth_overlap, int *best_prior_idx,
                float *best_prior_overlap, void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    int rc = jabd_assign_match(priors, P, gt, gt_off, B, sumG, flags, workspace, workspace_bytes, stream);
    if (rc != JABD_OK) return rc;
    return jabd_assign_encode(priors, P, gt, gt_off, B, sumG, threshold, var0, var1, label_mode, encode_mode, loc_t, conf_t,
                              landm_t, best_truth_idx, best_truth_overlap, best_prior_idx, best_prior_overlap, workspace,
                              workspace_bytes, stream);
}

size_t jabd_assign_host_scratch_bytes(int B, int64_t P, int64_t sumG, int with_landm)
{
    if (B < 0 || P < 0 || sumG < 0) return 0;
    size_t n = round_up(sizeof(float) * JABD_GT_ROW * (size_t)(sumG > 0 ? sumG : 1), 256);
    n += round_up(sizeof(int) * (size_t)(B + 1), 256);
    n += round_up(sizeof(float) * 4 * (size_t)B * P, 256);
    n += round_up(sizeof(int64_t) * (size_t)B * P, 256);
    if (with_landm) n += round_up(sizeof(float) * 10 * (size_t)B * P, 256);
    n += assign_ws_layout(B, P, sumG, nullptr, nullptr);
    return n;
}

int jabd_assign_host(const float *priors_dev, int64_t P, const float *gt_host, const int *gt_off_host, int B, float threshold,
                     float var0, float var1, int label_mode, int encode_mode, int flags, float *loc_t_host,
                     int64_t *conf_t_host, float *landm_t_host, void *dev_scratch, size_t dev_scratch_bytes,
                     jabd_stream_t stream)
{
    JABD_REQUIRE(B >= 0 && P >= 0, JABD_EINVAL, "assign_host: negative size");
    if (B == 0 || P == 0) return JABD_OK;
    JABD_REQUIRE(gt_off_host && loc_t_host && conf_t_host, JABD_EINVAL, "assign_host: null host pointer");
    const int64_t sumG = gt_off_host[B];
    JABD_REQUIRE(gt_off_host[0] == 0 && sumG >= 0, JABD_EINVAL, "assign_host: gt_off must start at 0 and be non-decreasing");
    for (int b = 0; b < B; ++b)
        JABD_REQUIRE(gt_off_host[b + 1] >= gt_off_host[b], JABD_EINVAL, "assign_host: gt_off decreases at image %d", b);
    JABD_REQUIRE(gt_host || sumG == 0, JABD_EINVAL, "assign_host: gt_host is null");
    const int with_landm = landm_t_host != nullptr;
    JABD_REQUIRE(dev_scratch && aligned_to(dev_scratch, 256), JABD_EWORKSPACE, "assign_host: dev_scratch null or not 256-byte aligned");
    JABD_REQUIRE(dev_scratch_bytes >= jabd_assign_host_scratch_bytes(B, P, sumG, with_landm), JABD_EWORKSPACE,
                 "assign_host: dev_scratch too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *base = static_cast<char *>(dev_scratch);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = base + off; off += round_up(bytes, 256); return q; };
    float *d_gt = reinterpret_cast<float *>(take(sizeof(float) * JABD_GT_ROW * (size_t)(sumG > 0 ? sumG : 1)));
    int *d_off = reinterpret_cast<int *>(take(sizeof(int) * (size_t)(B + 1)));
    float *d_loc = reinterpret_cast<float *>(take(sizeof(float) * 4 * (size_t)B * P));
    int64_t *d_conf = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * (size_t)B * P));
    float *d_landm = with_landm ? reinterpret_cast<float *>(take(sizeof(float) * 10 * (size_t)B * P)) : nullptr;
    void *d_ws = base + off;
    const size_t ws_bytes = dev_scratch_bytes - off;
    if (sumG > 0) JABD_CUDA(cudaMemcpyAsync(d_gt, gt_host, sizeof(float) * JABD_GT_ROW * (size_t)sumG, cudaMemcpyHostToDevice, st));
    JABD_CUDA(cudaMemcpyAsync(d_off, gt_off_host, sizeof(int) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
    int rc = jabd_assign(priors_dev, P, d_gt, d_off, B, sumG, threshold, var0, var1, label_mode, encode_mode, flags, d_loc,
                         d_conf, d_landm, nullptr, nullptr, nullptr, nullptr, d_ws, ws_bytes, stream);
    if (rc != JABD_OK) return rc;
    JABD_CUDA(cudaMemcpyAsync(loc_t_host, d_loc, sizeof(float) * 4 * (size_t)B * P, cudaMemcpyDeviceToHost, st));
    JABD_CUDA(cudaMemcpyAsync(conf_t_host, d_conf, sizeof(int64_t) * (size_t)B * P, cudaMemcpyDeviceToHost, st));
    if (with_landm)
        JABD_CUDA(cudaMemcpyAsync(landm_t_host, d_landm, sizeof(float) * 10 * (size_t)B * P, cudaMemcpyDeviceToHost, st));
    JABD_CUDA(cudaStreamSynchronize(st));
    return JABD_OK;
}

} // extern "C"
