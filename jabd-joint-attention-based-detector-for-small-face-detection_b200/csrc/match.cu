// match.cu -- batched target assignment: priors x GT IoU, both argmaxes, force-match, SSD encode.
//
// Replaces the per-image loop of MultiBoxLoss.forward (R/nets/retinaface_training.py:197-214) and match()
// (:93-162).  Three launches per batch, all on the caller's stream; nothing is allocated, there is no host
// sync, no atomic and no memset; the [G,P] IoU matrix is never stored.
//
//   assign_prep_kernel  image CTAs: GT rows -> packed 16-byte boxes in the workspace + a per-image "well
//                       formed" flag; tile CTAs: bounding box, smallest prior area and a sanity flag for
//                       every 256-prior tile (32 B of metadata per tile).
//   assign_match_kernel two roles in one launch, chosen by blockIdx.y:
//     column CTAs       (image, 256-prior tile), one prior per thread: best GT per prior (overlaps.max(0),
//                       :120).  The image's GT boxes are staged into shared memory by the TMA engine
//                       (cp.async.bulk + mbarrier, 1024 boxes per copy, double buffered).  Each warp tests 32 GT
//                       per ballot against the bounding box of its 32 priors and visits only the hits, four
//                       per step so that their loads and divisions overlap.  GT are visited in ascending
//                       order and only a strictly larger IoU replaces the best: ties keep the lowest index.
//     row warps         one warp per GT: best prior per GT (overlaps.max(1), :111).  32 tiles are tested per
//                       ballot against the GT box; a tile is skipped when it cannot intersect the GT or when
//                       area_gt / min_prior_area < current best (no prior of it can reach the best); the
//                       surviving tiles are scanned 8 priors per lane; a REDUX max on the IoU bits and a REDUX
//                       min on the prior index among the ties give the lowest index of the maximum.
//   match_encode_kernel one CTA per (image, tile): force-match (:127-130, largest j wins), gather of the
//                       matched GT row, threshold (:143), encode (:61-84), coalesced stores.
//
// Why culling is exact: with well-formed inputs (finite coordinates, GT area in [0, 2^40], prior area in
// [2^-40, 2^40]) every IoU is >= +0 and a pair whose boxes do not intersect has IoU == +0 exactly, which can
// neither replace a best (strict >, both argmaxes start at value 0 / index 0 like torch.max over zeros) nor tie
// with a positive one.  JABD_ASSIGN_DENSE disables culling and pruning and evaluates all P*G pairs twice (once per
// argmax); tests compare the two bit for bit.  Malformed inputs (negative or non-finite areas) select a generic
// dense path with torch.max's NaN/ordering semantics per warp.
#include "common.cuh"

namespace jabd {

constexpr int kTile = 256;    // priors per column CTA, one per thread
constexpr int kChunk = 1024;  // GT boxes per bulk copy (16 KB), two buffers
constexpr int kMaxRowCtas = 64; // row-role CTAs per image (8 warps each, one GT per warp per pass)
constexpr int kWide = 4;      // GT hits processed per step by a column warp

struct TileMeta {
    float4 box; // bounding box of the tile's priors (point form)
    float amin; // smallest prior area of the tile
    int ok;     // every prior area within [2^-40, 2^40]
    int pad0, pad1;
};

struct AssignWorkspace {
    float4 *gtbox;   // [sumG] x1 y1 x2 y2
    int *bpi;        // [sumG] best prior per GT
    float *bpo;      // [sumG] its IoU
    TileMeta *tiles; // [ceil(P/256)]
    int *bti;        // [B,P] best GT per prior (before the force-match override)
    float *bto;      // [B,P] its IoU
    int *img_ok;     // [B]
};

static size_t assign_ws_layout(int B, int64_t P, int64_t sumG, AssignWorkspace *w, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += round_up(bytes, 256);
        return o;
    };
    const size_t ng = (size_t)(sumG > 0 ? sumG : 1);
    const size_t nt = (size_t)((P + kTile - 1) / kTile) + 1;
    size_t o_box = take(sizeof(float4) * ng);
    size_t o_bpi = take(sizeof(int) * ng);
    size_t o_bpo = take(sizeof(float) * ng);
    size_t o_tiles = take(sizeof(TileMeta) * nt);
    size_t o_bti = take(sizeof(int) * (size_t)B * (size_t)P);
    size_t o_bto = take(sizeof(float) * (size_t)B * (size_t)P);
    size_t o_ok = take(sizeof(int) * (size_t)(B > 0 ? B : 1));
    if (w) {
        w->gtbox = reinterpret_cast<float4 *>(base + o_box);
        w->bpi = reinterpret_cast<int *>(base + o_bpi);
        w->bpo = reinterpret_cast<float *>(base + o_bpo);
        w->tiles = reinterpret_cast<TileMeta *>(base + o_tiles);
        w->bti = reinterpret_cast<int *>(base + o_bti);
        w->bto = reinterpret_cast<float *>(base + o_bto);
        w->img_ok = reinterpret_cast<int *>(base + o_ok);
    }
    return off;
}

// -------------------------------------------------------------------------------------------------------
// blockIdx.x < B: image role; otherwise tile role.
__global__ void __launch_bounds__(kTile) assign_prep_kernel(const float *__restrict__ gt, const int *__restrict__ gt_off,
                                                            const float4 *__restrict__ priors, int P, int B, AssignWorkspace ws)
{
    __shared__ float red[5][kTile / 32];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < B) {
        const int b = blockIdx.x;
        const int g0 = gt_off[b];
        const int G = gt_off[b + 1] - g0;
        int ok = 1;
        for (int g = tid; g < G; g += kTile) {
            const float *r = gt + (size_t)(g0 + g) * JABD_GT_ROW;
            const float4 a = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
            ws.gtbox[g0 + g] = a;
            const float aa = box_area(a);
            ok &= (aa >= 0.0f && aa <= 0x1p40f) ? 1 : 0; // false for NaN / inf coordinates too
        }
        ok = __syncthreads_and(ok);
        if (tid == 0) ws.img_ok[b] = ok;
        return;
    }
    const int t = (int)blockIdx.x - B;
    const int p = t * kTile + tid;
    const bool valid = p < P;
    float x1 = CUDART_INF_F, y1 = CUDART_INF_F, x2 = -CUDART_INF_F, y2 = -CUDART_INF_F, amin = CUDART_INF_F;
    int ok = 1;
    if (valid) {
        const float4 pb = to_point_form(__ldg(priors + p));
        const float area = box_area(pb);
        x1 = pb.x; y1 = pb.y; x2 = pb.z; y2 = pb.w; amin = area;
        ok = (area >= 0x1p-40f && area <= 0x1p40f) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x1 = fminf(x1, __shfl_xor_sync(kFull, x1, o));
        y1 = fminf(y1, __shfl_xor_sync(kFull, y1, o));
        x2 = fmaxf(x2, __shfl_xor_sync(kFull, x2, o));
        y2 = fmaxf(y2, __shfl_xor_sync(kFull, y2, o));
        amin = fminf(amin, __shfl_xor_sync(kFull, amin, o));
    }
    if ((tid & 31) == 0) {
        const int w = tid >> 5;
        red[0][w] = x1; red[1][w] = y1; red[2][w] = x2; red[3][w] = y2; red[4][w] = amin;
    }
    ok = __syncthreads_and(ok);
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < kTile / 32; ++w) {
            x1 = fminf(x1, red[0][w]); y1 = fminf(y1, red[1][w]);
            x2 = fmaxf(x2, red[2][w]); y2 = fmaxf(y2, red[3][w]);
            amin = fminf(amin, red[4][w]);
        }
        TileMeta m;
        m.box = make_float4(x1, y1, x2, y2);
        m.amin = amin;
        m.ok = ok;
        m.pad0 = m.pad1 = 0;
        ws.tiles[t] = m;
    }
}

// -------------------------------------------------------------------------------------------------------
// IoU of a well-formed pair (union in [2^-40, 2^42], inter >= 0): the bits of __fdiv_rn(inter, union) unless the
// intersection is a non-zero value below 2^-60, which sets `slow` instead (the caller then redoes its whole block
// with the generic divide).  Straight-line code: callers unroll several pairs and rely on the compiler to overlap
// their loads, min/max chains and divisions -- a branch per pair would serialise them.
__device__ __forceinline__ float iou_wellformed(float4 a, float area_a, float4 b, float area_b, bool &slow)
{
    const float w = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float h = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float inter = fmul(fmaxf(w, 0.0f), fmaxf(h, 0.0f)); // == clamp(min=0) for finite values
    const float uni = fsub(fadd(area_a, area_b), inter);
    slow = slow || (inter < 0x1p-60f && inter != 0.0f);
    return fdiv_fast(inter, uni, rcp_refined(uni));
}

struct ColSmem {
    float4 raw[2][kChunk]; // TMA destinations
    uint64_t mbar[2];
};

// MODE 0: culled (ballot of 32 GT against the warp's bounding box), MODE 1: dense, MODE 2: dense generic
template <int MODE>
__device__ __forceinline__ void column_consume(const float4 *__restrict__ raw, int n, int c0, float4 pb, float area_p, float4 wbox,
                                               bool valid, float &best, int &bidx, bool &have_best)
{
    const unsigned lane = lane_id();
    for (int base = 0; base < n; base += 32) {
        const int e = base + (int)lane;
        bool hit = e < n;
        if (MODE == 0 && hit) {
            const float4 a = raw[e];
            const float w = fsub(fminf(a.z, wbox.z), fmaxf(a.x, wbox.x));
            const float h = fsub(fminf(a.w, wbox.w), fmaxf(a.y, wbox.y));
            hit = (w > 0.0f) && (h > 0.0f);
        }
        unsigned m = __ballot_sync(kFull, hit);
        if (MODE != 2) {
            while (m) {
                float v[kWide];
                int j[kWide];
                bool live[kWide];
                float4 a[kWide];
#pragma unroll
                for (int k = 0; k < kWide; ++k) {
                    j[k] = base + (m ? (__ffs(m) - 1) : 0);
                    live[k] = m != 0u;
                    m &= m - 1; // no-op once m == 0
                    a[k] = raw[j[k]];
                }
                bool slow = false;
#pragma unroll
                for (int k = 0; k < kWide; ++k) v[k] = iou_wellformed(a[k], box_area(a[k]), pb, area_p, slow);
                if (slow) { // a sliver intersection below 2^-60 somewhere in this step: generic IEEE divide
#pragma unroll
                    for (int k = 0; k < kWide; ++k) v[k] = iou_ref(a[k], box_area(a[k]), pb, area_p);
                }
#pragma unroll
                for (int k = 0; k < kWide; ++k) { // ascending GT index, strict > : ties keep the lowest index
                    if (live[k] && v[k] > best) { best = v[k]; bidx = c0 + j[k]; } // !live: padding of a short step
                }
            }
        } else {
            while (m) {
                const int j = base + __ffs(m) - 1;
                m &= m - 1;
                const float4 a = raw[j];
                const float v = iou_ref(a, box_area(a), pb, area_p);
                if (valid) {
                    if (!have_best) { best = v; bidx = c0 + j; have_best = true; }
                    else if (!(best != best) && ((v != v) || v > best)) { best = v; bidx = c0 + j; }
                }
            }
        }
    }
}

__device__ __forceinline__ void column_role(const float4 *__restrict__ priors, int P, int b, int tile, int g0, int G,
                                            const AssignWorkspace &ws, int dense, ColSmem &s)
{
    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int p = tile * kTile + tid;
    const bool valid = p < P;
    // out-of-range threads carry a box that never has a positive intersection
    float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
    float area_p = 1.0f;
    if (valid) {
        pb = to_point_form(__ldg(priors + p));
        area_p = box_area(pb);
    }
    if (tid == 0) { mbar_init(&s.mbar[0], 1); mbar_init(&s.mbar[1], 1); }
    __syncthreads();
    const int nchunks = (G + kChunk - 1) / kChunk;
    if (tid == 0) { // first GT chunk: in flight while the warp bounding boxes are reduced
        const int n0 = G < kChunk ? G : kChunk;
        mbar_arrive_expect_tx(&s.mbar[0], (uint32_t)n0 * 16u);
        bulk_g2s(s.raw[0], ws.gtbox + g0, (uint32_t)n0 * 16u, &s.mbar[0]);
    }
    const bool prior_ok = !valid || (area_p >= 0x1p-40f && area_p <= 0x1p40f);
    float bx1 = valid ? pb.x : CUDART_INF_F, by1 = valid ? pb.y : CUDART_INF_F;
    float bx2 = valid ? pb.z : -CUDART_INF_F, by2 = valid ? pb.w : -CUDART_INF_F;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx1 = fminf(bx1, __shfl_xor_sync(kFull, bx1, o));
        by1 = fminf(by1, __shfl_xor_sync(kFull, by1, o));
        bx2 = fmaxf(bx2, __shfl_xor_sync(kFull, bx2, o));
        by2 = fmaxf(by2, __shfl_xor_sync(kFull, by2, o));
    }
    const float4 wbox = make_float4(bx1, by1, bx2, by2);
    const int mode = (__all_sync(kFull, prior_ok) && ws.img_ok[b]) ? (dense ? 1 : 0) : 2; // per warp

    float best = 0.0f;
    int bidx = 0;
    bool have_best = false;
    uint32_t ph0 = 0, ph1 = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        const int c0 = c * kChunk;
        const int n = (G - c0) < kChunk ? (G - c0) : kChunk;
        if (tid == 0 && c + 1 < nchunks) { // the other buffer was released by the barrier that ended step c-1
            const int nn = (G - c0 - kChunk) < kChunk ? (G - c0 - kChunk) : kChunk;
            mbar_arrive_expect_tx(&s.mbar[buf ^ 1], (uint32_t)nn * 16u);
            bulk_g2s(s.raw[buf ^ 1], ws.gtbox + g0 + c0 + kChunk, (uint32_t)nn * 16u, &s.mbar[buf ^ 1]);
        }
        if (buf == 0) { mbar_wait(&s.mbar[0], ph0); ph0 ^= 1u; }
        else { mbar_wait(&s.mbar[1], ph1); ph1 ^= 1u; }
        if (mode == 0) column_consume<0>(s.raw[buf], n, c0, pb, area_p, wbox, valid, best, bidx, have_best);
        else if (mode == 1) column_consume<1>(s.raw[buf], n, c0, pb, area_p, wbox, valid, best, bidx, have_best);
        else column_consume<2>(s.raw[buf], n, c0, pb, area_p, wbox, valid, best, bidx, have_best);
        if (nchunks > 1) __syncthreads();
    }
    if (valid) {
        ws.bti[(size_t)b * P + p] = bidx;
        ws.bto[(size_t)b * P + p] = best;
    }
}

// One warp per GT: best prior (value, lowest index).
__device__ __forceinline__ void row_role(const float4 *__restrict__ priors, int P, int b, int r, int g0, int G,
                                         const AssignWorkspace &ws, int dense, int n_tiles, int row_ctas)
{
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    if (r * (kTile / 32) + warp >= G) return; // nothing for this warp
    bool ok = true;
    for (int t = (int)lane; t < n_tiles; t += 32) ok = ok && (ws.tiles[t].ok != 0);
    const int mode = (__all_sync(kFull, ok) && ws.img_ok[b]) ? (dense ? 1 : 0) : 2;
    for (int g = r * (kTile / 32) + warp; g < G; g += row_ctas * (kTile / 32)) {
        const float4 a = ws.gtbox[g0 + g];
        const float aa = box_area(a);
        float best = 0.0f, wbest = 0.0f;
        uint32_t bp = 0;
        bool have = false;
        for (int tbase = 0; tbase < n_tiles; tbase += 32) {
            const int t = tbase + (int)lane;
            bool hit = t < n_tiles;
            float amin = 0.0f;
            if ((mode == 0) && hit) {
                const TileMeta tm = ws.tiles[t];
                amin = tm.amin;
                const float w = fsub(fminf(a.z, tm.box.z), fmaxf(a.x, tm.box.x));
                const float h = fsub(fminf(a.w, tm.box.w), fmaxf(a.y, tm.box.y));
                hit = (w > 0.0f) && (h > 0.0f);
            }
            unsigned m = __ballot_sync(kFull, hit);
            while (m) {
                const int tl = __ffs(m) - 1;
                m &= m - 1;
                if ((mode == 0)) {
                    // IoU <= area_gt / area_prior: the tile cannot reach the current best (margin covers rounding)
                    const float am = __shfl_sync(kFull, amin, tl);
                    if (aa < fmul(fmul(wbest, am), 0.999999f)) continue;
                }
                const int pbase = (tbase + tl) * kTile + (int)lane;
                if (mode != 2) {
                    float v[kTile / 32];
                    float4 pr[kTile / 32];
#pragma unroll
                    for (int k = 0; k < kTile / 32; ++k) { // all loads first (clamped index: always in range)
                        const int p = pbase + 32 * k;
                        pr[k] = __ldg(priors + (p < P ? p : P - 1));
                    }
                    bool slow = false;
#pragma unroll
                    for (int k = 0; k < kTile / 32; ++k) {
                        const float4 pb = to_point_form(pr[k]);
                        v[k] = iou_wellformed(a, aa, pb, box_area(pb), slow);
                    }
                    if (slow) {
#pragma unroll
                        for (int k = 0; k < kTile / 32; ++k) {
                            const float4 pb = to_point_form(pr[k]);
                            v[k] = iou_ref(a, aa, pb, box_area(pb));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kTile / 32; ++k) { // ascending prior index per lane, strict >
                        const int p = pbase + 32 * k;
                        if (p < P && v[k] > best) { best = v[k]; bp = (uint32_t)p; }
                    }
                    wbest = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(best)));
                } else {
                    for (int k = 0; k < kTile / 32; ++k) {
                        const int p = pbase + 32 * k;
                        if (p < P) {
                            const float4 pb = to_point_form(__ldg(priors + p));
                            const float v = iou_ref(a, aa, pb, box_area(pb));
                            if (!have) { best = v; bp = (uint32_t)p; have = true; }
                            else if (!(best != best) && ((v != v) || v > best)) { best = v; bp = (uint32_t)p; }
                        }
                    }
                }
            }
        }
        // maximum over lanes, lowest prior index among the ties (torch.max returns the first maximum; NaN wins)
        const uint32_t u = (mode != 2) ? (__float_as_uint(best) | 0x80000000u) : (have ? ord_of(best) : 0u);
        const uint32_t wmax = __reduce_max_sync(kFull, u);
        const uint32_t pmin = __reduce_min_sync(kFull, (u == wmax) ? bp : 0xffffffffu);
        if (lane == 0) {
            ws.bpi[g0 + g] = (int)pmin;
            ws.bpo[g0 + g] = ord_inv(wmax);
        }
    }
}

// grid: x = image, y = role.  y < row_ctas: row warps; then the prior tiles in REVERSE order.  Launch order is
// x-fastest, so every image's last tiles -- the large priors of the coarse pyramid levels, which intersect most GT
// and carry the longest per-warp loops -- start first and the cheap fine-level tiles fill the tail of the launch.
__global__ void __launch_bounds__(kTile) assign_match_kernel(const float4 *__restrict__ priors, int P,
                                                             const int *__restrict__ gt_off, AssignWorkspace ws, int dense,
                                                             int n_tiles)
{
    __shared__ __align__(16) ColSmem s;
    const int b = blockIdx.x;
    const int g0 = gt_off[b];
    const int G = gt_off[b + 1] - g0;
    if (G <= 0) return; // encode kernel writes zeros for this image
    const int row_ctas = (int)gridDim.y - n_tiles;
    if ((int)blockIdx.y < row_ctas) row_role(priors, P, b, (int)blockIdx.y, g0, G, ws, dense, n_tiles, row_ctas);
    else column_role(priors, P, b, (int)gridDim.y - 1 - (int)blockIdx.y, g0, G, ws, dense, s);
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
        const uint32_t bp = (uint32_t)ws.bpi[g0 + g];
        if (bp >= (uint32_t)p0 && bp < (uint32_t)(p0 + kTile)) atomicMax(&s_forced[bp - p0], g);
        if (blockIdx.x == 0) {
            if (a.out_bpi) a.out_bpi[g0 + g] = (int)bp;
            if (a.out_bpo) a.out_bpo[g0 + g] = ws.bpo[g0 + g];
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
        // 16 IEEE divisions per prior share 5 divisors: var0*w, var0*h (centre + 10 landmark coordinates), w, h
        // (size ratio) and var1 -- one refined reciprocal each, see fdiv_shared().
        const float dx = fmul(a.var0, pr.z), dy = fmul(a.var0, pr.w);
        const bool safe = divisor_safe(dx) && divisor_safe(dy) && divisor_safe(pr.z) && divisor_safe(pr.w) && divisor_safe(a.var1);
        if (safe) {
            const float rdx = rcp_refined(dx), rdy = rcp_refined(dy);
            if (a.encode_mode) {
                const float rw = rcp_refined(pr.z), rh = rcp_refined(pr.w), rv = rcp_refined(a.var1);
                loc.x = fdiv_shared(fsub(fmul(fadd(m.x, m.z), 0.5f), pr.x), dx, rdx);
                loc.y = fdiv_shared(fsub(fmul(fadd(m.y, m.w), 0.5f), pr.y), dy, rdy);
                loc.z = fdiv_shared(log_rn(fdiv_shared(fsub(m.z, m.x), pr.z, rw)), a.var1, rv);
                loc.w = fdiv_shared(log_rn(fdiv_shared(fsub(m.w, m.y), pr.w, rh)), a.var1, rv);
            } else {
                loc = m;
            }
            if (a.landm_t) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    lm[2 * k] = fdiv_shared(fsub(__ldg(r + 4 + 2 * k), pr.x), dx, rdx);
                    lm[2 * k + 1] = fdiv_shared(fsub(__ldg(r + 5 + 2 * k), pr.y), dy, rdy);
                }
            }
        } else {
            loc = a.encode_mode ? encode_box(m, pr, a.var0, a.var1) : m;
            if (a.landm_t) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    lm[2 * k] = fdiv(fsub(__ldg(r + 4 + 2 * k), pr.x), dx);
                    lm[2 * k + 1] = fdiv(fsub(__ldg(r + 5 + 2 * k), pr.y), dy);
                }
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
    JABD_REQUIRE(P <= (65535ll - kMaxRowCtas) * kTile && sumG < (1ll << 31), JABD_EINVAL,
                 "assign: at most %lld priors per image and 2^31 GT rows per call", (65535ll - kMaxRowCtas) * kTile);
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
    const unsigned n_tiles = (unsigned)((P + kTile - 1) / kTile);
    assign_prep_kernel<<<(unsigned)B + n_tiles, kTile, 0, st>>>(gt, gt_off, reinterpret_cast<const float4 *>(priors), (int)P, B, ws);
    JABD_LAUNCH_CHECK("assign_prep_kernel");
    // row warps: one per GT.  The host only knows the mean GT count, so provision ~2.5x the mean (one pass for
    // most images); warps beyond an image's G exit at once and larger images take extra passes.
    long long rows = (5 * sumG / (2 * (long long)B) + kTile / 32 - 1) / (kTile / 32);
    rows = rows < 4 ? 4 : (rows > kMaxRowCtas ? kMaxRowCtas : rows);
    const dim3 grid((unsigned)B, n_tiles + (unsigned)rows);
    assign_match_kernel<<<grid, kTile, 0, st>>>(reinterpret_cast<const float4 *>(priors), (int)P, gt_off, ws,
                                                (flags & JABD_ASSIGN_DENSE) ? 1 : 0, (int)n_tiles);
    JABD_LAUNCH_CHECK("assign_match_kernel");
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
