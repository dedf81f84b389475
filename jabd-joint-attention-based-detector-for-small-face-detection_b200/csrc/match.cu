// match.cu -- batched target assignment: priors x GT IoU, both argmaxes, force-match, SSD encode.
//
// Replaces the per-image loop of MultiBoxLoss.forward (R/nets/retinaface_training.py:197-214) and match()
// (:93-162).  Three launches per batch, all on the caller's stream; nothing is allocated, there is no host
// sync and no memset; the [G,P] IoU matrix is never stored.
//
//   assign_prep_kernel  image CTAs: GT rows -> 32-byte records (box, area, "image well formed" flag) for the match kernel and
//                       64-byte records (whole row + "fast division is exact" flag) for the encode kernel in the
//                       workspace, the image's row-argmax keys zeroed.  tile CTAs: per 256-prior tile a sanity flag,
//                       the bounding box of each of its 8 warps' priors, the column-argmax keys of the tile zeroed
//                       for every image.  one scan CTA: the work list -- every image's GT cut into segments (image, first GT,
//                       count, record offset) of <= 64 rows, or whatever the call's work-list shape says (AssignTune below:
//                       two regimes over the tiles, 16..192 rows) -- and the work-queue head.
//   assign_match_kernel persistent CTAs (4 per SM) pull work items (prior tile, GT segment) from an atomic queue,
//                       coarse-level tiles first (their large priors intersect most GT and make the longest items).
//                       One elected thread stages the next item while the eight consumer warps work on the current one:
//                       as soon as a stage of the two-deep ring is free it draws a queue ticket, reads the work-list entry
//                       and issues three bulk copies (cp.async.bulk + mbarrier: the GT segment, the tile's 256 priors,
//                       its warp boxes).  A CTA thus never owns more than two items -- with a deeper ring the queue ran
//                       empty while most CTAs still sat on a backlog nobody else could take, and the launch ended on
//                       that backlog.  Per item each warp tests 32 GT per ballot against the bounding box of its 32
//                       priors, compacts the hits into a byte list in shared memory (padded with a null GT record to a
//                       multiple of four) and visits them four per step so that their loads and divisions overlap (one
//                       prior per lane).
//       column argmax   best GT per prior (overlaps.max(0), :120): in registers over the segment (ascending GT index,
//                       strict >, so ties keep the lowest index), then one 64-bit atomicMax per prior on the packed key
//                       (ordered IoU bits << 32 | ~GT index) to combine the segments.
//       row argmax      best prior per GT (overlaps.max(1), :111): per visited GT a REDUX max over the warp's 32 IoU
//                       bit patterns; the lane(s) holding it issue a 64-bit atomicMax on the GT's packed key
//                       (ordered IoU bits << 32 | ~prior index).
//                       The largest key is the largest IoU and, among equal IoUs, the lowest index -- torch.max's
//                       first-maximum rule -- so indices match the reference bit for bit.
//   match_encode_kernel one CTA per (256-prior tile, 8 consecutive images): force-match scan (:127-130, largest j wins) while one
//                       bulk copy brings the images' 64-byte GT records into shared memory, then per image -- unpack the column
//                       key, read the matched GT's record from shared memory, threshold (:143), encode (:61-84), coalesced
//                       stores -- with the column keys four images ahead in flight.
//   jabd_assign_batches several independent batches per call, batch i on caller-owned side stream i % n_lanes (events fork the
//                       lanes from the caller's stream and join them back): one batch's staging and encode kernels run in the
//                       ramp and tail of another batch's persistent matching kernel.
//
// Why culling is exact: with well-formed inputs (finite coordinates, GT area in [0, 2^40], prior area in
// [2^-40, 2^40]) every IoU is >= +0 and a pair whose boxes do not intersect has IoU == +0 exactly, which can
// neither replace a best (strict >; both argmaxes start at value 0 / index 0 like torch.max over zeros -- a key
// that is still zero after the launch decodes to exactly that) nor tie with a positive one.  JABD_ASSIGN_DENSE
// disables culling and evaluates all P*G pairs; tests compare the two bit for bit.  If any prior, or any GT of an
// image, is malformed (negative or non-finite area) that image takes a generic dense path with torch.max's
// NaN-wins / first-index semantics.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace jabd {

constexpr int kTile = 256;       // priors per work item, one per thread
// GT per work item.  A run-time choice per call (AssignTune): small segments bound the longest per-warp dependency chain and
// the backlog a CTA can sit on when the queue runs dry -- what a launch running ALONE ends on --, large ones mean fewer items,
// i.e. fewer culling ballots, staging round trips and column-key atomics per pair -- what counts when other launches fill the
// tail anyway (jabd_assign_batches).  The work list has two regimes: the first `coarse_tiles` tiles of the processing order
// (coarse pyramid levels first) are cut into segments of `seg_a`, the remaining tiles into segments of `seg_b`.
constexpr int kSegMax = 192;     // shared-memory capacity of a stage; hit lists hold GT indices as bytes (< 256)
constexpr int kSegMin = 16;      // work-list capacity is sized for this
constexpr int kSegDefault = 64;
constexpr int kWide = 4;         // GT hits processed per step by a warp
#ifndef JABD_MATCH_CTAS
#define JABD_MATCH_CTAS 4
#endif
constexpr int kMatchCtasPerSm = JABD_MATCH_CTAS;
#ifndef JABD_KSTAGES
#define JABD_KSTAGES 2
#endif
constexpr int kStages = JABD_KSTAGES; // staged work items per CTA: consumer warps may run up to kStages-1 items apart.  Two, not more:
                                     // what a CTA has staged it must finish itself, and once the queue is empty the launch lasts
                                     // as long as the deepest backlog (4 stages + a prefetched ticket measured 23.3 us, 2 stages with
                                     // the ticket drawn only when a stage is free 20.5 us)
constexpr int kMatchThreads = kTile + 32; // 8 consumer warps + 1 producer warp

struct __align__(32) GtRec {
    float4 box;  // x1 y1 x2 y2
    float area;
    int img_ok;  // every GT of the image has a finite area in [0, 2^40]
    int pad0, pad1;
};

// What the encode kernel gathers for a prior's matched GT: the whole row as ONE aligned 64-byte record (four 16-byte
// loads instead of fifteen scalar ones) plus the verdict of the range test that lets every division of that row take
// the shared-reciprocal fast path (see match_encode_kernel).
struct __align__(64) EncRec {
    float4 box;     // x1 y1 x2 y2
    float lm[10];   // five landmark points
    float label;
    int fast_ok;    // all 14 coordinates finite with magnitude <= 2^59, x2-x1 and y2-y1 within [2^-60, 2^60]
};
static_assert(sizeof(EncRec) == 64, "EncRec is one 64-byte record");

struct AssignWorkspace {
    GtRec *gtrec;               // [sumG]
    EncRec *encrec;             // [sumG]
    unsigned long long *rowkey; // [sumG] best prior per GT (0 = value +0 at prior 0)
    unsigned long long *colkey; // [B,P]  best GT per prior before the force-match override (0 = value +0 at GT 0)
    float4 *wbox;               // [n_tiles*8] bounding box of each warp's 32 priors (point form)
    int *tile_ok;               // [n_tiles] every prior area of the tile within [2^-40, 2^40]
    int4 *segs;                 // work list, regime A: (image, first GT within the image, count, record offset)
    int4 *segs_b;               // regime B
    int *ctl;                   // [0] segments in list A, [1] work-queue head, [2] segments in list B
};

// Per-call shape of the work list (see kSegMax): results never depend on it.
struct AssignTune {
    int seg_a, seg_b;   // GT per item in the two regimes, kSegMin..kSegMax
    int coarse_pct;     // share of the tiles (processing order: coarse levels first) that use seg_a, 0..100
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
    const size_t nseg = (size_t)(B > 0 ? B : 0) + (size_t)(sumG > 0 ? sumG : 0) / kSegMin + 1;
    size_t o_rec = take(sizeof(GtRec) * ng);
    size_t o_enc = take(sizeof(EncRec) * ng);
    size_t o_row = take(sizeof(unsigned long long) * ng);
    size_t o_col = take(sizeof(unsigned long long) * (size_t)B * (size_t)P);
    size_t o_wbox = take(sizeof(float4) * nt * (kTile / 32));
    size_t o_tiles = take(sizeof(int) * nt);
    size_t o_segs = take(sizeof(int4) * nseg);
    size_t o_segs_b = take(sizeof(int4) * nseg);
    size_t o_ctl = take(sizeof(int) * 4);
    if (w) {
        w->gtrec = reinterpret_cast<GtRec *>(base + o_rec);
        w->encrec = reinterpret_cast<EncRec *>(base + o_enc);
        w->rowkey = reinterpret_cast<unsigned long long *>(base + o_row);
        w->colkey = reinterpret_cast<unsigned long long *>(base + o_col);
        w->wbox = reinterpret_cast<float4 *>(base + o_wbox);
        w->tile_ok = reinterpret_cast<int *>(base + o_tiles);
        w->segs = reinterpret_cast<int4 *>(base + o_segs);
        w->segs_b = reinterpret_cast<int4 *>(base + o_segs_b);
        w->ctl = reinterpret_cast<int *>(base + o_ctl);
    }
    return off;
}

// -------------------------------------------------------------------------------------------------------
// blockIdx.x < B: image role; < B + n_tiles: tile role; == B + n_tiles: scan role.
__global__ void __launch_bounds__(kTile) assign_prep_kernel(const float *__restrict__ gt, const int *__restrict__ gt_off,
                                                            const float4 *__restrict__ priors, int P, int B, int n_tiles,
                                                            int queue_head, AssignWorkspace ws, int seg_a, int seg_b)
{
    __shared__ int s_scan[kTile / 32];
    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = tid >> 5;
    if ((int)blockIdx.x < B) { // ---- image role
        const int b = blockIdx.x;
        const int g0 = gt_off[b];
        const int G = gt_off[b + 1] - g0;
        // one pass over the rows (one load round trip): everything but the "whole image well formed" flag, which needs the
        // CTA's vote, is written at once; the flag follows after the barrier
        int ok = 1;
        float area0 = 0.0f; // area of this thread's first row (G <= kTile: its only one)
        for (int g = tid; g < G; g += kTile) {
            const float *r = gt + (size_t)(g0 + g) * JABD_GT_ROW;
            float v[JABD_GT_ROW];
#pragma unroll
            for (int k = 0; k < JABD_GT_ROW; ++k) v[k] = __ldg(r + k);
            const float4 box = make_float4(v[0], v[1], v[2], v[3]);
            const float area = box_area(box);
            ok &= (area >= 0.0f && area <= 0x1p40f) ? 1 : 0; // false for NaN / inf coordinates too
            if (g == tid) area0 = area;
            ws.gtrec[g0 + g].box = box;
            ws.rowkey[g0 + g] = 0ull;
            float mx = 0.0f;
            bool finite = true;
#pragma unroll
            for (int k = 0; k < JABD_GT_ROW - 1; ++k) {
                mx = fmaxf(mx, fabsf(v[k]));       // fmaxf drops a NaN operand, hence the separate test
                finite = finite && (v[k] == v[k]);
            }
            const float nw = fsub(v[2], v[0]), nh = fsub(v[3], v[1]);
            const int fast_ok = (finite && mx <= 0x1p59f && nw >= 0x1p-60f && nw <= 0x1p60f && nh >= 0x1p-60f && nh <= 0x1p60f) ? 1 : 0;
            float4 *e = reinterpret_cast<float4 *>(ws.encrec + g0 + g);
            e[0] = box;
            e[1] = make_float4(v[4], v[5], v[6], v[7]);
            e[2] = make_float4(v[8], v[9], v[10], v[11]);
            e[3] = make_float4(v[12], v[13], v[14], __int_as_float(fast_ok));
        }
        ok = __syncthreads_and(ok);
        for (int g = tid; g < G; g += kTile) {
            float area = area0;
            if (g != tid) area = box_area(ws.gtrec[g0 + g].box); // this thread's own store
            // second half of the record: area, img_ok, padding
            reinterpret_cast<float4 *>(ws.gtrec + g0 + g)[1] = make_float4(area, __int_as_float(ok), 0.0f, 0.0f);
        }
        return;
    }
    const int t = (int)blockIdx.x - B;
    if (t < n_tiles) { // ---- tile role
        const int p = t * kTile + tid;
        const bool valid = p < P;
        int ok = 1;
        float x1 = CUDART_INF_F, y1 = CUDART_INF_F, x2 = -CUDART_INF_F, y2 = -CUDART_INF_F;
        if (valid) {
            const float4 pb = to_point_form(__ldg(priors + p));
            const float area = box_area(pb);
            ok = (area >= 0x1p-40f && area <= 0x1p40f) ? 1 : 0;
            x1 = pb.x; y1 = pb.y; x2 = pb.z; y2 = pb.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x1 = fminf(x1, __shfl_xor_sync(kFull, x1, o));
            y1 = fminf(y1, __shfl_xor_sync(kFull, y1, o));
            x2 = fmaxf(x2, __shfl_xor_sync(kFull, x2, o));
            y2 = fmaxf(y2, __shfl_xor_sync(kFull, y2, o));
        }
        if (lane == 0) ws.wbox[t * (kTile / 32) + warp] = make_float4(x1, y1, x2, y2);
        ok = __syncthreads_and(ok);
        if (tid == 0) ws.tile_ok[t] = ok;
        if (valid)
            for (int b = 0; b < B; ++b) ws.colkey[(size_t)b * P + p] = 0ull;
        return;
    }
    // ---- scan role: the two segment lists in image order (list B only if some tile uses it: seg_b > 0)
    for (int pass = 0; pass < 2; ++pass) {
        const int seg = pass == 0 ? seg_a : seg_b;
        int4 *list = pass == 0 ? ws.segs : ws.segs_b;
        int running = 0;
        for (int base = 0; seg > 0 && base < B; base += kTile) {
            const int b = base + tid;
            int g0 = 0, G = 0;
            if (b < B) { g0 = gt_off[b]; G = gt_off[b + 1] - g0; }
            const int c = G > 0 ? (G + seg - 1) / seg : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if ((int)lane >= o) incl += v;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kTile / 32; ++w) {
                const int v = s_scan[w];
                if (w < warp) before += v;
                total += v;
            }
            const int first = running + before + incl - c;
            for (int k = 0; k < c; ++k) {
                const int c0 = k * seg;
                list[first + k] = make_int4(b, c0, (G - c0) < seg ? (G - c0) : seg, g0 + c0);
            }
            running += total;
            __syncthreads();
        }
        if (tid == 0) ws.ctl[pass == 0 ? 0 : 2] = running;
    }
    if (tid == 0) ws.ctl[1] = queue_head;
}

// -------------------------------------------------------------------------------------------------------
// IoU of a well-formed pair (union in [2^-40, 2^42], inter >= 0): the bits of __fdiv_rn(inter, union) unless the
// intersection is a non-zero value below 2^-60, which sets `slow` instead (the caller then redoes its whole step
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

// atomicMax on a 64-bit key assembled from two 32-bit halves, predicated (no branch, no return value): RED.MAX.64
__device__ __forceinline__ void red_max_key(unsigned long long *addr, uint32_t hi, uint32_t lo, bool pred)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b64 k;\n\t"
        "setp.ne.u32 q, %3, 0;\n\t"
        "mov.b64 k, {%2, %1};\n\t"
        "@q red.global.max.u64 [%0], k;\n\t}"
        :
        : "l"(addr), "r"(hi), "r"(lo), "r"((uint32_t)pred)
        : "memory");
}

struct ItemMeta {
    int tile;   // < 0: queue drained
    int image;
    int c0;     // first GT of the segment within the image
    int n;      // GT in the segment
    int rec0;   // offset of the segment in gtrec / rowkey
    int pad0, pad1, pad2;
};

struct MatchSmem {
    GtRec gt[kStages][kSegMax + 1];   // bulk-copy destinations (kStages x 6 KB); record kSegMax of every stage is the null GT
                                      // (an inverted box far away: zero intersection with everything), never overwritten
    float4 pri[kStages][kTile];       //                        (kStages x 4 KB)
    float4 wbox[kStages][kTile / 32]; //                        (kStages x 128 B)
    ItemMeta meta[kStages];
    uint64_t full[kStages];           // producer -> consumers: item staged (transaction count)
    uint64_t empty[kStages];          // consumers -> producer: one arrival per consumer warp
    uint32_t hits[kTile / 32][12];    // per consumer warp: the GT of the current 32-GT group that hit its bounding box, one
                                      // byte each, padded with the null GT to a multiple of four (<= 36 bytes used)
};
constexpr int kNullGt = kSegMax;      // index of the null GT record of a stage
static_assert(kSegMax + 1 <= 256, "hit lists hold GT indices as bytes");

// One GT segment against the warp's 32 priors.  MODE 0: culled (ballot of 32 GT against the warp's bounding box),
// MODE 1: dense, MODE 2: dense with torch.max's NaN / ordering semantics (malformed input).
//   rowkeys: row-argmax keys of this segment's GT;  p: this lane's prior index;  best/bidx: column argmax over the segment
//   (bidx relative to the segment).
template <int MODE>
__device__ __forceinline__ void consume_segment(const GtRec *__restrict__ rec, int n, unsigned long long *rowkeys, float4 pb,
                                                float area_p, float4 wbox, int p, bool valid, float &best, int &bidx,
                                                bool &have_best, uint32_t *hits)
{
    const unsigned lane = lane_id();
    if (MODE == 2) {
        for (int j = 0; j < n; ++j) {
            const float4 a = rec[j].box;
            const float v = iou_ref(a, rec[j].area, pb, area_p);
            uint32_t hi = 0, lo = 0;
            if (valid) {
                if (!have_best) { best = v; bidx = j; have_best = true; }
                else if (!(best != best) && ((v != v) || v > best)) { best = v; bidx = j; }
                hi = ord_of(v);
                lo = 0xffffffffu - (uint32_t)p;
            }
            const uint32_t whi = __reduce_max_sync(kFull, hi);
            const uint32_t wlo = __reduce_max_sync(kFull, hi == whi ? lo : 0u);
            if (lane == 0 && (whi | wlo)) atomicMax(rowkeys + j, ((unsigned long long)whi << 32) | wlo);
        }
        return;
    }
    const uint32_t notp = 0xffffffffu - (uint32_t)p;
    uint8_t *list = reinterpret_cast<uint8_t *>(hits);
    for (int base = 0; base < n; base += 32) {
        const int e = base + (int)lane;
        bool hit = e < n;
        if (MODE == 0 && hit) {
            const float4 a = rec[e].box;
            const float w = fsub(fminf(a.z, wbox.z), fmaxf(a.x, wbox.x));
            const float h = fsub(fminf(a.w, wbox.w), fmaxf(a.y, wbox.y));
            hit = (w > 0.0f) && (h > 0.0f);
        }
        const unsigned m = __ballot_sync(kFull, hit);
        if (m == 0u) continue;
        // the group's hits as a byte list in shared memory, padded with the null GT to a multiple of kWide: the steps below read
        // four indices with one broadcast load and need neither bit scans nor a "repeat the last hit" fix-up
        const int cnt = __popc(m);
        if (hit) list[__popc(m & lanemask_lt())] = (uint8_t)e;
        if (lane < (unsigned)(kWide - 1)) list[cnt + (int)lane] = (uint8_t)kNullGt;
        __syncwarp();
        for (int i = 0; i < cnt; i += kWide) {
            const uint32_t four = hits[i >> 2];
            int j[kWide];
            float4 a[kWide];
            float aa[kWide];
            float v[kWide];
#pragma unroll
            for (int k = 0; k < kWide; ++k) {
                j[k] = (int)__byte_perm(four, 0u, 0x4440u + (unsigned)k);   // byte k, zero extended: one PRMT
                a[k] = rec[j[k]].box;
                aa[k] = rec[j[k]].area;
            }
            bool slow = false;
#pragma unroll
            for (int k = 0; k < kWide; ++k) v[k] = iou_wellformed(a[k], aa[k], pb, area_p, slow);
            if (slow) { // a sliver intersection below 2^-60 somewhere in this step: generic IEEE divide
#pragma unroll
                for (int k = 0; k < kWide; ++k) v[k] = iou_ref(a[k], aa[k], pb, area_p);
            }
#pragma unroll
            for (int k = 0; k < kWide; ++k) {
                // column argmax: ascending GT index, strict > : ties keep the lowest index (the null GT scores +0: never taken)
                if (v[k] > best) { best = v[k]; bidx = j[k]; }
                // row argmax: the warp's largest IoU for this GT; every lane that holds it (ties are rare) pushes its
                // own key, the 64-bit max keeps the lowest prior index.  A well-formed pair has v >= +0, so the bit patterns
                // order like the values; a warp whose best is +0 (no lane intersects this GT, or the null GT) pushes nothing.
                const uint32_t bits = __float_as_uint(v[k]);
                const uint32_t wmax = __reduce_max_sync(kFull, bits);
                if (wmax != 0u) red_max_key(rowkeys + j[k], bits | 0x80000000u, notp, bits == wmax);
            }
        }
        __syncwarp(); // every lane has read the list before the next group overwrites it
    }
}

__global__ void __launch_bounds__(kMatchThreads, kMatchCtasPerSm) assign_match_kernel(const float4 *__restrict__ priors, int P,
                                                                                      AssignWorkspace ws, int dense,
                                                                                      int n_tiles, int coarse_tiles)
{
    __shared__ __align__(128) MatchSmem s;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    // items 0 .. items_a-1: the first coarse_tiles tiles of the processing order x list A; then the other tiles x list B
    const int n_seg_a = ws.ctl[0], n_seg_b = ws.ctl[2];
    const long long items_a = (long long)n_seg_a * coarse_tiles;
    const long long n_items = items_a + (long long)n_seg_b * (n_tiles - coarse_tiles);

    int all_ok = 1;
    for (int t = tid; t < n_tiles; t += kMatchThreads) all_ok &= ws.tile_ok[t];
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], kTile / 32); }
    }
    if (tid < kStages) {
        GtRec null_rec;
        null_rec.box = make_float4(1e30f, 1e30f, -1e30f, -1e30f);   // min(x2) - max(x1) is hugely negative against any box: inter = +0
        null_rec.area = 1.0f;
        null_rec.img_ok = 1;
        null_rec.pad0 = null_rec.pad1 = 0;
        s.gt[tid][kNullGt] = null_rec;
    }
    all_ok = __syncthreads_and(all_ok);

    if (warp == kTile / 32) {
        // ---- producer warp (one lane)
        if (tid != kTile) return;
        long long item = blockIdx.x;                      // the first item is static, the rest come from the queue
        for (int k = 0;; ++k) {
            const int slot = k % kStages;
            if (k >= kStages) mbar_wait_relaxed(&s.empty[slot], (uint32_t)((k / kStages - 1) & 1));
            // a ticket is drawn only when there is a free stage for it: a CTA never owns more than kStages items, so the tail of
            // the launch (queue empty, CTAs finishing what they hold) is at most that deep
            if (k > 0) item = atomicAdd(&ws.ctl[1], 1);
            ItemMeta m;
            m.tile = -1; m.image = m.c0 = m.n = m.rec0 = m.pad0 = m.pad1 = m.pad2 = 0;
            if (item >= n_items) { // queue drained: publish the sentinel and stop
                s.meta[slot] = m;
                mbar_arrive(&s.full[slot]);
                return;
            }
            int4 seg;
            int tile_ord;
            if (item < items_a) {
                seg = ws.segs[item % n_seg_a];
                tile_ord = (int)(item / n_seg_a);
            } else {
                const long long it = item - items_a;
                seg = ws.segs_b[it % n_seg_b];
                tile_ord = coarse_tiles + (int)(it / n_seg_b);
            }
            const int tile = n_tiles - 1 - tile_ord; // coarse pyramid levels (last tiles) first
            const int np = (P - tile * kTile) < kTile ? (P - tile * kTile) : kTile;
            m.tile = tile; m.image = seg.x; m.c0 = seg.y; m.n = seg.z; m.rec0 = seg.w;
            s.meta[slot] = m;
            mbar_arrive_expect_tx(&s.full[slot], (uint32_t)(seg.z * 32 + np * 16 + (kTile / 32) * 16));
            bulk_g2s(s.gt[slot], ws.gtrec + seg.w, (uint32_t)seg.z * 32u, &s.full[slot]);
            bulk_g2s(s.pri[slot], priors + (size_t)tile * kTile, (uint32_t)np * 16u, &s.full[slot]);
            bulk_g2s(s.wbox[slot], ws.wbox + (size_t)tile * (kTile / 32), (kTile / 32) * 16u, &s.full[slot]);
        }
    }

    // ---- consumer warps: no CTA-wide barrier; a warp releases a stage as soon as it is done with it
    const unsigned lane = lane_id();
    for (int k = 0;; ++k) {
        const int slot = k % kStages;
        mbar_wait(&s.full[slot], (uint32_t)((k / kStages) & 1));
        const ItemMeta meta = s.meta[slot];
        if (meta.tile < 0) break;
        const int n = meta.n;
        const int p = meta.tile * kTile + tid;
        const bool valid = p < P;
        // out-of-range threads carry a box that never has a positive intersection
        float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
        float area_p = 1.0f;
        if (valid) {
            pb = to_point_form(s.pri[slot][tid]);
            area_p = box_area(pb);
        }
        const float4 wbox = s.wbox[slot][warp];
        const GtRec *rec = s.gt[slot];
        unsigned long long *rowkeys = ws.rowkey + meta.rec0;
        const int mode = (all_ok && rec[0].img_ok) ? (dense ? 1 : 0) : 2; // uniform over the image
        float best = 0.0f;
        int bidx = 0;
        bool have_best = false;
        if (mode == 0) consume_segment<0>(rec, n, rowkeys, pb, area_p, wbox, p, valid, best, bidx, have_best, s.hits[warp]);
        else if (mode == 1) consume_segment<1>(rec, n, rowkeys, pb, area_p, wbox, p, valid, best, bidx, have_best, s.hits[warp]);
        else consume_segment<2>(rec, n, rowkeys, pb, area_p, wbox, p, valid, best, bidx, have_best, s.hits[warp]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[slot]); // this warp no longer reads the stage
        // combine the segments of the image: max key = largest IoU, then lowest GT index
        unsigned long long *ck = ws.colkey + (size_t)meta.image * P + p;
        const uint32_t lo = 0xffffffffu - (uint32_t)(meta.c0 + bidx);
        if (mode != 2) red_max_key(ck, __float_as_uint(best) | 0x80000000u, lo, valid && best > 0.0f);
        else if (valid && have_best) atomicMax(ck, make_key(ord_of(best), (uint32_t)(meta.c0 + bidx)));
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

// One CTA = kEncThreads consecutive priors (one per thread) x kEncImages consecutive images.  What depends on the prior alone
// (five refined reciprocals, the range test) is computed once per thread; the images are then encoded in a software pipeline:
// while image i is being encoded the matched GT record of image i+1 is already in flight (one aligned 64-byte EncRec, four
// 16-byte loads) and the column keys of images i+2, i+3 are being fetched.  The force-match scan of all the CTA's images runs
// once, before the loop, between the CTA's only two barriers.
//
// Divisions.  16 IEEE quotients per prior and image share 5 divisors (var0*w, var0*h for the centre and the ten landmark
// coordinates; w, h for the size ratio; var1) -- one refined reciprocal each (fdiv_fast).  The fast quotient is the bits of
// div.rn when the divisor's magnitude lies in [2^-60, 2^60] and the numerator is zero or lies in that range.  That is
// established ONCE per prior and once per GT row instead of per quotient:
//   * prior:  |cx|, |cy|, w, h, var0*w, var0*h within [2^-36, 2^59] (prior_ok), var1 within [2^-60, 2^60];
//   * GT row (EncRec.fast_ok, assign_prep_kernel): 14 finite coordinates of magnitude <= 2^59, x2-x1 and y2-y1 in [2^-60, 2^60].
//   Then every numerator  t - c  (t a GT coordinate or box centre, |t| <= 2^59; c = cx or cy, 2^-36 <= |c| <= 2^59) is at most
//   2^60, and if it is not zero it is at least 2^-60: either |t| < |c|/2 and the difference exceeds |c|/2 >= 2^-37, or both are
//   multiples of ulp(2^-37) = 2^-60 and so is their exact difference.  t - c is never -0 (c != 0).  The size ratios are normal
//   positive numbers, their logarithms are zero or of magnitude in [2^-25, 2^7].
// A row or prior that fails takes the compiler's generic divide (bit-identical results, tests force both paths).
#ifndef JABD_ENC_THREADS
#define JABD_ENC_THREADS 256
#endif
#ifndef JABD_ENC_IMAGES
#define JABD_ENC_IMAGES 8
#endif
#ifndef JABD_ENC_MINB
#define JABD_ENC_MINB 3
#endif
#ifndef JABD_ENC_RECCAP
#define JABD_ENC_RECCAP 896
#endif
constexpr int kEncThreads = JABD_ENC_THREADS;
constexpr int kEncImages = JABD_ENC_IMAGES;
constexpr int kEncScan = 4;                   // row keys per thread in flight during the force-match scan
constexpr int kEncRecCap = JABD_ENC_RECCAP;   // GT records of a CTA's images staged in shared memory (56 KB: three CTAs per SM);
                                              // what does not fit is read from HBM

struct EncRow {   // the GT a prior is matched with in one image
    float4 e0, e1, e2, e3;   // its EncRec
    int idx;
    float ov;
};

// Per-thread constants of the encode loop.
struct EncPrior {
    float4 pr;
    float dx, dy, rdx, rdy, rw, rh, rv;
    bool ok;
};

struct EncOut {   // one prior's targets for one image
    float4 loc;
    long long conf;
    float lm[10];
    int idx;
    float ov;
};

// kLandm: landm_t requested; kEncode: SSD encode (else raw matched boxes, match_iou); kExtra: label_mode / the optional
// best_truth_* outputs are looked at (the MultiBoxLoss path compiles without them).
template <bool kLandm, bool kEncode, bool kExtra>
__device__ __forceinline__ EncOut encode_compute(const EncodeArgs &a, const EncPrior &q, const EncRow &r, bool has_gt)
{
    EncOut o;
    const float4 m = r.e0;
    float c = r.e3.z;
    if (kExtra && a.label_mode) c = fadd(c, 1.0f); // R/utils/box_utils.py:315
    if (r.ov < a.threshold) c = 0.0f;              // :143
    o.conf = (long long)c;                         // float -> int64 store truncates
    o.idx = r.idx;
    o.ov = r.ov;
    const bool fast = q.ok && __float_as_int(r.e3.w) != 0;
    o.loc = m;
    float nl[10];
    if (kLandm) {
        const float t[10] = {r.e1.x, r.e1.y, r.e1.z, r.e1.w, r.e2.x, r.e2.y, r.e2.z, r.e2.w, r.e3.x, r.e3.y};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            nl[2 * k] = fsub(t[2 * k], q.pr.x);
            nl[2 * k + 1] = fsub(t[2 * k + 1], q.pr.y);
            o.lm[2 * k] = fdiv_fast(nl[2 * k], q.dx, q.rdx);
            o.lm[2 * k + 1] = fdiv_fast(nl[2 * k + 1], q.dy, q.rdy);
        }
    }
    if (kEncode) {
        const float ncx = fsub(fmul(fadd(m.x, m.z), 0.5f), q.pr.x), ncy = fsub(fmul(fadd(m.y, m.w), 0.5f), q.pr.y);
        const float nw = fsub(m.z, m.x), nh = fsub(m.w, m.y);
        o.loc.x = fdiv_fast(ncx, q.dx, q.rdx);
        o.loc.y = fdiv_fast(ncy, q.dy, q.rdy);
        const float lw = log_f32(fdiv_fast(nw, q.pr.z, q.rw)), lh = log_f32(fdiv_fast(nh, q.pr.w, q.rh));
        o.loc.z = fdiv_fast(lw, a.var1, q.rv);
        o.loc.w = fdiv_fast(lh, a.var1, q.rv);
    }
    if (!fast) {
        if (kEncode) o.loc = encode_box(m, q.pr, a.var0, a.var1);
        if (kLandm) {
#pragma unroll
            for (int k = 0; k < 10; ++k) o.lm[k] = fdiv(nl[k], (k & 1) ? q.dy : q.dx);
        }
    }
    if (!has_gt) { // an image without GT: all-zero targets (the reference's freshly allocated rows stay untouched)
        o.loc = make_float4(0.f, 0.f, 0.f, 0.f);
        o.conf = 0;
        o.idx = 0;
        o.ov = 0.0f;
        if (kLandm) {
#pragma unroll
            for (int k = 0; k < 10; ++k) o.lm[k] = 0.0f;
        }
    }
    return o;
}

template <bool kLandm, bool kExtra>
__device__ __forceinline__ void encode_store(const EncodeArgs &a, const EncOut &o, size_t row, size_t wrow, bool valid, bool vec_ok,
                                             int n_valid, float *sl, unsigned lane)
{
    if (valid) {
        a.loc_t[row] = o.loc;
        a.conf_t[row] = o.conf;
        if (kExtra) {
            if (a.out_bti) a.out_bti[row] = o.idx;
            if (a.out_bto) a.out_bto[row] = o.ov;
        }
    }
    if (kLandm) {
        // the warp's [32,10] rows through shared memory so that the global stores are contiguous 16-byte vectors
        __syncwarp(); // the previous image's rows have been read
#pragma unroll
        for (int k = 0; k < 10; k += 2) *reinterpret_cast<float2 *>(sl + lane * 10 + k) = make_float2(o.lm[k], o.lm[k + 1]);
        __syncwarp();
        float *dst = a.landm_t + wrow * 10;
        if (vec_ok && (wrow & 1) == 0) { // 80 aligned vectors, 2.5 per lane
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            const float4 *s4 = reinterpret_cast<const float4 *>(sl);
            d4[lane] = s4[lane];
            d4[lane + 32] = s4[lane + 32];
            if (lane < 16) d4[lane + 64] = s4[lane + 64];
        } else {
            for (int j = (int)lane; j < n_valid * 10; j += 32) dst[j] = sl[j];
        }
    }
}

template <bool kLandm, bool kEncode, bool kExtra>
__global__ void __launch_bounds__(kEncThreads, JABD_ENC_MINB) match_encode_kernel(EncodeArgs a, AssignWorkspace ws, int B, int rec_cap)
{
    // the 64-byte records of the CTA's images (contiguous in the workspace), brought in by ONE bulk copy while the force-match
    // scan runs: the gather of a prior's matched record is then a shared-memory read instead of a dependent L2 round trip per image
    extern __shared__ __align__(128) unsigned char enc_dyn[];
    EncRec *s_rec = reinterpret_cast<EncRec *>(enc_dyn);
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_off[kEncImages + 1];
    __shared__ int s_forced[kEncImages][kEncThreads];
    __shared__ __align__(16) float s_lm[kLandm ? kEncThreads * 10 : 4];   // 320 per warp

    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int P = a.P;
    const int p0 = blockIdx.x * kEncThreads;
    const int wp0 = p0 + (tid & ~31);                  // first prior of this warp
    const int p = p0 + tid;
    const bool valid = p < P;
    const int b0 = blockIdx.y * kEncImages;
    const int nb = (B - b0) < kEncImages ? (B - b0) : kEncImages;

    if (tid <= kEncImages) s_off[tid] = a.gt_off[b0 + (tid < nb ? tid : nb)];
    if (tid == 0) mbar_init(&s_bar, 1);
    // column keys (0: no positive IoU, i.e. value +0 at GT 0) of the first two images
    const unsigned long long *ckp = ws.colkey + (size_t)b0 * P + (valid ? p : 0);
    unsigned long long ck0 = ckp[0];
    unsigned long long ck1 = nb > 1 ? ckp[P] : 0ull;
#pragma unroll
    for (int i = 0; i < kEncImages; ++i) s_forced[i][tid] = -1;
    EncPrior q;
    q.pr = __ldg(a.priors + (valid ? p : 0));
    __syncthreads();
    const int rec0 = s_off[0];
    // staged: what fits, and no more records than half the CTA's lookups (a single 2048^2 image with 1,500 faces is gathered
    // from HBM as before: 256 lookups do not pay for 96 KB of staging per CTA)
    int n_staged = s_off[kEncImages] - rec0;
    n_staged = n_staged < rec_cap ? n_staged : rec_cap;
    n_staged = n_staged < nb * (kEncThreads / 2) ? n_staged : nb * (kEncThreads / 2);
    if (tid == 0 && n_staged > 0) {
        mbar_arrive_expect_tx(&s_bar, (uint32_t)n_staged * 64u);
        bulk_g2s(s_rec, ws.encrec + rec0, (uint32_t)n_staged * 64u, &s_bar);
    }

    // force-match: best_truth_idx[best_prior_idx[j]] = j for j ascending -> the largest j wins (:129-130).  The row keys of
    // the CTA's images are contiguous; every thread takes kEncScan of them per round, all loads of a round in flight together.
    // (Encoding speculatively by the column argmax while the scan is in flight, and the few forced priors again at the end,
    // was measured: the extra barriers and the second pass cost more than the round trip they hide, 11.5 vs 10.9 us.)
    {
        int off[kEncImages + 1];
#pragma unroll
        for (int i = 0; i <= kEncImages; ++i) off[i] = s_off[i];
        const int gtot = off[kEncImages] - off[0];
        for (int base = 0; base < gtot; base += kEncScan * kEncThreads) {
            unsigned long long rk[kEncScan];
#pragma unroll
            for (int k = 0; k < kEncScan; ++k) {
                const int g = base + k * kEncThreads + tid;
                rk[k] = (g < gtot) ? ws.rowkey[off[0] + g] : 0ull;
            }
#pragma unroll
            for (int k = 0; k < kEncScan; ++k) {
                const int g = off[0] + base + k * kEncThreads + tid;
                if (g < off[kEncImages]) {
                    int img = 0;
#pragma unroll
                    for (int i = 1; i < kEncImages; ++i) img += (g >= off[i]) ? 1 : 0;
                    const unsigned long long key = rk[k];                             // 0: no positive IoU, i.e. value +0 at prior 0
                    const uint32_t bp = key ? key_idx(key) : 0u;
                    if (bp - (uint32_t)p0 < (uint32_t)kEncThreads) atomicMax(&s_forced[img][bp - (uint32_t)p0], g - s_off[img]);
                    if (kExtra && p0 == 0) {
                        if (a.out_bpi) a.out_bpi[g] = (int)bp;
                        if (a.out_bpo) a.out_bpo[g] = key ? ord_inv(key_ord(key)) : 0.0f;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (wp0 >= P) return;

    // the GT row a prior is matched with in image i: the forced one if any, else its column argmax
    auto fetch = [&](int i, unsigned long long key) {
        EncRow r;
        r.idx = 0;
        r.ov = 0.0f;
        if (key) { r.idx = (int)key_idx(key); r.ov = ord_inv(key_ord(key)); }
        const int f = s_forced[i][tid];
        if (f >= 0) { r.idx = f; r.ov = 2.0f; }    // :127
        // (an image without GT: the record belongs to the next image or is the workspace's spare one -- loaded, never used)
        const int rel = s_off[i] + r.idx - rec0;
        if (rel < n_staged) {
            const float4 *e = reinterpret_cast<const float4 *>(s_rec + rel);
            r.e0 = e[0]; r.e1 = e[1]; r.e2 = e[2]; r.e3 = e[3];
        } else {
            const float4 *e = reinterpret_cast<const float4 *>(ws.encrec + rec0 + rel);
            r.e0 = __ldg(e); r.e1 = __ldg(e + 1); r.e2 = __ldg(e + 2); r.e3 = __ldg(e + 3);
        }
        return r;
    };
    unsigned long long cka = nb > 2 ? ckp[(size_t)2 * P] : 0ull, ckb = nb > 3 ? ckp[(size_t)3 * P] : 0ull; // images i+2, i+3

    auto prior_consts = [&](EncPrior &c) {
        c.dx = fmul(a.var0, c.pr.z);
        c.dy = fmul(a.var0, c.pr.w);
        const float4 pr = c.pr;
        const float lo = fminf(fminf(fminf(fabsf(pr.x), fabsf(pr.y)), fminf(fabsf(pr.z), fabsf(pr.w))), fminf(fabsf(c.dx), fabsf(c.dy)));
        const float hi = fmaxf(fmaxf(fmaxf(fabsf(pr.x), fabsf(pr.y)), fmaxf(fabsf(pr.z), fabsf(pr.w))), fmaxf(fabsf(c.dx), fabsf(c.dy)));
        // fminf/fmaxf drop a NaN operand, hence the explicit NaN tests
        c.ok = lo >= 0x1p-36f && hi <= 0x1p59f && (pr.x == pr.x) && (pr.y == pr.y) && (pr.z == pr.z) && (pr.w == pr.w) && mag_safe(a.var1);
        c.rdx = rcp_refined(c.dx);
        c.rdy = rcp_refined(c.dy);
        c.rw = rcp_refined(c.pr.z);
        c.rh = rcp_refined(c.pr.w);
        c.rv = rcp_refined(a.var1);
    };
    prior_consts(q);
    float *sl = s_lm + (kLandm ? (tid & ~31) * 10 : 0);
    const int n_valid = (P - wp0) < 32 ? (P - wp0) : 32;
    const bool vec_ok = n_valid == 32 && (reinterpret_cast<uintptr_t>(a.landm_t) & 15u) == 0;

    if (n_staged > 0) mbar_wait(&s_bar, 0u);   // the records have landed (issued before the scan: normally long since)
    // Two images per trip; the column keys run four images ahead (plain coalesced loads, the only global reads of the loop).
    size_t row = (size_t)b0 * P + p, wrow = (size_t)b0 * P + wp0;
#pragma unroll 1
    for (int i = 0; i < nb; i += 2) {
        const unsigned long long ck4 = (i + 4 < nb) ? ckp[(size_t)(i + 4) * P] : 0ull;
        const unsigned long long ck5 = (i + 5 < nb) ? ckp[(size_t)(i + 5) * P] : 0ull;
        const EncOut o0 = encode_compute<kLandm, kEncode, kExtra>(a, q, fetch(i, ck0), s_off[i + 1] > s_off[i]);
        encode_store<kLandm, kExtra>(a, o0, row, wrow, valid, vec_ok, n_valid, sl, lane);
        if (i + 1 >= nb) break;
        row += P; wrow += P;
        const EncOut o1 = encode_compute<kLandm, kEncode, kExtra>(a, q, fetch(i + 1, ck1), s_off[i + 2] > s_off[i + 1]);
        encode_store<kLandm, kExtra>(a, o1, row, wrow, valid, vec_ok, n_valid, sl, lane);
        row += P; wrow += P;
        ck0 = cka; ck1 = ckb;
        cka = ck4; ckb = ck5;
    }
}

// -------------------------------------------------------------------------------------------------------
static int check_assign_common(const float *priors, int64_t P, const float *gt, const int *gt_off, int B, int64_t sumG,
                               void *workspace, size_t workspace_bytes)
{
    JABD_REQUIRE(B >= 0 && P >= 0 && sumG >= 0, JABD_EINVAL, "assign: negative size (B=%d P=%lld sumG=%lld)", B,
                 (long long)P, (long long)sumG);
    JABD_REQUIRE(B <= 65535, JABD_EINVAL, "assign: B=%d exceeds 65535 images per call", B);
    JABD_REQUIRE(P <= 65535ll * kTile && sumG < (1ll << 31), JABD_EINVAL,
                 "assign: at most %lld priors per image and 2^31 GT rows per call", 65535ll * kTile);
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

// cudaFuncAttributeMaxDynamicSharedMemorySize for one instantiation of the encode kernel, once per device (a memo of a device
// property, like detect.cu's: a race only repeats an idempotent runtime call)
template <typename K>
static int enc_smem_optin(K kernel, int variant)
{
    static std::atomic<bool> done[8][64];
    int dev = 0;
    JABD_CUDA(cudaGetDevice(&dev));
    const bool memo = dev >= 0 && dev < 64 && variant >= 0 && variant < 8;
    if (memo && done[variant][dev].load(std::memory_order_acquire)) return JABD_OK;
    JABD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kEncRecCap * sizeof(EncRec))));
    if (memo) done[variant][dev].store(true, std::memory_order_release);
    return JABD_OK;
}

// JABD_ASSIGN_TUNE(seg_a, seg_b, coarse_pct) in bits 8..30 of a call's flags; all-zero bits: the default shape of a call
// that runs alone (see kSegMax).
static int assign_tune_of(int flags, AssignTune *t)
{
    const int a = (flags >> 8) & 255, b = (flags >> 16) & 255, pct = (flags >> 24) & 127;
    if (a == 0 && b == 0 && pct == 0) {
        t->seg_a = kSegDefault;
        t->seg_b = kSegDefault;
        t->coarse_pct = 100;
        return JABD_OK;
    }
    JABD_REQUIRE(a >= kSegMin && a <= kSegMax && b >= kSegMin && b <= kSegMax && pct <= 100, JABD_EINVAL,
                 "assign: JABD_ASSIGN_TUNE(seg_a, seg_b, coarse_pct) needs %d <= seg <= %d and 0 <= coarse_pct <= 100", kSegMin, kSegMax);
    t->seg_a = a;
    t->seg_b = b;
    t->coarse_pct = pct;
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
    if (rc != JABD_OK) return rc;
    AssignTune tune;
    rc = assign_tune_of(flags, &tune);
    if (rc != JABD_OK || B == 0 || P == 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AssignWorkspace ws;
    assign_ws_layout(B, P, sumG, &ws, static_cast<char *>(workspace));
    const unsigned n_tiles = (unsigned)((P + kTile - 1) / kTile);
    int dev = 0, sms = 0;
    JABD_CUDA(cudaGetDevice(&dev));
    JABD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // persistent grid: one CTA per resident slot, never more than there can be work items
    const int coarse_tiles = (int)(((long long)n_tiles * tune.coarse_pct + 99) / 100);
    const int seg_small = coarse_tiles == 0 ? tune.seg_b
                                            : (coarse_tiles == (int)n_tiles || tune.seg_a < tune.seg_b ? tune.seg_a : tune.seg_b);
    const long long max_items = (long long)n_tiles * ((long long)B + sumG / seg_small);
    long long grid = (long long)sms * kMatchCtasPerSm;
    grid = grid < max_items ? grid : max_items;
    grid = grid < 1 ? 1 : grid;
    assign_prep_kernel<<<(unsigned)B + n_tiles + 1u, kTile, 0, st>>>(gt, gt_off, reinterpret_cast<const float4 *>(priors), (int)P, B,
                                                                     (int)n_tiles, (int)grid, ws, coarse_tiles > 0 ? tune.seg_a : 0,
                                                                     coarse_tiles < (int)n_tiles ? tune.seg_b : 0);
    JABD_LAUNCH_CHECK("assign_prep_kernel");
    if (flags & JABD_ASSIGN_PREP_ONLY) return JABD_OK;
    assign_match_kernel<<<(unsigned)grid, kMatchThreads, 0, st>>>(reinterpret_cast<const float4 *>(priors), (int)P, ws,
                                                          (flags & JABD_ASSIGN_DENSE) ? 1 : 0, (int)n_tiles, coarse_tiles);
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
    const dim3 grid((unsigned)((P + kEncThreads - 1) / kEncThreads), (unsigned)((B + kEncImages - 1) / kEncImages));
    const bool extra = label_mode != 0 || best_truth_idx || best_truth_overlap || best_prior_idx || best_prior_overlap;
    const int variant = (landm_t ? 4 : 0) | (encode_mode ? 2 : 0) | (extra ? 1 : 0);
    // shared-memory staging of the GT records: as many as the call has, at most kEncRecCap (opt-in above 48 KB, once per device)
    const int rec_cap = (int)(sumG < kEncRecCap ? sumG : kEncRecCap);
    const size_t dyn = (size_t)rec_cap * sizeof(EncRec);
#define JABD_ENC_LAUNCH(L_, E_, X_)                                                                                   \
    do {                                                                                                              \
        rc = enc_smem_optin(match_encode_kernel<L_, E_, X_>, variant);                                                \
        if (rc == JABD_OK) match_encode_kernel<L_, E_, X_><<<grid, kEncThreads, dyn, st>>>(a, ws, B, rec_cap);         \
    } while (0)
    switch (variant) {
    case 0: JABD_ENC_LAUNCH(false, false, false); break;
    case 1: JABD_ENC_LAUNCH(false, false, true); break;
    case 2: JABD_ENC_LAUNCH(false, true, false); break;
    case 3: JABD_ENC_LAUNCH(false, true, true); break;
    case 4: JABD_ENC_LAUNCH(true, false, false); break;
    case 5: JABD_ENC_LAUNCH(true, false, true); break;
    case 6: JABD_ENC_LAUNCH(true, true, false); break;
    default: JABD_ENC_LAUNCH(true, true, true); break;
    }
#undef JABD_ENC_LAUNCH
    if (rc != JABD_OK) return rc;
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

int jabd_assign_batches(const float *priors, int64_t P, const jabd_assign_batch_t *batches, int n_batches, float threshold,
                        float var0, float var1, int label_mode, int encode_mode, int flags, const jabd_stream_t *lanes,
                        int n_lanes, jabd_stream_t stream)
{
    JABD_REQUIRE(n_batches >= 0 && (n_batches == 0 || batches), JABD_EINVAL, "assign_batches: null batch list or negative count");
    int rc = lanes_check(lanes, n_lanes, stream, "assign_batches");
    if (rc != JABD_OK) return rc;
    const int used = n_lanes < n_batches ? n_lanes : n_batches;
    // everything that can be refused is refused before the first lane is forked
    for (int i = 0; i < n_batches; ++i) {
        const jabd_assign_batch_t &b = batches[i];
        rc = check_assign_common(priors, P, b.gt, b.gt_off, b.B, b.sumG, b.workspace, b.workspace_bytes);
        if (rc != JABD_OK) return rc;
        if (b.B == 0 || P == 0) continue;
        JABD_REQUIRE(b.loc_t && b.conf_t, JABD_EINVAL, "assign_batches: batch %d: loc_t/conf_t must not be null", i);
        JABD_REQUIRE(aligned_to(b.loc_t, 16) && aligned_to(b.conf_t, 8) && aligned_to(b.landm_t, 4), JABD_EALIGN,
                     "assign_batches: batch %d: loc_t needs 16-byte, conf_t 8-byte, landm_t 4-byte alignment", i);
        // the same workspace twice is fine on one lane (stream order), a race on two
        for (int j = 0; j < i; ++j)
            JABD_REQUIRE(batches[j].workspace != b.workspace || batches[j].B == 0 || used == 0 || i % used == j % used, JABD_EINVAL,
                         "assign_batches: batches %d and %d share a workspace on different lanes", j, i);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = lanes_fork(st, lanes, used);
    if (rc != JABD_OK) return rc;
    // batches that overlap on lanes are cut into large items (fewer culling ballots, staging round trips and column-key
    // atomics per pair; the longer tail of each launch is filled by its neighbours) unless the caller chose a shape itself
    if (used > 1 && (flags & JABD_ASSIGN_TUNE(255, 255, 127)) == 0) flags |= JABD_ASSIGN_TUNE(kSegMax, kSegMax, 100);
    for (int i = 0; i < n_batches && rc == JABD_OK; ++i) {
        const jabd_assign_batch_t &b = batches[i];
        rc = jabd_assign(priors, P, b.gt, b.gt_off, b.B, b.sumG, threshold, var0, var1, label_mode, encode_mode, flags, b.loc_t,
                         b.conf_t, b.landm_t, nullptr, nullptr, nullptr, nullptr, b.workspace, b.workspace_bytes,
                         used > 0 ? lanes[i % used] : stream);
    }
    return lanes_join(st, lanes, used, rc);
}

int64_t jabd_pack_gt_rows(const float *const *rows, const int *counts, int B, float *gt_packed, int64_t capacity_rows, int *gt_off)
{
    JABD_REQUIRE(B >= 0 && gt_off && (B == 0 || (rows && counts)), JABD_EINVAL, "pack_gt_rows: null pointer or negative B");
    int64_t total = 0;
    gt_off[0] = 0;
    for (int b = 0; b < B; ++b) {
        const int g = counts[b];
        JABD_REQUIRE(g >= 0 && (g == 0 || rows[b]), JABD_EINVAL, "pack_gt_rows: image %d has a negative count or a null array", b);
        JABD_REQUIRE(total + g <= capacity_rows && total + g < (1ll << 31), JABD_EWORKSPACE,
                     "pack_gt_rows: %lld rows exceed the packed buffer's capacity %lld", (long long)(total + g), (long long)capacity_rows);
        if (g) {
            JABD_REQUIRE(gt_packed != nullptr, JABD_EINVAL, "pack_gt_rows: gt_packed is null");
            memcpy(gt_packed + total * JABD_GT_ROW, rows[b], sizeof(float) * JABD_GT_ROW * (size_t)g);
        }
        total += g;
        gt_off[b + 1] = (int)total;
    }
    return total;
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

// Layout of the three target tensors inside the device staging area of jabd_assign_host (and of a host block that wants them
// back in one copy): loc_t | conf_t | landm_t, each rounded up to 256 bytes.
static void host_out_offsets(int B, int64_t P, int with_landm, size_t out[4])
{
    const size_t bp = (size_t)(B > 0 ? B : 0) * (size_t)(P > 0 ? P : 0);
    out[0] = 0;
    out[1] = round_up(sizeof(float) * 4 * bp, 256);
    out[2] = out[1] + round_up(sizeof(int64_t) * bp, 256);
    out[3] = out[2] + (with_landm ? round_up(sizeof(float) * 10 * bp, 256) : 0);
}

int jabd_assign_host_out_offsets(int B, int64_t P, int with_landm, size_t *offsets4)
{
    JABD_REQUIRE(B >= 0 && P >= 0 && offsets4, JABD_EINVAL, "assign_host_out_offsets: negative size or null pointer");
    host_out_offsets(B, P, with_landm, offsets4);
    return JABD_OK;
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
    const bool dev_out = (flags & JABD_ASSIGN_DEVICE_OUT) != 0; // outputs are device buffers: written in place, not staged
    float *d_loc = reinterpret_cast<float *>(take(sizeof(float) * 4 * (size_t)B * P));
    int64_t *d_conf = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * (size_t)B * P));
    float *d_landm = with_landm ? reinterpret_cast<float *>(take(sizeof(float) * 10 * (size_t)B * P)) : nullptr;
    if (dev_out) {
        d_loc = loc_t_host;
        d_conf = conf_t_host;
        d_landm = landm_t_host;
    }
    void *d_ws = base + off;
    const size_t ws_bytes = dev_scratch_bytes - off;
    if (sumG > 0) JABD_CUDA(cudaMemcpyAsync(d_gt, gt_host, sizeof(float) * JABD_GT_ROW * (size_t)sumG, cudaMemcpyHostToDevice, st));
    JABD_CUDA(cudaMemcpyAsync(d_off, gt_off_host, sizeof(int) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
    int rc = jabd_assign(priors_dev, P, d_gt, d_off, B, sumG, threshold, var0, var1, label_mode, encode_mode,
                         flags & JABD_ASSIGN_DENSE, d_loc,
                         d_conf, d_landm, nullptr, nullptr, nullptr, nullptr, d_ws, ws_bytes, stream);
    if (rc != JABD_OK) return rc;
    if (!dev_out) {
        // Host outputs laid out like the staging area (jabd_assign_host_out_offsets: one block, three views) come back
        // in ONE copy; three unrelated buffers take three.
        size_t o[4];
        host_out_offsets(B, P, with_landm, o);
        const char *h0 = reinterpret_cast<const char *>(loc_t_host);
        const bool one_block = reinterpret_cast<const char *>(conf_t_host) == h0 + o[1] &&
                               (!with_landm || reinterpret_cast<const char *>(landm_t_host) == h0 + o[2]);
        const size_t bp = (size_t)B * (size_t)P;
        if (one_block) {
            const size_t bytes = with_landm ? o[2] + sizeof(float) * 10 * bp : o[1] + sizeof(int64_t) * bp;
            JABD_CUDA(cudaMemcpyAsync(loc_t_host, d_loc, bytes, cudaMemcpyDeviceToHost, st));
        } else {
            JABD_CUDA(cudaMemcpyAsync(loc_t_host, d_loc, sizeof(float) * 4 * bp, cudaMemcpyDeviceToHost, st));
            JABD_CUDA(cudaMemcpyAsync(conf_t_host, d_conf, sizeof(int64_t) * bp, cudaMemcpyDeviceToHost, st));
            if (with_landm) JABD_CUDA(cudaMemcpyAsync(landm_t_host, d_landm, sizeof(float) * 10 * bp, cudaMemcpyDeviceToHost, st));
        }
    }
    if (!(flags & JABD_ASSIGN_ASYNC)) JABD_CUDA(cudaStreamSynchronize(st));
    return JABD_OK;
}

} // extern "C"
