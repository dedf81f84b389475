// loss.cu -- MultiBox loss on top of the assigned targets: hard-negative mining and the three loss reductions.
//
// Replaces the second half of MultiBoxLoss.forward (R/nets/retinaface_training.py:229-303) -- SURVEY 8(f) rank 1, the
// immediate consumer of loc_t / conf_t / landm_t -- and its autograd backward:
//   pos  = conf_t != 0, pos1 = conf_t > 0                                                     (:236, :243)
//   loss_landm = sum smooth_l1(landm_data - landm_t) over pos1;  loss_l = sum smooth_l1(loc_data - loc_t) over pos
//   rank value  loss_c = log_sum_exp(conf) - conf[:, conf_t], 0 for pos, log_sum_exp with the GLOBAL maximum (:86-88)
//   neg  = rank of loss_c within the image (descending, stable) < min(negpos_ratio * num_pos, P - 1)    (:270-281)
//   loss_c = sum cross_entropy(conf, conf_t in {0,1}) over pos | neg;  all divided by N = max(sum num_pos, 1) (N1 for
//   the landmarks).
// The reference ranks with two full [B,P] sorts (:270-271); here one 1024-thread CTA per image radix-selects the
// num_neg-th largest rank value in three histogram passes over the order-preserving bits (11+11+10), which gives the
// same set: everything above the cut, plus the lowest-index elements at the cut value (what a stable descending sort
// ranks first).  Sums are accumulated per thread in fp32, reduced per image and over the batch in fp64 in a fixed
// order (deterministic; no float atomics).
//
//   mbl_max_kernel       global maximum of conf_data (ordered-uint atomicMax)
//   mbl_forward_kernel   per image: counts, radix select, selection mask [B,P] (bit0 pos, bit1 pos1, bit2 neg), partial sums
//   mbl_finalize_kernel  batch sums -> losses[3], norms[2]
//   mbl_backward_kernel  per prior: gradients of the three losses w.r.t. loc_data / conf_data / landm_data
#include <cooperative_groups.h>

#include "iou_family.cuh"

namespace jabd {

namespace cg = cooperative_groups;

constexpr int kLossThreads = 1024;
constexpr int kLossBins = 2048;
constexpr int kMaxParts = 148 * 8; // CTAs of mbl_max_kernel
#ifndef JABD_LOSS_CLUSTER
#define JABD_LOSS_CLUSTER 4
#endif
constexpr int kLossCluster = JABD_LOSS_CLUSTER; // CTAs (SMs) per image: a thread-block cluster whose histograms meet in distributed shared memory

struct LossWs {
    unsigned *xmax;   // [kMaxParts] ordered bits of max(conf_data) per CTA of mbl_max_kernel (no memset, no atomics)
    double *partial;  // [B*kLossCluster,4] sum_l, sum_c, sum_landm, (unused): one row per CTA
    int *counts;      // [B,2] num_pos, num_pos1
};

static size_t loss_ws_layout(int B, LossWs *w, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += round_up(bytes, 256);
        return o;
    };
    const size_t nb = (size_t)(B > 0 ? B : 1);
    size_t o_x = take(sizeof(unsigned) * kMaxParts);
    size_t o_p = take(sizeof(double) * 4 * nb * kLossCluster);
    size_t o_c = take(sizeof(int) * 2 * nb);
    if (w) {
        w->xmax = reinterpret_cast<unsigned *>(base + o_x);
        w->partial = reinterpret_cast<double *>(base + o_p);
        w->counts = reinterpret_cast<int *>(base + o_c);
    }
    return off;
}

__device__ __forceinline__ float smooth_l1(float d)
{
    const float a = fabsf(d);
    return a < 1.0f ? fmul(fmul(0.5f, d), d) : fsub(a, 0.5f); // F.smooth_l1_loss, beta = 1
}
__device__ __forceinline__ float smooth_l1_grad(float d) { return fabsf(d) < 1.0f ? d : (d > 0.0f ? 1.0f : -1.0f); }

// log_sum_exp(x)[row] - x[row, 0] with the global maximum M (R/nets/retinaface_training.py:86-88, :265)
__device__ __forceinline__ float rank_value(float2 c, float M)
{
    const float s = fadd(exp_f32(fsub(c.x, M)), exp_f32(fsub(c.y, M)));
    return fsub(fadd(log_f32(s), M), c.x);
}

struct LossFwdArgs;
__device__ __forceinline__ uint32_t rank_bits(const LossFwdArgs &a, const uint32_t *s_rank, int cache_cap, long long row0, long long p,
                                              long long p_lo, float M);

__global__ void __launch_bounds__(256) mbl_max_kernel(const float *__restrict__ x, long long n, unsigned *out)
{
    __shared__ unsigned s_m[8];
    unsigned m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned u = ord_of(__ldg(x + i));
        m = u > m ? u : m;
    }
    m = __reduce_max_sync(kFull, m);
    if (lane_id() == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned r = 0;
        for (int w = 0; w < 8; ++w) r = s_m[w] > r ? s_m[w] : r;
        out[blockIdx.x] = r; // one partial maximum per CTA; mbl_forward_kernel reduces them
    }
}

struct LossSmem {
    unsigned local[kLossBins];   // this CTA's histogram (read by the other CTAs of the cluster)
    unsigned hist[kLossBins];    // the image's histogram: sum over the cluster
    unsigned before[kLossBins];  // per bin: count in the CTAs of lower rank (index order of tied values)
    int cnt_local[2];            // this CTA's num_pos, num_pos1
    unsigned wsum[32];
    unsigned wsum2[32];
    double dsum[3][32];
    unsigned found_bin, found_above, found_cnt;
    int total[2];
};

// bin d (from the top) that holds the `want`-th largest element: above < want <= above + hist[d]; warp 0 only
__device__ __forceinline__ void loss_find_bin(LossSmem &sm, int nbins, unsigned want)
{
    if (threadIdx.x < 32) {
        const unsigned lane = lane_id();
        const int per = nbins / 32;
        const int hi = nbins - 1 - (int)lane * per; // this lane owns bins hi, hi-1, ..., hi-per+1
        unsigned s = 0;
        for (int k = 0; k < per; ++k) s += sm.hist[hi - k];
        unsigned inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(kFull, inc, o);
            if (lane >= (unsigned)o) inc += y;
        }
        unsigned above = inc - s;
        if (above < want && want <= inc) {
            for (int k = 0; k < per; ++k) {
                const unsigned h = sm.hist[hi - k];
                if (want <= above + h) { sm.found_bin = (unsigned)(hi - k); sm.found_above = above; sm.found_cnt = h; break; }
                above += h;
            }
        }
    }
    __syncthreads();
}

struct LossFwdArgs {
    const float4 *loc_data;
    const float2 *conf_data;
    const float *landm_data;
    const float4 *loc_t;
    const long long *conf_t;
    const float *landm_t;
    long long P;
    int negpos_ratio;
    unsigned char *mask;
    LossWs ws;
    // box regression term: 0 = smooth-L1 on encoded offsets (R/nets/retinaface_training.py:252), 1..4 = IouLoss of
    // R/nets/retinaface_training_DIOU.py:491-525 ('Iou' / 'Giou' / 'Diou' / 'Ciou') on decode(loc_data, priors) against the
    // raw matched boxes that match_iou leaves in loc_t (:590-591)
    int loc_loss;
    const float4 *priors;
    float var0, var1;
};

__device__ __forceinline__ uint32_t rank_bits(const LossFwdArgs &a, const uint32_t *s_rank, int cache_cap, long long row0, long long p,
                                              long long p_lo, float M)
{
    if (p - p_lo < cache_cap) return s_rank[p - p_lo];
    const bool pos = a.conf_t[row0 + p] != 0;
    return ord_of(pos ? 0.0f : rank_value(__ldg(a.conf_data + row0 + p), M));
}

// One cluster of kLossCluster CTAs per image, CTA r owning the contiguous prior range r: the order-preserving bits of every
// prior's rank value are computed once (pass 1) and kept in dynamic shared memory (`cache_cap` entries per CTA; priors
// beyond that are recomputed in the later passes), each histogram is built locally and summed over the cluster through
// distributed shared memory, so that every CTA finds the same cut without a global round trip.
__device__ __forceinline__ void cluster_merge_hist(LossSmem &sm, int nbins, cg::cluster_group &cluster, bool with_counts)
{
    cluster.sync(); // every CTA's local histogram is complete
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < nbins; i += kLossThreads) {
        unsigned sum = 0, bef = 0;
        for (unsigned r = 0; r < (unsigned)kLossCluster; ++r) {
            const unsigned v = cluster.map_shared_rank(sm.local, r)[i];
            sum += v;
            if (r < rank) bef += v;
        }
        sm.hist[i] = sum;
        sm.before[i] = bef;
    }
    if (with_counts && threadIdx.x == 0) {
        int s0 = 0, s1 = 0;
        for (unsigned r = 0; r < (unsigned)kLossCluster; ++r) {
            const int *c = cluster.map_shared_rank(sm.cnt_local, r);
            s0 += c[0];
            s1 += c[1];
        }
        sm.total[0] = s0;
        sm.total[1] = s1;
    }
    cluster.sync(); // nobody overwrites its local histogram while a neighbour still reads it
}

__global__ void __launch_bounds__(kLossThreads, 1) mbl_forward_kernel(LossFwdArgs a, int cache_cap, int n_parts)
{
    extern __shared__ __align__(16) uint32_t s_rank[];
    __shared__ LossSmem sm;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = tid >> 5;
    const int crank = (int)cluster.block_rank();
    const int b = blockIdx.x / kLossCluster;
    const long long P = a.P;
    const long long row0 = (long long)b * P;
    const long long per = (P + kLossCluster - 1) / kLossCluster;
    const long long p_lo = crank * per < P ? crank * per : P, p_hi = (p_lo + per) < P ? (p_lo + per) : P;
    // global maximum of conf_data: reduce mbl_max_kernel's per-CTA partials (<= kMaxParts of them)
    {
        unsigned m = 0;
        for (int i = tid; i < n_parts; i += kLossThreads) { const unsigned u = a.ws.xmax[i]; m = u > m ? u : m; }
        m = __reduce_max_sync(kFull, m);
        if (lane == 0) sm.wsum[warp] = m;
        __syncthreads();
        if (tid == 0) {
            unsigned r = 0;
            for (int w = 0; w < kLossThreads / 32; ++w) r = sm.wsum[w] > r ? sm.wsum[w] : r;
            sm.found_bin = r;
        }
        __syncthreads();
    }
    const float M = ord_inv(sm.found_bin);
    __syncthreads();

    // ---- pass 1: counts + top 11 bits of the rank value
    for (int i = tid; i < kLossBins; i += kLossThreads) sm.local[i] = 0;
    __syncthreads();
    int npos = 0, npos1 = 0;
    for (long long base = p_lo + tid; base < p_hi; base += 4ll * kLossThreads) { // four loads in flight before the first atomic
        long long ct[4];
        float2 cd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long p = base + (long long)k * kLossThreads;
            ct[k] = p < p_hi ? a.conf_t[row0 + p] : 1;
            cd[k] = p < p_hi ? __ldg(a.conf_data + row0 + p) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long p = base + (long long)k * kLossThreads;
            if (p >= p_hi) continue;
            const bool pos = ct[k] != 0;
            npos += pos ? 1 : 0;
            npos1 += ct[k] > 0 ? 1 : 0;
            const uint32_t u = ord_of(pos ? 0.0f : rank_value(cd[k], M));
            if (p - p_lo < cache_cap) s_rank[p - p_lo] = u;
            atomicAdd(&sm.local[u >> 21], 1u);
        }
    }
    npos = __reduce_add_sync(kFull, npos);
    npos1 = __reduce_add_sync(kFull, npos1);
    if (lane == 0) { sm.wsum[warp] = (unsigned)npos; sm.wsum2[warp] = (unsigned)npos1; }
    __syncthreads();
    if (tid == 0) {
        int s0 = 0, s1 = 0;
        for (int w = 0; w < kLossThreads / 32; ++w) { s0 += (int)sm.wsum[w]; s1 += (int)sm.wsum2[w]; }
        sm.cnt_local[0] = s0;
        sm.cnt_local[1] = s1;
    }
    cluster_merge_hist(sm, kLossBins, cluster, true);
    const int num_pos = sm.total[0], num_pos1 = sm.total[1];
    long long want_ll = (long long)a.negpos_ratio * num_pos; // torch.clamp(negpos_ratio * num_pos, max = P - 1), :280
    if (want_ll > P - 1) want_ll = P - 1;
    const unsigned want = want_ll > 0 ? (unsigned)want_ll : 0u;

    // ---- passes 2, 3: exact cut value T, number of elements above it, quota of ties (same decisions in every CTA)
    uint32_t T = 0xffffffffu;
    unsigned quota = 0, eq_total = 0, eq_before = 0;
    if (want > 0) {
        loss_find_bin(sm, kLossBins, want);
        const uint32_t b1 = sm.found_bin;
        const unsigned above1 = sm.found_above;
        __syncthreads();
        for (int i = tid; i < kLossBins; i += kLossThreads) sm.local[i] = 0;
        __syncthreads();
        for (long long p = p_lo + tid; p < p_hi; p += kLossThreads) {
            const uint32_t u = rank_bits(a, s_rank, cache_cap, row0, p, p_lo, M);
            if ((u >> 21) == b1) atomicAdd(&sm.local[(u >> 10) & 0x7ffu], 1u);
        }
        cluster_merge_hist(sm, kLossBins, cluster, false);
        loss_find_bin(sm, kLossBins, want - above1);
        const uint32_t b2 = sm.found_bin;
        const unsigned above2 = sm.found_above;
        __syncthreads();
        const uint32_t pre = (b1 << 11) | b2;
        for (int i = tid; i < 1024; i += kLossThreads) sm.local[i] = 0;
        __syncthreads();
        for (long long p = p_lo + tid; p < p_hi; p += kLossThreads) {
            const uint32_t u = rank_bits(a, s_rank, cache_cap, row0, p, p_lo, M);
            if ((u >> 10) == pre) atomicAdd(&sm.local[u & 0x3ffu], 1u);
        }
        cluster_merge_hist(sm, 1024, cluster, false);
        loss_find_bin(sm, 1024, want - above1 - above2);
        T = (pre << 10) | sm.found_bin;
        eq_total = sm.found_cnt;
        eq_before = sm.before[sm.found_bin];
        quota = want - (above1 + above2 + sm.found_above);
        __syncthreads();
    }
    const bool ordered_ties = want > 0 && quota < eq_total; // only some of the elements at the cut value are taken

    // ---- final pass: selection mask and the three sums
    float sl = 0.0f, sc = 0.0f, sn = 0.0f;
    unsigned eq_seen = eq_before; // ties at the cut value in the prior ranges of the lower-ranked CTAs come first
    for (long long base = p_lo; base < p_hi; base += kLossThreads) {
        const long long p = base + tid;
        bool pos = false, pos1 = false, neg = false, tie = false;
        float2 c = make_float2(0.f, 0.f);
        if (p < p_hi) {
            const long long ct = a.conf_t[row0 + p];
            pos = ct != 0;
            pos1 = ct > 0;
            c = __ldg(a.conf_data + row0 + p);
            if (want > 0) {
                const uint32_t u = (p - p_lo) < cache_cap ? s_rank[p - p_lo] : ord_of(pos ? 0.0f : rank_value(c, M));
                neg = u > T;
                tie = u == T;
            }
        }
        if (ordered_ties) { // block-wide rank of the ties in index order (rare path)
            const unsigned bal = __ballot_sync(kFull, tie);
            if (lane == 0) sm.wsum[warp] = __popc(bal);
            __syncthreads();
            unsigned before = 0, tot = 0;
            for (int w = 0; w < kLossThreads / 32; ++w) {
                const unsigned v = sm.wsum[w];
                if (w < warp) before += v;
                tot += v;
            }
            const unsigned rank = eq_seen + before + __popc(bal & lanemask_lt());
            neg = neg || (tie && rank < quota);
            eq_seen += tot;
            __syncthreads();
        } else {
            neg = neg || tie;
        }
        if (p < p_hi) {
            a.mask[row0 + p] = (unsigned char)((pos ? 1 : 0) | (pos1 ? 2 : 0) | (neg ? 4 : 0));
            if (pos || neg) { // cross entropy with target = pos ? 1 : 0 (conf_t[pos] = 1, :259), stable log-softmax
                const float mx = fmaxf(c.x, c.y);
                const float lse = fadd(log_f32(fadd(exp_f32(fsub(c.x, mx)), exp_f32(fsub(c.y, mx)))), mx);
                sc = fadd(sc, fsub(lse, pos ? c.y : c.x));
            }
            if (pos) {
                const float4 lp = __ldg(a.loc_data + row0 + p), lt = __ldg(a.loc_t + row0 + p);
                if (a.loc_loss == 0) {
                    sl = fadd(sl, fadd(fadd(smooth_l1(fsub(lp.x, lt.x)), smooth_l1(fsub(lp.y, lt.y))),
                                       fadd(smooth_l1(fsub(lp.z, lt.z)), smooth_l1(fsub(lp.w, lt.w)))));
                } else {
                    const float4 box = decode_box(lp, __ldg(a.priors + p), a.var0, a.var1);
                    sl = fadd(sl, fsub(1.0f, iou_family<float>(a.loc_loss, box_of(box), box_of(lt))));
                }
            }
            if (pos1) {
                const float *mp = a.landm_data + (row0 + p) * 10, *mt = a.landm_t + (row0 + p) * 10;
#pragma unroll
                for (int k = 0; k < 10; ++k) sn = fadd(sn, smooth_l1(fsub(__ldg(mp + k), __ldg(mt + k))));
            }
        }
    }
    // fixed-order reduction in fp64: lanes (butterfly), then warps in index order
    double dl = sl, dc = sc, dn = sn;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dl += __shfl_xor_sync(kFull, dl, o);
        dc += __shfl_xor_sync(kFull, dc, o);
        dn += __shfl_xor_sync(kFull, dn, o);
    }
    if (lane == 0) { sm.dsum[0][warp] = dl; sm.dsum[1][warp] = dc; sm.dsum[2][warp] = dn; }
    __syncthreads();
    if (tid == 0) {
        double tl = 0.0, tc = 0.0, tn = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) { tl += sm.dsum[0][w]; tc += sm.dsum[1][w]; tn += sm.dsum[2][w]; }
        double *out = a.ws.partial + 4 * ((size_t)b * kLossCluster + crank);
        out[0] = tl;
        out[1] = tc;
        out[2] = tn;
        if (crank == 0) {
            a.ws.counts[2 * b + 0] = num_pos;
            a.ws.counts[2 * b + 1] = num_pos1;
        }
    }
}

__global__ void __launch_bounds__(32) mbl_finalize_kernel(LossWs ws, int B, float *losses, float *norms)
{
    if (threadIdx.x != 0) return;
    double tl = 0.0, tc = 0.0, tn = 0.0;
    long long n = 0, n1 = 0;
    for (int b = 0; b < B; ++b) {
        for (int r = 0; r < kLossCluster; ++r) { // fixed order: images, then the cluster's CTAs
            const double *q = ws.partial + 4 * ((size_t)b * kLossCluster + r);
            tl += q[0];
            tc += q[1];
            tn += q[2];
        }
        n += ws.counts[2 * b + 0];
        n1 += ws.counts[2 * b + 1];
    }
    const float N = n > 1 ? (float)n : 1.0f, N1 = n1 > 1 ? (float)n1 : 1.0f; // max(num_pos.sum().float(), 1), :293, :299
    losses[0] = fdiv((float)tl, N);
    losses[1] = fdiv((float)tc, N);
    losses[2] = fdiv((float)tn, N1);
    norms[0] = N;
    norms[1] = N1;
}

struct LossBwdArgs {
    const float4 *loc_data;
    const float2 *conf_data;
    const float *landm_data;
    const float4 *loc_t;
    const float *landm_t;
    const unsigned char *mask;
    const float *norms;
    const float *grad_losses;
    long long n; // B * P
    float4 *g_loc;
    float2 *g_conf;
    float *g_landm;
    int loc_loss;       // see LossFwdArgs
    const float4 *priors;
    long long P;
    float var0, var1;
};

__global__ void __launch_bounds__(256) mbl_backward_kernel(LossBwdArgs a)
{
    __shared__ __align__(16) float s_lm[256 * 10];
    const int tid = threadIdx.x;
    const long long i0 = (long long)blockIdx.x * 256;
    const long long i = i0 + tid;
    const float gl = fdiv(__ldg(a.grad_losses + 0), __ldg(a.norms + 0));
    const float gc = fdiv(__ldg(a.grad_losses + 1), __ldg(a.norms + 0));
    const float gn = fdiv(__ldg(a.grad_losses + 2), __ldg(a.norms + 1));
    float lm[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) lm[k] = 0.0f;
    if (i < a.n) {
        const unsigned m = a.mask[i];
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 g2 = make_float2(0.f, 0.f);
        if (m & 1u) {
            const float4 lp = __ldg(a.loc_data + i), lt = __ldg(a.loc_t + i);
            if (a.loc_loss == 0) {
                g4.x = fmul(gl, smooth_l1_grad(fsub(lp.x, lt.x)));
                g4.y = fmul(gl, smooth_l1_grad(fsub(lp.y, lt.y)));
                g4.z = fmul(gl, smooth_l1_grad(fsub(lp.z, lt.z)));
                g4.w = fmul(gl, smooth_l1_grad(fsub(lp.w, lt.w)));
            } else {
                const float4 r = iou_loss_grad(a.loc_loss, lp, __ldg(a.priors + i % a.P), lt, a.var0, a.var1);
                g4 = make_float4(gl * r.x, gl * r.y, gl * r.z, gl * r.w);
            }
        }
        if (m & 5u) { // softmax - onehot(target), target = pos ? 1 : 0
            const float2 c = __ldg(a.conf_data + i);
            const float mx = fmaxf(c.x, c.y);
            const float e0 = exp_f32(fsub(c.x, mx)), e1 = exp_f32(fsub(c.y, mx));
            const float s = fadd(e0, e1);
            const float p0 = fdiv(e0, s), p1 = fdiv(e1, s);
            g2.x = fmul(gc, (m & 1u) ? p0 : fsub(p0, 1.0f));
            g2.y = fmul(gc, (m & 1u) ? fsub(p1, 1.0f) : p1);
        }
        if (m & 2u) {
            const float *mp = a.landm_data + i * 10, *mt = a.landm_t + i * 10;
#pragma unroll
            for (int k = 0; k < 10; ++k) lm[k] = fmul(gn, smooth_l1_grad(fsub(__ldg(mp + k), __ldg(mt + k))));
        }
        a.g_loc[i] = g4;
        a.g_conf[i] = g2;
    }
    // landmark gradients through shared memory: contiguous 16-byte stores
#pragma unroll
    for (int k = 0; k < 10; ++k) s_lm[tid * 10 + k] = lm[k];
    __syncthreads();
    const long long left = a.n - i0;
    const int nv = left < 256 ? (int)left : 256;
    float *dst = a.g_landm + i0 * 10;
    const int nf = nv * 10;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int n4 = nf >> 2;
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        const float4 *s4 = reinterpret_cast<const float4 *>(s_lm);
        for (int k = tid; k < n4; k += 256) d4[k] = s4[k];
        for (int k = (n4 << 2) + tid; k < nf; k += 256) dst[k] = s_lm[k];
    } else {
        for (int k = tid; k < nf; k += 256) dst[k] = s_lm[k];
    }
}

} // namespace jabd

using namespace jabd;

extern "C" {

size_t jabd_multibox_loss_workspace_bytes(int B)
{
    if (B < 0) return 0;
    return loss_ws_layout(B, nullptr, nullptr);
}

int jabd_multibox_loss_forward_ex(const float *loc_data, const float *conf_data, const float *landm_data, const float *loc_t,
                                  const int64_t *conf_t, const float *landm_t, int B, int64_t P, int negpos_ratio, int loc_loss,
                                  const float *priors, float var0, float var1, float *losses, float *norms,
                                  unsigned char *sel_mask, void *workspace, size_t workspace_bytes, jabd_stream_t stream)
{
    JABD_REQUIRE(loc_loss >= 0 && loc_loss <= kCiou, JABD_EINVAL, "multibox_loss: loc_loss must be 0 (smooth-L1) or 1..4 (IoU family)");
    JABD_REQUIRE(loc_loss == 0 || (priors && aligned_to(priors, 16)), JABD_EINVAL,
                 "multibox_loss: the IoU-family box loss needs 16-byte aligned priors");
    JABD_REQUIRE(B >= 0 && P >= 0 && negpos_ratio >= 0, JABD_EINVAL, "multibox_loss: negative size");
    JABD_REQUIRE(B <= (1 << 20) && (int64_t)B * P < (1ll << 40), JABD_EINVAL, "multibox_loss: batch too large");
    JABD_REQUIRE(losses && norms, JABD_EINVAL, "multibox_loss: null output pointer");
    JABD_REQUIRE(workspace && aligned_to(workspace, 256), JABD_EWORKSPACE, "multibox_loss: workspace null or not 256-byte aligned");
    JABD_REQUIRE(workspace_bytes >= loss_ws_layout(B, nullptr, nullptr), JABD_EWORKSPACE, "multibox_loss: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LossWs ws;
    loss_ws_layout(B, &ws, static_cast<char *>(workspace));
    if (B > 0 && P > 0) {
        JABD_REQUIRE(loc_data && conf_data && landm_data && loc_t && conf_t && landm_t && sel_mask, JABD_EINVAL,
                     "multibox_loss: null input pointer");
        JABD_REQUIRE(aligned_to(loc_data, 16) && aligned_to(loc_t, 16) && aligned_to(conf_data, 8) && aligned_to(conf_t, 8),
                     JABD_EALIGN, "multibox_loss: loc needs 16-byte, conf 8-byte alignment");
        const long long n = 2ll * B * P;
        long long grid = (n + 256 * 8 - 1) / (256 * 8);
        grid = grid > kMaxParts ? kMaxParts : (grid < 1 ? 1 : grid);
        mbl_max_kernel<<<(unsigned)grid, 256, 0, st>>>(conf_data, n, ws.xmax);
        JABD_LAUNCH_CHECK("mbl_max_kernel");
        LossFwdArgs a;
        a.loc_data = reinterpret_cast<const float4 *>(loc_data);
        a.conf_data = reinterpret_cast<const float2 *>(conf_data);
        a.landm_data = landm_data;
        a.loc_t = reinterpret_cast<const float4 *>(loc_t);
        a.conf_t = reinterpret_cast<const long long *>(conf_t);
        a.landm_t = landm_t;
        a.P = P;
        a.negpos_ratio = negpos_ratio;
        a.mask = sel_mask;
        a.ws = ws;
        a.loc_loss = loc_loss;
        a.priors = reinterpret_cast<const float4 *>(priors);
        a.var0 = var0;
        a.var1 = var1;
        // rank-value cache: as many priors of a CTA's range as fit next to the kernel's static shared memory (opt-in > 48 KB)
        constexpr long long kCacheMax = 48 * 1024;   // entries (192 KB)
        const long long per = (P + kLossCluster - 1) / kLossCluster;
        const int cache_cap = (int)(per < kCacheMax ? per : kCacheMax);
        const size_t dyn = sizeof(uint32_t) * (size_t)cache_cap;
        static bool attr_done[64] = {};
        int devi = 0;
        JABD_CUDA(cudaGetDevice(&devi));
        if (!(devi >= 0 && devi < 64 && attr_done[devi])) {
            JABD_CUDA(cudaFuncSetAttribute(mbl_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) * kCacheMax)));
            if (devi >= 0 && devi < 64) attr_done[devi] = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)B * kLossCluster);
        cfg.blockDim = dim3(kLossThreads);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kLossCluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        JABD_CUDA(cudaLaunchKernelEx(&cfg, mbl_forward_kernel, a, cache_cap, (int)grid));
        JABD_LAUNCH_CHECK("mbl_forward_kernel");
    }
    mbl_finalize_kernel<<<1, 32, 0, st>>>(ws, (B > 0 && P > 0) ? B : 0, losses, norms);
    JABD_LAUNCH_CHECK("mbl_finalize_kernel");
    return JABD_OK;
}

int jabd_multibox_loss_forward(const float *loc_data, const float *conf_data, const float *landm_data, const float *loc_t,
                               const int64_t *conf_t, const float *landm_t, int B, int64_t P, int negpos_ratio, float *losses,
                               float *norms, unsigned char *sel_mask, void *workspace, size_t workspace_bytes,
                               jabd_stream_t stream)
{
    return jabd_multibox_loss_forward_ex(loc_data, conf_data, landm_data, loc_t, conf_t, landm_t, B, P, negpos_ratio, 0, nullptr,
                                         0.0f, 0.0f, losses, norms, sel_mask, workspace, workspace_bytes, stream);
}

int jabd_multibox_loss_backward(const float *loc_data, const float *conf_data, const float *landm_data, const float *loc_t,
                                const float *landm_t, const unsigned char *sel_mask, const float *norms,
                                const float *grad_losses, int B, int64_t P, float *g_loc, float *g_conf, float *g_landm,
                                jabd_stream_t stream)
{
    return jabd_multibox_loss_backward_ex(loc_data, conf_data, landm_data, loc_t, landm_t, sel_mask, norms, grad_losses, B, P, 0,
                                          nullptr, 0.0f, 0.0f, g_loc, g_conf, g_landm, stream);
}

int jabd_multibox_loss_backward_ex(const float *loc_data, const float *conf_data, const float *landm_data, const float *loc_t,
                                   const float *landm_t, const unsigned char *sel_mask, const float *norms,
                                   const float *grad_losses, int B, int64_t P, int loc_loss, const float *priors, float var0,
                                   float var1, float *g_loc, float *g_conf, float *g_landm, jabd_stream_t stream)
{
    JABD_REQUIRE(loc_loss >= 0 && loc_loss <= kCiou, JABD_EINVAL, "multibox_loss_backward: loc_loss must be 0..4");
    JABD_REQUIRE(loc_loss == 0 || (priors && aligned_to(priors, 16)), JABD_EINVAL,
                 "multibox_loss_backward: the IoU-family box loss needs 16-byte aligned priors");
    JABD_REQUIRE(B >= 0 && P >= 0, JABD_EINVAL, "multibox_loss_backward: negative size");
    if (B == 0 || P == 0) return JABD_OK;
    JABD_REQUIRE(loc_data && conf_data && landm_data && loc_t && landm_t && sel_mask && norms && grad_losses && g_loc && g_conf &&
                     g_landm, JABD_EINVAL, "multibox_loss_backward: null pointer");
    JABD_REQUIRE(aligned_to(loc_data, 16) && aligned_to(loc_t, 16) && aligned_to(g_loc, 16) && aligned_to(conf_data, 8) &&
                     aligned_to(g_conf, 8), JABD_EALIGN, "multibox_loss_backward: loc needs 16-byte, conf 8-byte alignment");
    LossBwdArgs a;
    a.loc_data = reinterpret_cast<const float4 *>(loc_data);
    a.conf_data = reinterpret_cast<const float2 *>(conf_data);
    a.landm_data = landm_data;
    a.loc_t = reinterpret_cast<const float4 *>(loc_t);
    a.landm_t = landm_t;
    a.mask = sel_mask;
    a.norms = norms;
    a.grad_losses = grad_losses;
    a.n = (long long)B * P;
    a.g_loc = reinterpret_cast<float4 *>(g_loc);
    a.g_conf = reinterpret_cast<float2 *>(g_conf);
    a.g_landm = g_landm;
    a.loc_loss = loc_loss;
    a.priors = reinterpret_cast<const float4 *>(priors);
    a.P = P;
    a.var0 = var0;
    a.var1 = var1;
    mbl_backward_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    JABD_LAUNCH_CHECK("mbl_backward_kernel");
    return JABD_OK;
}

} // extern "C"
