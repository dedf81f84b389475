// boxcodec.cu -- prior generation and the stand-alone box operators of the reference's public surface:
// Anchors.get_anchors (R/utils/anchors.py:9-42), point_form / jaccard / encode / encode_landm
// (R/nets/retinaface_training.py:8-84), decode / decode_landm (R/utils/utils_bbox.py:29-46).
// All are streaming, HBM-bound kernels: 16-byte vector loads/stores, one pass, no shared memory.
#include "common.cuh"

namespace jabd {

constexpr int kMaxLevels = 16;
constexpr int kMaxSizes = 64;

struct PriorCfg {
    int n_levels;
    int H, W, clip;
    int step[kMaxLevels];
    int fw[kMaxLevels];
    int ns[kMaxLevels];        // sizes per level
    int soff[kMaxLevels];      // first size of the level in min_size[]
    long long poff[kMaxLevels + 1]; // first prior of the level
    double min_size[kMaxSizes];
};

__global__ void __launch_bounds__(256) priors_kernel(PriorCfg c, float4 *__restrict__ out, long long P)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P) return;
    int k = 0;
#pragma unroll 1
    while (k + 1 < c.n_levels && n >= c.poff[k + 1]) ++k;
    const long long local = n - c.poff[k];
    const int ns = c.ns[k];
    const long long cell = local / ns;
    const int si = (int)(local - cell * ns);
    const int i = (int)(cell / c.fw[k]);
    const int j = (int)(cell - (long long)i * c.fw[k]);
    const double ms = c.min_size[c.soff[k] + si];
    const double step = (double)c.step[k];
    // Python float64 arithmetic, one rounding to fp32 (torch.Tensor(list)); R/utils/anchors.py:31-34
    float4 v;
    v.x = (float)(((double)j + 0.5) * step / (double)c.W);
    v.y = (float)(((double)i + 0.5) * step / (double)c.H);
    v.z = (float)(ms / (double)c.W);
    v.w = (float)(ms / (double)c.H);
    if (c.clip) { // output.clamp_(max=1, min=0), :39-40
        v.x = fminf(fmaxf(v.x, 0.f), 1.f);
        v.y = fminf(fmaxf(v.y, 0.f), 1.f);
        v.z = fminf(fmaxf(v.z, 0.f), 1.f);
        v.w = fminf(fmaxf(v.w, 0.f), 1.f);
    }
    out[n] = v;
}

static int fill_prior_cfg(const int *steps, const double *min_sizes, const int *sizes_off, int n_levels, int H, int W,
                          int clip, PriorCfg *c)
{
    JABD_REQUIRE(steps && sizes_off, JABD_EINVAL, "priors: null config pointer");
    JABD_REQUIRE(n_levels > 0 && n_levels <= kMaxLevels, JABD_EINVAL, "priors: n_levels=%d outside 1..%d", n_levels, kMaxLevels);
    JABD_REQUIRE(H > 0 && W > 0, JABD_EINVAL, "priors: image size must be positive (H=%d W=%d)", H, W);
    JABD_REQUIRE(sizes_off[0] == 0 && sizes_off[n_levels] <= kMaxSizes, JABD_EINVAL, "priors: at most %d min_sizes", kMaxSizes);
    c->n_levels = n_levels;
    c->H = H;
    c->W = W;
    c->clip = clip ? 1 : 0;
    long long off = 0;
    for (int k = 0; k < n_levels; ++k) {
        JABD_REQUIRE(steps[k] > 0, JABD_EINVAL, "priors: step %d must be positive", k);
        const int ns = sizes_off[k + 1] - sizes_off[k];
        JABD_REQUIRE(ns >= 0, JABD_EINVAL, "priors: sizes_off must be non-decreasing");
        const int fh = (H + steps[k] - 1) / steps[k]; // ceil(H/step), :21
        const int fw = (W + steps[k] - 1) / steps[k];
        c->step[k] = steps[k];
        c->fw[k] = fw;
        c->ns[k] = ns > 0 ? ns : 1;
        c->soff[k] = sizes_off[k];
        c->poff[k] = off;
        off += (long long)fh * fw * ns;
    }
    c->poff[n_levels] = off;
    for (int k = n_levels + 1; k <= kMaxLevels; ++k) c->poff[k] = off;
    if (min_sizes)
        for (int s = 0; s < sizes_off[n_levels]; ++s) c->min_size[s] = min_sizes[s];
    return JABD_OK;
}

// ---- elementwise -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) point_form_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = to_point_form(__ldg(in + i));
}

// thread per column b, 16 rows of a per thread: the b box stays in registers, stores are coalesced along b
template <bool INTER_ONLY>
__global__ void __launch_bounds__(256) jaccard_kernel(const float4 *__restrict__ box_a, long long A, const float4 *__restrict__ box_b,
                                                      long long B, float *__restrict__ out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    const float4 b = __ldg(box_b + j);
    const float area_b = box_area(b);
    const long long a0 = (long long)blockIdx.y * 16;
    const long long a1 = (a0 + 16 < A) ? a0 + 16 : A;
    for (long long i = a0; i < a1; ++i) {
        const float4 a = __ldg(box_a + i);
        if (INTER_ONLY) { // intersect(), R/nets/retinaface_training.py:22-39
            const float w = clamp0(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)));
            const float h = clamp0(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)));
            out[i * B + j] = fmul(w, h);
        } else {
            out[i * B + j] = iou_ref(a, box_area(a), b, area_b);
        }
    }
}

__global__ void __launch_bounds__(256) encode_kernel(const float4 *__restrict__ m, const float4 *__restrict__ p, long long n,
                                                     float var0, float var1, float4 *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = encode_box(__ldg(m + i), __ldg(p + i), var0, var1);
}

// landmarks as [n*5] (x,y) pairs: 8-byte loads/stores, prior shared by 5 consecutive threads
template <bool ENCODE>
__global__ void __launch_bounds__(256) landm_kernel(const float2 *__restrict__ in, const float4 *__restrict__ priors, long long P,
                                                    float var0, float2 *__restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; // pair index inside the image
    if (e >= P * 5) return;
    const long long img = (long long)blockIdx.y * P * 5;
    const float4 pr = __ldg(priors + e / 5);
    const float2 v = __ldg(in + img + e);
    float2 o;
    if (ENCODE) {
        o.x = encode_pt(v.x, pr.x, pr.z, var0);
        o.y = encode_pt(v.y, pr.y, pr.w, var0);
    } else {
        o.x = decode_pt(v.x, pr.x, pr.z, var0);
        o.y = decode_pt(v.y, pr.y, pr.w, var0);
    }
    out[img + e] = o;
}

// D1 for a batch: thread = one prior for kDecImgs consecutive images -- the prior is read once, the kDecImgs loc
// vectors are independent 16-byte loads in flight together, stores are 16-byte and coalesced along the prior axis.
constexpr int kDecImgs = 4;
__global__ void __launch_bounds__(256) decode_kernel(const float4 *__restrict__ loc, const float4 *__restrict__ priors, long long P,
                                                     int batch, float var0, float var1, float4 *__restrict__ out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int b0 = blockIdx.y * kDecImgs;
    const float4 pr = __ldg(priors + p);
    float4 l[kDecImgs];
#pragma unroll
    for (int k = 0; k < kDecImgs; ++k)
        if (b0 + k < batch) l[k] = __ldcs(loc + (long long)(b0 + k) * P + p); // streamed once: evict-first
#pragma unroll
    for (int k = 0; k < kDecImgs; ++k)
        if (b0 + k < batch) __stcs(out + (long long)(b0 + k) * P + p, decode_box(l[k], pr, var0, var1));
}

// Box post-processing of Retinaface.detect_image on the kept rows: undo the letterbox padding
// (retinaface_correct_boxes, R/utils/utils_bbox.py:9-24) and scale to pixels (R/predict.py:194-195).  The reference does
// both in numpy: float32 rows op float64 vectors -> float64, stored back to float32 -- so every step here is evaluated in
// fp64 and rounded to fp32 once per step.  post[b] = {offset_x, offset_y, scale_x, scale_y, width, height}.
__global__ void __launch_bounds__(256) correct_boxes_kernel(float *__restrict__ dets, const int *__restrict__ counts,
                                                            const double *__restrict__ post, int B, int K, int letterbox,
                                                            int to_pixels)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K), k = (int)(row % K);
    if (counts && k >= counts[b]) return;
    const double *q = post + 6 * b;
    float *r = dets + row * JABD_DET_ROW;
#pragma unroll
    for (int c = 0; c < JABD_DET_ROW; ++c) {
        if (c == 4) continue; // score
        const int xy = (c < 4 ? c : c - 5) & 1; // 0: x column, 1: y column
        float v = r[c];
        if (letterbox) v = (float)(((double)v - q[xy]) * q[2 + xy]);
        if (to_pixels) v = (float)((double)v * q[4 + xy]);
        r[c] = v;
    }
}

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

} // namespace jabd

using namespace jabd;

extern "C" {

int64_t jabd_priors_count(const int *steps, const int *sizes_off, int n_levels, int H, int W)
{
    PriorCfg c;
    if (fill_prior_cfg(steps, nullptr, sizes_off, n_levels, H, W, 0, &c) != JABD_OK) return -1;
    return (int64_t)c.poff[n_levels];
}

int jabd_priors(const int *steps, const double *min_sizes, const int *sizes_off, int n_levels, int H, int W, int clip,
                float *out, int64_t P, jabd_stream_t stream)
{
    PriorCfg c;
    JABD_REQUIRE(min_sizes != nullptr, JABD_EINVAL, "priors: min_sizes is null");
    int rc = fill_prior_cfg(steps, min_sizes, sizes_off, n_levels, H, W, clip, &c);
    if (rc != JABD_OK) return rc;
    JABD_REQUIRE(P == (int64_t)c.poff[n_levels], JABD_EINVAL, "priors: out holds %lld rows but the config yields %lld",
                 (long long)P, c.poff[n_levels]);
    if (P == 0) return JABD_OK;
    JABD_REQUIRE(out && aligned_to(out, 16), JABD_EALIGN, "priors: out must be non-null and 16-byte aligned");
    priors_kernel<<<blocks_for(P, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(c, reinterpret_cast<float4 *>(out), P);
    JABD_LAUNCH_CHECK("priors_kernel");
    return JABD_OK;
}

int jabd_point_form(const float *boxes, int64_t n, float *out, jabd_stream_t stream)
{
    JABD_REQUIRE(n >= 0, JABD_EINVAL, "point_form: negative n");
    if (n == 0) return JABD_OK;
    JABD_REQUIRE(boxes && out, JABD_EINVAL, "point_form: null pointer");
    JABD_REQUIRE(aligned_to(boxes, 16) && aligned_to(out, 16), JABD_EALIGN, "point_form: 16-byte alignment required");
    point_form_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(boxes), reinterpret_cast<float4 *>(out), n);
    JABD_LAUNCH_CHECK("point_form_kernel");
    return JABD_OK;
}

static int jaccard_common(bool inter_only, const float *box_a, int64_t A, const float *box_b, int64_t B, float *out,
                          jabd_stream_t stream)
{
    JABD_REQUIRE(A >= 0 && B >= 0, JABD_EINVAL, "jaccard: negative size");
    if (A == 0 || B == 0) return JABD_OK;
    JABD_REQUIRE(box_a && box_b && out, JABD_EINVAL, "jaccard: null pointer");
    JABD_REQUIRE(aligned_to(box_a, 16) && aligned_to(box_b, 16) && aligned_to(out, 4), JABD_EALIGN,
                 "jaccard: boxes need 16-byte alignment");
    JABD_REQUIRE((A + 15) / 16 <= 65535, JABD_EINVAL, "jaccard: A=%lld too large (max 1048560 rows)", (long long)A);
    const dim3 grid(blocks_for(B, 256), (unsigned)((A + 15) / 16));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (inter_only)
        jaccard_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(box_a), A,
                                                   reinterpret_cast<const float4 *>(box_b), B, out);
    else
        jaccard_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(box_a), A,
                                                    reinterpret_cast<const float4 *>(box_b), B, out);
    JABD_LAUNCH_CHECK("jaccard_kernel");
    return JABD_OK;
}

int jabd_jaccard(const float *box_a, int64_t A, const float *box_b, int64_t B, float *out, jabd_stream_t stream)
{
    return jaccard_common(false, box_a, A, box_b, B, out, stream);
}

int jabd_intersect(const float *box_a, int64_t A, const float *box_b, int64_t B, float *out, jabd_stream_t stream)
{
    return jaccard_common(true, box_a, A, box_b, B, out, stream);
}

int jabd_encode(const float *matched, const float *priors, int64_t n, float var0, float var1, float *out, jabd_stream_t stream)
{
    JABD_REQUIRE(n >= 0, JABD_EINVAL, "encode: negative n");
    if (n == 0) return JABD_OK;
    JABD_REQUIRE(matched && priors && out, JABD_EINVAL, "encode: null pointer");
    JABD_REQUIRE(aligned_to(matched, 16) && aligned_to(priors, 16) && aligned_to(out, 16), JABD_EALIGN,
                 "encode: 16-byte alignment required");
    encode_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(matched), reinterpret_cast<const float4 *>(priors), n, var0, var1,
        reinterpret_cast<float4 *>(out));
    JABD_LAUNCH_CHECK("encode_kernel");
    return JABD_OK;
}

static int landm_common(const char *name, bool enc, const float *in, const float *priors, int64_t P, int batch, float var0,
                        float *out, jabd_stream_t stream)
{
    JABD_REQUIRE(P >= 0 && batch >= 0 && batch <= 65535, JABD_EINVAL, "%s: bad size (P=%lld batch=%d)", name, (long long)P, batch);
    if (P == 0 || batch == 0) return JABD_OK;
    JABD_REQUIRE(in && priors && out, JABD_EINVAL, "%s: null pointer", name);
    JABD_REQUIRE(aligned_to(in, 8) && aligned_to(out, 8) && aligned_to(priors, 16), JABD_EALIGN,
                 "%s: landmarks need 8-byte, priors 16-byte alignment", name);
    const dim3 grid(blocks_for(P * 5, 256), (unsigned)batch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (enc)
        landm_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const float2 *>(in), reinterpret_cast<const float4 *>(priors), P,
                                                 var0, reinterpret_cast<float2 *>(out));
    else
        landm_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const float2 *>(in), reinterpret_cast<const float4 *>(priors), P,
                                                  var0, reinterpret_cast<float2 *>(out));
    JABD_LAUNCH_CHECK(name);
    return JABD_OK;
}

int jabd_encode_landm(const float *matched, const float *priors, int64_t n, float var0, float *out, jabd_stream_t stream)
{
    return landm_common("encode_landm", true, matched, priors, n, 1, var0, out, stream);
}

int jabd_decode_landm(const float *pre, const float *priors, int64_t P, int batch, float var0, float *out, jabd_stream_t stream)
{
    return landm_common("decode_landm", false, pre, priors, P, batch, var0, out, stream);
}

int jabd_correct_boxes(float *dets, const int *counts, const double *post, int B, int K, int letterbox, int to_pixels,
                       jabd_stream_t stream)
{
    JABD_REQUIRE(B >= 0 && K >= 0, JABD_EINVAL, "correct_boxes: negative size");
    if (B == 0 || K == 0 || (!letterbox && !to_pixels)) return JABD_OK;
    JABD_REQUIRE(dets && post, JABD_EINVAL, "correct_boxes: null pointer");
    JABD_REQUIRE(aligned_to(post, 8) && aligned_to(dets, 4), JABD_EALIGN, "correct_boxes: post must be 8-byte aligned");
    correct_boxes_kernel<<<blocks_for((long long)B * K, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dets, counts, post, B, K,
                                                                                                        letterbox, to_pixels);
    JABD_LAUNCH_CHECK("correct_boxes_kernel");
    return JABD_OK;
}

int jabd_decode(const float *loc, const float *priors, int64_t P, int batch, float var0, float var1, float *out,
                jabd_stream_t stream)
{
    JABD_REQUIRE(P >= 0 && batch >= 0 && batch <= 65535, JABD_EINVAL, "decode: bad size (P=%lld batch=%d)", (long long)P, batch);
    if (P == 0 || batch == 0) return JABD_OK;
    JABD_REQUIRE(loc && priors && out, JABD_EINVAL, "decode: null pointer");
    JABD_REQUIRE(aligned_to(loc, 16) && aligned_to(priors, 16) && aligned_to(out, 16), JABD_EALIGN,
                 "decode: 16-byte alignment required");
    const dim3 grid(blocks_for(P, 256), (unsigned)((batch + kDecImgs - 1) / kDecImgs));
    decode_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4 *>(loc),
                                                                       reinterpret_cast<const float4 *>(priors), P, batch, var0,
                                                                       var1, reinterpret_cast<float4 *>(out));
    JABD_LAUNCH_CHECK("decode_kernel");
    return JABD_OK;
}

} // extern "C"
