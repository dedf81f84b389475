// iou_loss.cu -- element-wise IoU-family overlaps and the IouLoss regression loss.
//
//   jabd_bbox_overlaps_family  bbox_overlaps_{iou,giou,diou,ciou}(bboxes1, bboxes2), R/utils/box_utils.py:5-158
//   jabd_iou_loss_forward      IouLoss.forward, R/nets/retinaface_training_DIOU.py:491-525: sum (or mean) over rows of
//                              1 - overlap(decode(loc_p, priors) | loc_p, loc_t)
//   jabd_iou_loss_backward     its gradient with respect to loc_p (what autograd derives in the reference)
// The formulas live in iou_family.cuh; the MultiBox loss (loss.cu) uses the same functions for its DIoU variant.
#include "iou_family.cuh"

namespace jabd {

__global__ void __launch_bounds__(256) overlaps_family_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, long long n,
                                                              int kind, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = iou_family<float>(kind, box_of(__ldg(a + i)), box_of(__ldg(b + i)));
}

// one CTA: per-thread fp32 partial sums in index order, fp64 tree over the CTA (fixed order, deterministic)
__global__ void __launch_bounds__(1024) iou_loss_forward_kernel(const float4 *__restrict__ loc_p, const float4 *__restrict__ loc_t,
                                                                const float4 *__restrict__ priors, long long n, float var0, float var1,
                                                                int kind, int size_sum, float *__restrict__ per_row,
                                                                float *__restrict__ loss)
{
    __shared__ double s_sum[32];
    float acc = 0.0f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        float4 box = __ldg(loc_p + i);
        if (priors) box = decode_box(box, __ldg(priors + i), var0, var1); // pred_mode == 'Center'
        const float l = fsub(1.0f, iou_family<float>(kind, box_of(box), box_of(__ldg(loc_t + i))));
        if (per_row) per_row[i] = l;
        acc = fadd(acc, l);
    }
    double d = acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
    if (lane_id() == 0) s_sum[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_sum[w];
        float r = (float)t;
        if (!size_sum) r = fdiv(r, (float)n); // loss / num
        loss[0] = r;
    }
}

__global__ void __launch_bounds__(256) iou_loss_backward_kernel(const float4 *__restrict__ loc_p, const float4 *__restrict__ loc_t,
                                                                const float4 *__restrict__ priors, long long n, float var0, float var1,
                                                                int kind, int size_sum, const float *__restrict__ grad_loss,
                                                                float4 *__restrict__ g_loc)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = __ldg(grad_loss);
    if (!size_sum) g = g / (float)n;
    float4 r;
    if (priors) {
        r = iou_loss_grad(kind, __ldg(loc_p + i), __ldg(priors + i), __ldg(loc_t + i), var0, var1);
    } else {
        const Dual ov = iou_family<Dual>(kind, box_seed(__ldg(loc_p + i)), box_const(__ldg(loc_t + i)));
        r = make_float4(-ov.d[0], -ov.d[1], -ov.d[2], -ov.d[3]);
    }
    g_loc[i] = make_float4(g * r.x, g * r.y, g * r.z, g * r.w);
}

static int check_kind(int kind, const char *who)
{
    JABD_REQUIRE(kind >= kIou && kind <= kCiou, JABD_EINVAL, "%s: kind must be 1 (iou), 2 (giou), 3 (diou) or 4 (ciou)", who);
    return JABD_OK;
}

} // namespace jabd

using namespace jabd;

extern "C" {

int jabd_bbox_overlaps_family(const float *boxes1, const float *boxes2, int64_t N, int kind, float *out, jabd_stream_t stream)
{
    JABD_REQUIRE(N >= 0, JABD_EINVAL, "bbox_overlaps_family: negative size");
    int rc = check_kind(kind, "bbox_overlaps_family");
    if (rc != JABD_OK || N == 0) return rc;
    JABD_REQUIRE(boxes1 && boxes2 && out, JABD_EINVAL, "bbox_overlaps_family: null pointer");
    JABD_REQUIRE(aligned_to(boxes1, 16) && aligned_to(boxes2, 16) && aligned_to(out, 4), JABD_EALIGN,
                 "bbox_overlaps_family: boxes need 16-byte alignment");
    overlaps_family_kernel<<<(unsigned)((N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(boxes1), reinterpret_cast<const float4 *>(boxes2), (long long)N, kind, out);
    JABD_LAUNCH_CHECK("overlaps_family_kernel");
    return JABD_OK;
}

int jabd_iou_loss_forward(const float *loc_p, const float *loc_t, const float *priors, int64_t N, float var0, float var1, int kind,
                          int size_sum, float *per_row, float *loss, jabd_stream_t stream)
{
    JABD_REQUIRE(N >= 0, JABD_EINVAL, "iou_loss: negative size");
    int rc = check_kind(kind, "iou_loss");
    if (rc != JABD_OK) return rc;
    JABD_REQUIRE(loss && ((loc_p && loc_t) || N == 0), JABD_EINVAL, "iou_loss: null pointer");
    JABD_REQUIRE(aligned_to(loc_p, 16) && aligned_to(loc_t, 16) && aligned_to(priors, 16), JABD_EALIGN, "iou_loss: boxes need 16-byte alignment");
    iou_loss_forward_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(loc_p), reinterpret_cast<const float4 *>(loc_t), reinterpret_cast<const float4 *>(priors),
        (long long)N, var0, var1, kind, size_sum, per_row, loss);
    JABD_LAUNCH_CHECK("iou_loss_forward_kernel");
    return JABD_OK;
}

int jabd_iou_loss_backward(const float *loc_p, const float *loc_t, const float *priors, int64_t N, float var0, float var1, int kind,
                           int size_sum, const float *grad_loss, float *g_loc, jabd_stream_t stream)
{
    JABD_REQUIRE(N >= 0, JABD_EINVAL, "iou_loss_backward: negative size");
    int rc = check_kind(kind, "iou_loss_backward");
    if (rc != JABD_OK || N == 0) return rc;
    JABD_REQUIRE(loc_p && loc_t && grad_loss && g_loc, JABD_EINVAL, "iou_loss_backward: null pointer");
    JABD_REQUIRE(aligned_to(loc_p, 16) && aligned_to(loc_t, 16) && aligned_to(priors, 16) && aligned_to(g_loc, 16), JABD_EALIGN,
                 "iou_loss_backward: boxes need 16-byte alignment");
    iou_loss_backward_kernel<<<(unsigned)((N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(loc_p), reinterpret_cast<const float4 *>(loc_t), reinterpret_cast<const float4 *>(priors),
        (long long)N, var0, var1, kind, size_sum, grad_loss, reinterpret_cast<float4 *>(g_loc));
    JABD_LAUNCH_CHECK("iou_loss_backward_kernel");
    return JABD_OK;
}

} // extern "C"
