// iou_family.cuh -- the element-wise IoU / GIoU / DIoU / CIoU of two aligned box lists.
//
// Restates bbox_overlaps_{iou,giou,diou,ciou} of R/utils/box_utils.py:5-158 (the same four functions are repeated in
// R/nets/retinaface_training_DIOU.py:342-490): despite the [rows, cols] buffer they allocate, their torch.min / torch.max
// calls pair bboxes1[i] with bboxes2[i], so the result is one value per row.  The formulas are written once over a number
// type T: `float` evaluates them with the reference's operation order, one rounding per op; `Dual` carries, next to the
// same value, the derivative with respect to the four coordinates of the FIRST box, which is what autograd computes for
// IouLoss (R/nets/retinaface_training_DIOU.py:491-525): torch.min/max split the gradient on ties, clamp passes it inside
// the closed range, CIoU's alpha is a constant (torch.no_grad, :482-484).
#pragma once
#include "common.cuh"

namespace jabd {

enum IouKind { kIou = 1, kGiou = 2, kDiou = 3, kCiou = 4 };

struct Dual {
    float v;
    float d[4];
};

__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(const Dual &a) { return a.v; }

// ---- float: the reference's arithmetic
__device__ __forceinline__ float n_const(float, float c) { return c; }
__device__ __forceinline__ float n_add(float a, float b) { return fadd(a, b); }
__device__ __forceinline__ float n_sub(float a, float b) { return fsub(a, b); }
__device__ __forceinline__ float n_mul(float a, float b) { return fmul(a, b); }
__device__ __forceinline__ float n_div(float a, float b) { return fdiv(a, b); }
__device__ __forceinline__ float n_min(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }  // torch.min propagates NaN
__device__ __forceinline__ float n_max(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }
__device__ __forceinline__ float n_clamp(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float n_clamp_min(float x, float lo) { return x != x ? x : fmaxf(x, lo); }
__device__ __forceinline__ float n_atan(float x) { return atanf(x); }
__device__ __forceinline__ float n_detach(float x) { return x; }

// ---- Dual: same value, plus d/d(first box)
__device__ __forceinline__ Dual n_const(const Dual &, float c) { return Dual{c, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ Dual n_add(const Dual &a, const Dual &b)
{
    Dual r{fadd(a.v, b.v), {}};
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] + b.d[k];
    return r;
}
__device__ __forceinline__ Dual n_sub(const Dual &a, const Dual &b)
{
    Dual r{fsub(a.v, b.v), {}};
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] - b.d[k];
    return r;
}
__device__ __forceinline__ Dual n_mul(const Dual &a, const Dual &b)
{
    Dual r{fmul(a.v, b.v), {}};
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
__device__ __forceinline__ Dual n_div(const Dual &a, const Dual &b)
{
    Dual r{fdiv(a.v, b.v), {}};
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v;
    return r;
}
__device__ __forceinline__ Dual n_min(const Dual &a, const Dual &b)
{
    Dual r{n_min(a.v, b.v), {}};
    const float wa = a.v < b.v ? 1.0f : (a.v == b.v ? 0.5f : 0.0f), wb = 1.0f - wa; // ties: half each
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = wa * a.d[k] + wb * b.d[k];
    return r;
}
__device__ __forceinline__ Dual n_max(const Dual &a, const Dual &b)
{
    Dual r{n_max(a.v, b.v), {}};
    const float wa = a.v > b.v ? 1.0f : (a.v == b.v ? 0.5f : 0.0f), wb = 1.0f - wa;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = wa * a.d[k] + wb * b.d[k];
    return r;
}
__device__ __forceinline__ Dual n_clamp(const Dual &x, float lo, float hi)
{
    Dual r{n_clamp(x.v, lo, hi), {}};
    const float pass = (x.v >= lo && x.v <= hi) ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = pass * x.d[k];
    return r;
}
__device__ __forceinline__ Dual n_clamp_min(const Dual &x, float lo)
{
    Dual r{n_clamp_min(x.v, lo), {}};
    const float pass = x.v >= lo ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = pass * x.d[k];
    return r;
}
__device__ __forceinline__ Dual n_atan(const Dual &x)
{
    Dual r{atanf(x.v), {}};
    const float g = 1.0f / (1.0f + x.v * x.v);
#pragma unroll
    for (int k = 0; k < 4; ++k) r.d[k] = g * x.d[k];
    return r;
}
__device__ __forceinline__ Dual n_detach(const Dual &x) { return Dual{x.v, {0.f, 0.f, 0.f, 0.f}}; }

template <typename T>
struct Box {
    T x1, y1, x2, y2;
};

// one overlap value, clamped like the reference
template <typename T>
__device__ __forceinline__ T iou_family(int kind, const Box<T> &a, const Box<T> &b)
{
    const T one = n_const(a.x1, 1.0f);
    const T w1 = n_sub(a.x2, a.x1), h1 = n_sub(a.y2, a.y1), w2 = n_sub(b.x2, b.x1), h2 = n_sub(b.y2, b.y1);
    const T area1 = n_mul(w1, h1), area2 = n_mul(w2, h2);
    const T iw = n_clamp_min(n_sub(n_min(a.x2, b.x2), n_max(a.x1, b.x1)), 0.0f);
    const T ih = n_clamp_min(n_sub(n_min(a.y2, b.y2), n_max(a.y1, b.y1)), 0.0f);
    const T inter = n_mul(iw, ih);
    const T uni = n_sub(n_add(area1, area2), inter);
    if (kind == kIou) return n_clamp(n_div(inter, uni), 0.0f, 1.0f);                       // box_utils.py:96-119
    const T ow = n_clamp_min(n_sub(n_max(a.x2, b.x2), n_min(a.x1, b.x1)), 0.0f);
    const T oh = n_clamp_min(n_sub(n_max(a.y2, b.y2), n_min(a.y1, b.y1)), 0.0f);
    if (kind == kGiou) {                                                                   // :121-158
        const T closure = n_mul(ow, oh);
        return n_clamp(n_sub(n_div(inter, uni), n_div(n_sub(closure, uni), closure)), -1.0f, 1.0f);
    }
    const T half = n_const(a.x1, 0.5f);
    const T cx1 = n_mul(n_add(a.x2, a.x1), half), cy1 = n_mul(n_add(a.y2, a.y1), half);    // "/ 2" is exact
    const T cx2 = n_mul(n_add(b.x2, b.x1), half), cy2 = n_mul(n_add(b.y2, b.y1), half);
    const T dx = n_sub(cx2, cx1), dy = n_sub(cy2, cy1);
    const T inter_diag = n_add(n_mul(dx, dx), n_mul(dy, dy));
    const T outer_diag = n_add(n_mul(ow, ow), n_mul(oh, oh));
    if (kind == kDiou)                                                                     // :5-46
        return n_clamp(n_sub(n_div(inter, uni), n_div(inter_diag, outer_diag)), -1.0f, 1.0f);
    // CIoU, :48-94
    const T u = n_div(inter_diag, outer_diag);
    const T iou = n_div(inter, uni);
    const T da = n_sub(n_atan(n_div(w2, h2)), n_atan(n_div(w1, h1)));
    const T v = n_mul(n_const(a.x1, (float)(4.0 / (3.141592653589793 * 3.141592653589793))), n_mul(da, da));
    const T S = n_sub(one, n_detach(iou));
    const T alpha = n_div(n_detach(v), n_add(S, n_detach(v)));
    return n_clamp(n_sub(iou, n_add(u, n_mul(alpha, v))), -1.0f, 1.0f);
}

__device__ __forceinline__ Box<float> box_of(float4 b) { return Box<float>{b.x, b.y, b.z, b.w}; }
__device__ __forceinline__ Box<Dual> box_seed(float4 b) // d/d(x1, y1, x2, y2) of the box itself
{
    return Box<Dual>{Dual{b.x, {1.f, 0.f, 0.f, 0.f}}, Dual{b.y, {0.f, 1.f, 0.f, 0.f}}, Dual{b.z, {0.f, 0.f, 1.f, 0.f}},
                     Dual{b.w, {0.f, 0.f, 0.f, 1.f}}};
}
__device__ __forceinline__ Box<Dual> box_const(float4 b)
{
    return Box<Dual>{Dual{b.x, {0.f, 0.f, 0.f, 0.f}}, Dual{b.y, {0.f, 0.f, 0.f, 0.f}}, Dual{b.z, {0.f, 0.f, 0.f, 0.f}},
                     Dual{b.w, {0.f, 0.f, 0.f, 0.f}}};
}

// d(1 - overlap)/d(loc) for loc decoded against prior p (decode, R/utils/utils_bbox.py:29-34): chain rule through
// cx = p.x + l.x*var0*p.z, w = p.z*exp(l.z*var1), x1 = cx - w/2, x2 = w + x1
__device__ __forceinline__ float4 iou_loss_grad(int kind, float4 loc, float4 prior, float4 target, float var0, float var1)
{
    const float4 dbox = decode_box(loc, prior, var0, var1);
    const Dual ov = iou_family<Dual>(kind, box_seed(dbox), box_const(target));
    const float gx1 = -ov.d[0], gy1 = -ov.d[1], gx2 = -ov.d[2], gy2 = -ov.d[3];
    const float w = dbox.z - dbox.x, h = dbox.w - dbox.y;
    float4 g;
    g.x = (gx1 + gx2) * var0 * prior.z;
    g.y = (gy1 + gy2) * var0 * prior.w;
    g.z = 0.5f * (gx2 - gx1) * w * var1;
    g.w = 0.5f * (gy2 - gy1) * h * var1;
    return g;
}

} // namespace jabd
