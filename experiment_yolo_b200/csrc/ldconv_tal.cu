// ldconv_tal.cu -- the dense part of the task-aligned assigner of the DEAL-YOLO training criterion as three kernels (sm_100a).
//
// Restates /root/reference/ultralytics/utils/tal.py:13-290 (TaskAlignedAssigner: select_candidates_in_gts :226-243,
// get_box_metrics :98-122 with bbox_iou(CIoU=True) utils/metrics.py:75-128, select_topk_candidates :124-157,
// select_highest_overlaps :245-272, the normalisation terms :83-88).  The reference runs ~60 element-wise / masked-index passes
// over (batch, n_gt, n_anchors) tensors (68.8 M elements at batch 128, 16 boxes, 33600 anchors: 22 ms of device time in
// eager PyTorch, benchmarks/profile_loss.py); here each element is touched by
//   tal_metric_kernel    thread = (image, anchor): for every ground truth the in-box test, CIoU, score^alpha * CIoU^beta
//   (torch.topk picks the k best anchors per ground truth -- library call, between the kernels)
//   tal_topk_mask_kernel thread = (image, gt, j): marks the j-th top anchor of a valid gt when it lies inside the box
//   tal_resolve_kernel   thread = (image, anchor): an anchor claimed by several gts goes to the gt with the highest overlap
//                        (first index on ties, also when that gt did not claim it: the reference's argmax quirk), foreground
//                        flag, assigned gt, and the per-gt maxima of metric / overlap over its positives (atomicMax)
// Everything is fp32 like the reference's assigner; experiment_yolo_b200/loss.py keeps the same arithmetic in plain PyTorch
// (the statement pinned against the reference fixtures) and tests/test_gpu_loss.py compares the two on the GPU.
#include "common.cuh"

namespace ldc {

__device__ __forceinline__ float tal_ciou(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2, float by2)
{
    const float eps = 1e-7f;
    const float w1 = ax2 - ax1, h1 = ay2 - ay1 + eps, w2 = bx2 - bx1, h2 = by2 - by1 + eps;
    const float inter = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1), 0.f) * fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1), 0.f);
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(ax2, bx2) - fminf(ax1, bx1), ch = fmaxf(ay2, by2) - fminf(ay1, by1);
    const float c2 = cw * cw + ch * ch + eps;
    const float dx = bx1 + bx2 - ax1 - ax2, dy = by1 + by2 - ay1 - ay2;
    const float rho2 = (dx * dx + dy * dy) * 0.25f;
    const float da = atanf(w2 / h2) - atanf(w1 / h1);
    const float v = 0.40528473456935109f * da * da;      // 4 / pi^2
    const float alpha = v / (v - iou + (1.f + eps));
    return iou - (rho2 / c2 + v * alpha);
}

// Dense, gradient-free decode of the raw head outputs for the assigner (utils/loss.py:347-354 `bbox_decode` = DFL softmax-expectation
// (nn/modules/block.py:37-56) + dist2bbox in grid units, utils/tal.py:309-318; and `pred_scores.sigmoid()`, loss.py:388-394): one
// thread per anchor row of x (rows, 4*16 + nc).  The criterion itself differentiates only through the (few) foreground rows, so the
// reference's dense softmax / matmul / their backward over all 33600 anchors are not needed.
template <typename T>
__global__ void __launch_bounds__(256)
head_decode_rows_kernel(const T* __restrict__ x, const float* __restrict__ anc, float* __restrict__ boxes, float* __restrict__ scores,
                        long long rows, int na, int nc)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int no = 64 + nc;
    const T* xr = x + r * no;
    const int a = (int)(r % na);
    const float ax = anc[2 * a], ay = anc[2 * a + 1];
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = Elem<T>::to_f(xr[side * 16 + k]);
        float mx = v[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) mx = fmaxf(mx, v[k]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float e = expf(v[k] - mx);
            den += e;
            num = fmaf(e, (float)k, num);
        }
        d[side] = num / den;
    }
    reinterpret_cast<float4*>(boxes)[r] = make_float4(ax - d[0], ay - d[1], ax + d[2], ay + d[3]);
    for (int c = 0; c < nc; ++c) scores[r * nc + c] = 1.f / (1.f + expf(-Elem<T>::to_f(xr[64 + c])));
}

constexpr int kTalMaxGt = 256;

__global__ void __launch_bounds__(256)
tal_metric_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, const float* __restrict__ anc,
                  const int* __restrict__ labels, const float* __restrict__ gtb, const unsigned char* __restrict__ valid,
                  float* __restrict__ align, float* __restrict__ overlaps, int na, int n, int nc, float alpha, float beta, float eps)
{
    __shared__ float4 s_box[kTalMaxGt];
    __shared__ int s_lab[kTalMaxGt];      // -1: padded gt
    const int b = blockIdx.y;
    for (int g = threadIdx.x; g < n; g += blockDim.x) {
        s_box[g] = reinterpret_cast<const float4*>(gtb)[(size_t)b * n + g];
        const int l = labels[(size_t)b * n + g];
        s_lab[g] = valid[(size_t)b * n + g] ? min(max(l, 0), nc - 1) : -1;
    }
    __syncthreads();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= na) return;
    const float4 pb = reinterpret_cast<const float4*>(boxes)[(size_t)b * na + a];
    const float2 ap = reinterpret_cast<const float2*>(anc)[a];
    const float* sc = scores + ((size_t)b * na + a) * nc;
    for (int g = 0; g < n; ++g) {
        float al = 0.f, ov = 0.f;
        const int l = s_lab[g];
        if (l >= 0) {
            const float4 gb = s_box[g];
            const float d = fminf(fminf(ap.x - gb.x, ap.y - gb.y), fminf(gb.z - ap.x, gb.w - ap.y));
            if (d > eps) {
                ov = fmaxf(tal_ciou(gb.x, gb.y, gb.z, gb.w, pb.x, pb.y, pb.z, pb.w), 0.f);
                const float s = __ldg(sc + l);
                al = (alpha == 0.5f ? sqrtf(s) : powf(s, alpha)) * powf(ov, beta);
            }
        }
        const size_t o = ((size_t)b * n + g) * na + a;
        align[o] = al;
        overlaps[o] = ov;
    }
}

__global__ void __launch_bounds__(256)
tal_topk_mask_kernel(const long long* __restrict__ idx, const float* __restrict__ anc, const float* __restrict__ gtb,
                     const unsigned char* __restrict__ valid, unsigned char* __restrict__ mask, int na, int n, int k, long long total,
                     float eps)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long bg = t / k;      // image * n + gt
    if (!valid[bg]) return;
    const long long a = idx[t];
    const float4 gb = reinterpret_cast<const float4*>(gtb)[bg];
    const float2 ap = reinterpret_cast<const float2*>(anc)[a];
    const float d = fminf(fminf(ap.x - gb.x, ap.y - gb.y), fminf(gb.z - ap.x, gb.w - ap.y));
    if (d > eps) mask[bg * na + a] = 1;      // top-k indices of one gt are distinct: the reference's count == 1
}

__global__ void __launch_bounds__(256)
tal_resolve_kernel(const unsigned char* __restrict__ mask, const float* __restrict__ align, const float* __restrict__ overlaps,
                   unsigned char* __restrict__ fg, long long* __restrict__ gt_idx, float* __restrict__ align_sel,
                   float* __restrict__ pos_align, float* __restrict__ pos_over, int na, int n)
{
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= na) return;
    int cnt = 0, first = 0, best_g = 0;
    float best = -1.f;
    for (int g = 0; g < n; ++g) {
        const size_t o = ((size_t)b * n + g) * na + a;
        const float ov = overlaps[o];
        if (ov > best) { best = ov; best_g = g; }
        if (mask[o]) { if (cnt == 0) first = g; ++cnt; }
    }
    const int gs = cnt > 1 ? best_g : first;
    const size_t o = ((size_t)b * n + gs) * na + a;
    const float al = cnt ? align[o] : 0.f, ov = cnt ? overlaps[o] : 0.f;
    fg[(size_t)b * na + a] = cnt ? 1 : 0;
    gt_idx[(size_t)b * na + a] = gs;
    align_sel[(size_t)b * na + a] = al;
    if (cnt) {      // non-negative floats order like their bit patterns
        atomicMax(reinterpret_cast<int*>(pos_align) + (size_t)b * n + gs, __float_as_int(al));
        atomicMax(reinterpret_cast<int*>(pos_over) + (size_t)b * n + gs, __float_as_int(ov));
    }
}

}  // namespace ldc

using namespace ldc;

LDC_API int ldconv_tal_metric(const float* scores, const float* boxes, const float* anchors, const int32_t* gt_labels,
                              const float* gt_boxes, const unsigned char* gt_valid, float* align, float* overlaps, int B, int na,
                              int n, int nc, float alpha, float beta, float eps, void* stream)
{
    LDC_REQUIRE(scores && boxes && anchors && gt_labels && gt_boxes && gt_valid && align && overlaps, "ldconv_tal_metric: null pointer");
    LDC_REQUIRE(B >= 0 && na >= 1 && n >= 1 && n <= kTalMaxGt && nc >= 1 && B <= 65535, "ldconv_tal_metric: bad dims (n_gt <= 256)");
    LDC_REQUIRE(aligned16(boxes) && aligned16(gt_boxes) && (reinterpret_cast<uintptr_t>(anchors) & 7u) == 0, "ldconv_tal_metric: alignment");
    if (B == 0) return LDCONV_OK;
    tal_metric_kernel<<<dim3(cdiv(na, 256), B), 256, 0, (cudaStream_t)stream>>>(scores, boxes, anchors, gt_labels, gt_boxes, gt_valid,
                                                                                align, overlaps, na, n, nc, alpha, beta, eps);
    LDC_LAUNCH_CHECK("tal_metric_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_tal_assign(const long long* topk_idx, const float* anchors, const float* gt_boxes, const unsigned char* gt_valid,
                              const float* align, const float* overlaps, unsigned char* mask_ws, unsigned char* fg, long long* gt_idx,
                              float* align_sel, float* pos_align, float* pos_over, int B, int na, int n, int k, float eps,
                              void* stream)
{
    LDC_REQUIRE(topk_idx && anchors && gt_boxes && gt_valid && align && overlaps && mask_ws && fg && gt_idx && align_sel && pos_align &&
                    pos_over, "ldconv_tal_assign: null pointer");
    LDC_REQUIRE(B >= 0 && na >= 1 && n >= 1 && k >= 1 && B <= 65535, "ldconv_tal_assign: bad dims");
    if (B == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LDC_CUDA(cudaMemsetAsync(mask_ws, 0, (size_t)B * n * na, st));
    LDC_CUDA(cudaMemsetAsync(pos_align, 0, (size_t)B * n * sizeof(float), st));
    LDC_CUDA(cudaMemsetAsync(pos_over, 0, (size_t)B * n * sizeof(float), st));
    const long long total = (long long)B * n * k;
    tal_topk_mask_kernel<<<cdiv(total, 256), 256, 0, st>>>(topk_idx, anchors, gt_boxes, gt_valid, mask_ws, na, n, k, total, eps);
    LDC_LAUNCH_CHECK("tal_topk_mask_kernel");
    tal_resolve_kernel<<<dim3(cdiv(na, 256), B), 256, 0, st>>>(mask_ws, align, overlaps, fg, gt_idx, align_sel, pos_align, pos_over, na, n);
    LDC_LAUNCH_CHECK("tal_resolve_kernel");
    return LDCONV_OK;
}

// x (rows = b * na, 64 + nc) bf16 / fp32 raw head rows (4 sides x 16 DFL bins, then nc class logits); anc (na, 2) fp32 anchor centres in
// grid units -> boxes (rows, 4) fp32 xyxy in grid units, scores (rows, nc) fp32 = sigmoid(logits).  No gradient: assigner inputs.
LDC_API int ldconv_head_decode_rows(const void* x, const float* anc, float* boxes, float* scores, long long rows, int na, int nc,
                                    int reg_max, int dtype, void* stream)
{
    using namespace ldc;
    LDC_REQUIRE(x && anc && boxes && scores && rows >= 0 && na >= 1 && nc >= 1, "ldconv_head_decode_rows: bad arguments");
    LDC_REQUIRE(reg_max == 16, "ldconv_head_decode_rows: reg_max %d not supported (16)", reg_max);
    LDC_REQUIRE(dtype == LDCONV_BF16 || dtype == LDCONV_F32, "ldconv_head_decode_rows: unsupported dtype %d", dtype);
    LDC_REQUIRE(aligned16(boxes), "ldconv_head_decode_rows: boxes must be 16-byte aligned");
    if (rows == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_BF16)
        head_decode_rows_kernel<__nv_bfloat16><<<cdiv(rows, 256), 256, 0, st>>>((const __nv_bfloat16*)x, anc, boxes, scores, rows, na, nc);
    else
        head_decode_rows_kernel<float><<<cdiv(rows, 256), 256, 0, st>>>((const float*)x, anc, boxes, scores, rows, na, nc);
    LDC_LAUNCH_CHECK("head_decode_rows_kernel");
    return LDCONV_OK;
}
