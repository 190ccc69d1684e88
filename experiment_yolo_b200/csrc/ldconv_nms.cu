// ldconv_nms.cu -- the reference's post-processing on the device: confidence filter + class-offset soft-NMS (sm_100a).
//
// Restates /root/reference/ultralytics/utils/ops.py:292-427 (`non_max_suppression`, the path the predictor / validator take:
// single label, no masks, not rotated, classes=None, labels=()) INCLUDING the fork's own `soft_nms` (ops.py:260-290), which is
// what line 407 calls instead of torchvision.ops.nms, with its quirks:
//   * candidates keep their ANCHOR order (ops.py:363,385); the first box kept is the first candidate, not the best one;
//   * one step (ops.py:265-288): keep order[0]; IoU (bbox_iou_for_nms, ops.py:188-202: h + eps, union + eps) of that box against
//     the rest on class-offset boxes (+ cls * max_wh, ops.py:399,405); scores of boxes with IoU > iou_thres are multiplied IN PLACE
//     by exp(-IoU^2 / 0.5) (so the returned confidences are the decayed ones: `scores` is a view of x[:, 4]); boxes whose score
//     is <= 0.25 (soft_nms' own default, not conf_thres) leave; the best remaining box is swapped to the front;
//   * the loop runs `while order.numel() > 1` and its `numel() == 1` branch is unreachable: the last surviving box is NEVER kept,
//     an image with a single candidate returns nothing, and with exactly two boxes left no decay is applied (0-d squeeze);
//   * the kept list is cut to max_det afterwards (ops.py:408).
// One CTA per image: stable compaction of the candidates (anchor order), then the sequential loop with block-wide passes
// (IoU / decay / survivor flags -> stable compaction into the other order buffer -> arg-max, first occurrence).  Replaces a
// 43 MB device->host copy per 64 images by 0.46 MB and ~64 x (hundreds) of host-launched torch ops by one launch.
#include "common.cuh"

namespace ldc {

// The launch runs on the predictor's download stream beside the next batch's forward, whose persistent kernels it displaces from
// the SMs it occupies: the end-to-end rate follows the launch's DURATION (measured, 64 images per batch: 128 threads per CTA 20 672
// images/s, 512 threads 21 725, 1024 threads 21 859), so the CTA is as wide as the hardware allows.
static constexpr int kNmsThreads = 1024;

__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kNmsThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kNmsThreads / 32) s_warp[lane] = w;      // inclusive totals of the warps
    }
    __syncthreads();
    total = s_warp[kNmsThreads / 32 - 1];
    const int base = warp == 0 ? 0 : s_warp[warp - 1];
    __syncthreads();
    return base + inc - v;
}

// y (B, 4 + nc, A) bf16 or fp32: rows xywh (pixels), then class scores.  Workspace per image (cap candidates):
// box float4[cap] (xyxy, no class offset), score float[cap], cls int[cap], order int[2][cap].
template <typename T>
__global__ void __launch_bounds__(kNmsThreads)
nms_soft_kernel(const T* __restrict__ y, int A, int nc, float conf_thres, float iou_thres, float max_wh, int agnostic, int max_det,
                int cap, float4* __restrict__ ws_box, float* __restrict__ ws_score, int* __restrict__ ws_cls,
                int* __restrict__ ws_order, float* __restrict__ out, int* __restrict__ out_count)
{
    __shared__ int s_warp[kNmsThreads / 32];
    __shared__ float s_best[kNmsThreads / 32];
    __shared__ int s_bpos[kNmsThreads / 32];
    __shared__ int s_m, s_nk, s_stop;
    const int b = blockIdx.x, tid = threadIdx.x;
    const T* yb = y + (size_t)b * (4 + nc) * A;
    float4* box = ws_box + (size_t)b * cap;
    float* score = ws_score + (size_t)b * cap;
    int* cls = ws_cls + (size_t)b * cap;
    int* order0 = ws_order + (size_t)b * 2 * cap;
    int* order1 = order0 + cap;
    float* ob = out + (size_t)b * max_det * 6;

    // ---- candidates: best class score > conf_thres, in anchor order (ops.py:347,363,384-385) ------------------------------------
    // Each warp owns a contiguous range of anchors: pass 1 counts its candidates (ballot + popc, no block barrier inside the
    // loop), one block scan over the warp totals gives every warp its first slot, pass 2 writes them -- a stable compaction
    // with three block barriers in total instead of three per 256 anchors.
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = kNmsThreads / 32;
    const int per_warp = ((A + NWARP - 1) / NWARP + 31) & ~31;
    const int a_begin = warp * per_warp, a_end = min(A, a_begin + per_warp);
    auto best_of = [&](int a, float& best, int& bj) {
        best = -1.f;
        bj = 0;
        if (a < a_end) {
            for (int j = 0; j < nc; ++j) {      // torch.max: first maximal index
                const float v = Elem<T>::to_f(yb[(size_t)(4 + j) * A + a]);
                if (v > best) { best = v; bj = j; }
            }
        }
        return (a < a_end && best > conf_thres) ? 1 : 0;
    };
    int wcount = 0;
    for (int a0 = a_begin; a0 < a_end; a0 += 32) {
        float best;
        int bj;
        wcount += __popc(__ballot_sync(0xffffffffu, best_of(a0 + lane, best, bj)));
    }
    int n;
    const int wbase = block_excl_scan(lane == 0 ? wcount : 0, s_warp, n);      // lane 0 of every warp: exclusive sum of the warp totals
    int pos0 = __shfl_sync(0xffffffffu, wbase, 0);
    for (int a0 = a_begin; a0 < a_end; a0 += 32) {
        const int a = a0 + lane;
        float best;
        int bj;
        const int flag = best_of(a, best, bj);
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        const int pos = pos0 + __popc(bal & ((1u << lane) - 1u));
        if (flag && pos < cap) {
            const float cx = Elem<T>::to_f(yb[a]), cy = Elem<T>::to_f(yb[(size_t)A + a]);
            const float w = Elem<T>::to_f(yb[(size_t)2 * A + a]), h = Elem<T>::to_f(yb[(size_t)3 * A + a]);
            const float dw = w / 2, dh = h / 2;      // xywh2xyxy (ops.py:509-516)
            box[pos] = make_float4(__fsub_rn(cx, dw), __fsub_rn(cy, dh), __fadd_rn(cx, dw), __fadd_rn(cy, dh));
            score[pos] = best;
            cls[pos] = bj;
            order0[pos] = pos;
        }
        pos0 += __popc(bal);
    }
    if (n > cap) {      // more candidates than max_nms: the reference sorts by confidence and truncates (ops.py:395-396); not covered
        if (tid == 0) out_count[b] = -n;
        return;
    }
    if (tid == 0) { s_m = n; s_nk = 0; s_stop = 0; }
    __syncthreads();

    // ---- soft_nms (ops.py:260-290) -------------------------------------------------------------------------------------------------
    int* cur = order0;
    int* nxt = order1;
    const float off_scale = agnostic ? 0.f : max_wh;
    while (true) {
        const int m = s_m;
        if (m <= 1 || s_stop) break;
        const int i = cur[0];
        if (tid == 0) {      // keep.append(order[0]); rows beyond max_det are cut afterwards (ops.py:408)
            const int k = s_nk;
            if (k < max_det) {
                const float4 bi = box[i];
                ob[k * 6 + 0] = bi.x; ob[k * 6 + 1] = bi.y; ob[k * 6 + 2] = bi.z; ob[k * 6 + 3] = bi.w;
                ob[k * 6 + 4] = score[i];
                ob[k * 6 + 5] = (float)cls[i];
            }
            s_nk = k + 1;
        }
        const float4 bi = box[i];
        const float ci = (float)cls[i] * off_scale;
        const float ix1 = __fadd_rn(bi.x, ci), iy1 = __fadd_rn(bi.y, ci), ix2 = __fadd_rn(bi.z, ci), iy2 = __fadd_rn(bi.w, ci);
        const float w1 = __fsub_rn(ix2, ix1), h1 = __fadd_rn(__fsub_rn(iy2, iy1), 1e-7f);
        // this thread's contiguous chunk of order[1:], so that the compaction below is stable
        const int rest = m - 1;
        const int chunk = (rest + kNmsThreads - 1) / kNmsThreads;
        const int t0 = 1 + tid * chunk, t1 = min(m, t0 + chunk);
        int alive = 0;
        for (int t = t0; t < t1; ++t) {
            const int idx = cur[t];
            float sc = score[idx];
            if (rest > 1) {      // with a single box left `iou` is 0-d and `(iou > thr).nonzero().squeeze()` is empty: no decay
                const float4 bj = box[idx];
                const float cj = (float)cls[idx] * off_scale;
                const float jx1 = __fadd_rn(bj.x, cj), jy1 = __fadd_rn(bj.y, cj), jx2 = __fadd_rn(bj.z, cj), jy2 = __fadd_rn(bj.w, cj);
                const float w2 = __fsub_rn(jx2, jx1), h2 = __fadd_rn(__fsub_rn(jy2, jy1), 1e-7f);
                const float iw = fmaxf(__fsub_rn(fminf(ix2, jx2), fmaxf(ix1, jx1)), 0.f);
                const float ih = fmaxf(__fsub_rn(fminf(iy2, jy2), fmaxf(iy1, jy1)), 0.f);
                const float inter = __fmul_rn(iw, ih);
                const float uni = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(w1, h1), __fmul_rn(w2, h2)), inter), 1e-7f);
                const float iou = __fdiv_rn(inter, uni);
                if (iou > iou_thres) {
                    sc = __fmul_rn(sc, expf(__fdiv_rn(-__fmul_rn(iou, iou), 0.5f)));
                    score[idx] = sc;
                }
            }
            alive += sc > 0.25f ? 1 : 0;
        }
        int total;
        int pos = block_excl_scan(alive, s_warp, total);
        // stable compaction + arg-max (first occurrence) of the survivors' scores
        float best = -INFINITY;
        int bpos = 0x7fffffff;
        for (int t = t0; t < t1; ++t) {
            const int idx = cur[t];
            const float sc = score[idx];
            if (sc > 0.25f) {
                nxt[pos] = idx;
                if (sc > best) { best = sc; bpos = pos; }
                ++pos;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob2 = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (ob2 > best || (ob2 == best && op < bpos)) { best = ob2; bpos = op; }
        }
        if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_bpos[tid >> 5] = bpos; }
        __syncthreads();
        if (tid == 0) {
            for (int wv = 1; wv < kNmsThreads / 32; ++wv)
                if (s_best[wv] > best || (s_best[wv] == best && s_bpos[wv] < bpos)) { best = s_best[wv]; bpos = s_bpos[wv]; }
            if (total == 0) s_stop = 1;
            else if (bpos != 0) {      // newOrder[[0, max]] = newOrder[[max, 0]]
                const int tmp = nxt[0];
                nxt[0] = nxt[bpos];
                nxt[bpos] = tmp;
            }
            s_m = total;
        }
        __syncthreads();
        int* sw = cur; cur = nxt; nxt = sw;
    }
    if (tid == 0) out_count[b] = min(s_nk, max_det);
}

}  // namespace ldc

using namespace ldc;

LDC_API size_t ldconv_nms_workspace_bytes(int B, int max_nms)
{
    return (size_t)B * (size_t)max_nms * (16 + 4 + 4 + 8);
}

// y (B, 4+nc, A) decoded predictions of the Detect head (bf16 or fp32) -> out (B, max_det, 6) fp32 rows (x1, y1, x2, y2, conf, cls)
// in keep order, out_count (B) int32 (negative: -candidates, more than max_nms candidates -- raise conf_thres).
LDC_API int ldconv_nms(const void* y, void* out, int32_t* out_count, void* workspace, size_t workspace_bytes, int B, int A, int nc,
                       float conf_thres, float iou_thres, int agnostic, int max_det, int max_nms, float max_wh, int dtype,
                       void* stream)
{
    LDC_REQUIRE(y && out && out_count && workspace, "ldconv_nms: null pointer");
    LDC_REQUIRE(B >= 0 && A >= 1 && nc >= 1 && max_det >= 1 && max_nms >= 1, "ldconv_nms: bad dims");
    LDC_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f && iou_thres >= 0.f && iou_thres <= 1.f, "ldconv_nms: thresholds must be in [0, 1]");
    LDC_REQUIRE(dtype == LDCONV_BF16 || dtype == LDCONV_F32, "ldconv_nms: unsupported dtype %d", dtype);
    if (B == 0) return LDCONV_OK;
    const int cap = max_nms < A ? max_nms : A;
    LDC_REQUIRE(workspace_bytes >= ldconv_nms_workspace_bytes(B, cap), "ldconv_nms: workspace too small");
    LDC_REQUIRE(aligned16(workspace), "ldconv_nms: workspace must be 16-byte aligned");
    uint8_t* ws = (uint8_t*)workspace;
    float4* ws_box = (float4*)ws;
    float* ws_score = (float*)(ws + (size_t)B * cap * 16);
    int* ws_cls = (int*)(ws + (size_t)B * cap * 20);
    int* ws_order = (int*)(ws + (size_t)B * cap * 24);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_BF16)
        nms_soft_kernel<__nv_bfloat16><<<B, kNmsThreads, 0, st>>>((const __nv_bfloat16*)y, A, nc, conf_thres, iou_thres, max_wh, agnostic,
                                                                 max_det, cap, ws_box, ws_score, ws_cls, ws_order, (float*)out, out_count);
    else
        nms_soft_kernel<float><<<B, kNmsThreads, 0, st>>>((const float*)y, A, nc, conf_thres, iou_thres, max_wh, agnostic, max_det, cap,
                                                         ws_box, ws_score, ws_cls, ws_order, (float*)out, out_count);
    LDC_LAUNCH_CHECK("nms_soft_kernel");
    return LDCONV_OK;
}
