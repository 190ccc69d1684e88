// ldconv_graph_ops.cu -- the small HBM-bound glue ops between the conv blocks of the DEAL-YOLO-LD graph (bf16, NHWC,
// sm_100a), each a single 128-bit-vectorised pass that can read / write channel slices of wider NHWC buffers (pixel
// strides ld*), so no torch.cat / torch.stack copy is needed around them:
//   * nearest up-sampling (nn.Upsample rows of yolov8-LD-P2.yaml:26,33) straight into the concat buffer;
//   * SSFF tail (reference nn/extra_modules/block.py:3432-3443 + Add :3479-3484): max over the three pyramid levels (the
//     two coarser ones read through nearest up-sampling index arithmetic instead of being materialised) + the residual add;
//   * the three chained 5x5 max-pools of SPPF (nn/modules/block.py:166-171) = 5x5, 9x9, 13x13 maxima of the same input,
//     computed together.
#include "common.cuh"

namespace ldc {

using T = __nv_bfloat16;

// thread = one 16-byte vector of a SOURCE pixel: one load, f x f stores (a thread per output vector spent its time on the
// 64-bit index decomposition and re-read every source vector f x f times: 2.1 TB/s at P2)
__global__ void __launch_bounds__(256)
upsample_nearest_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int H, int W, int CV, int f,
                        long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    const long long pix = t / CV;                    // source pixel (b, i, j)
    const int j = (int)(pix % W);
    const long long bi = pix / W;                    // b * H + i
    const uint4 v = *reinterpret_cast<const uint4*>(x + pix * ldx + cv * 8);
    const long long Wo = (long long)W * f;
    T* o = out + ((bi * f) * Wo + (long long)j * f) * ldo + cv * 8;
    for (int di = 0; di < f; ++di)
        for (int dj = 0; dj < f; ++dj) *reinterpret_cast<uint4*>(o + (di * Wo + dj) * ldo) = v;
}

// backward of the nearest up-sampling (training graph): grad_x[b, i, j, :] = sum of the f x f block of grad_out it was copied to,
// fp32 sum, one rounding.  Same thread mapping: one 16-byte vector of a source pixel, f x f loads, one store.  `ldg` lets grad_out
// be a channel slice of a wider NHWC tensor (the gradient of the Concat that follows the Upsample rows of the YAML).
__global__ void __launch_bounds__(256)
upsample_nearest_bwd_kernel(const T* __restrict__ gout, int ldg, T* __restrict__ gx, int ldx, int H, int W, int CV, int f, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    const long long pix = t / CV;
    const int j = (int)(pix % W);
    const long long bi = pix / W;
    const long long Wo = (long long)W * f;
    const T* g = gout + ((bi * f) * Wo + (long long)j * f) * ldg + cv * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, v[8];
    for (int di = 0; di < f; ++di)
        for (int dj = 0; dj < f; ++dj) {
            Vec16<T>::load(g + (di * Wo + dj) * ldg, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
    Vec16<T>::store(gx + pix * ldx + cv * 8, acc);
}

// SSFF tail of the TRAINING graph (nn/extra_modules/block.py:3438-3443: BatchNorm3d -> LeakyReLU(0.1) -> MaxPool3d((3,1,1)) over the
// depth axis of the stacked volume).  `pre` holds the three depth slices as (3 M, C) rows (slice d = rows d M ... d M + M - 1), scale /
// shift = the batch-statistics BatchNorm folded per channel.  Forward: out[m] = max_d leaky(pre[d M + m] * scale + shift).  Backward
// routing: dz[d M + m] = grad_out[m] * leaky'(z_d) for the FIRST slice that attains the maximum (torch's max_pool3d keeps the first
// index on ties), 0 for the other two -- the gradient w.r.t. the BatchNorm output, which the generic bn_act_bwd passes then take
// with act = none.
__device__ __forceinline__ float leaky01(float z) { return z > 0.f ? z : 0.1f * z; }

__global__ void __launch_bounds__(256)
ssff_max_fwd_kernel(const T* __restrict__ pre, const float* __restrict__ scale, const float* __restrict__ shift, T* __restrict__ out,
                    long long M, int CV, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    float sc[8], sh[8], a[8], b[8], c[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = scale[cv * 8 + e]; sh[e] = shift[cv * 8 + e]; }
    Vec16<T>::load(pre + t * 8, a);
    Vec16<T>::load(pre + (t + M * CV) * 8, b);
    Vec16<T>::load(pre + (t + 2 * M * CV) * 8, c);
#pragma unroll
    for (int e = 0; e < 8; ++e)
        a[e] = fmaxf(fmaxf(leaky01(fmaf(a[e], sc[e], sh[e])), leaky01(fmaf(b[e], sc[e], sh[e]))), leaky01(fmaf(c[e], sc[e], sh[e])));
    Vec16<T>::store(out + t * 8, a);
}

__global__ void __launch_bounds__(256)
ssff_max_bwd_kernel(const T* __restrict__ pre, const float* __restrict__ scale, const float* __restrict__ shift,
                    const T* __restrict__ gout, T* __restrict__ dz, long long M, int CV, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    float sc[8], sh[8], z[3][8], g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = scale[cv * 8 + e]; sh[e] = shift[cv * 8 + e]; }
#pragma unroll
    for (int d = 0; d < 3; ++d) Vec16<T>::load(pre + (t + d * M * CV) * 8, z[d]);
    Vec16<T>::load(gout + t * 8, g);
    float o[3][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float zz[3], y[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { zz[d] = fmaf(z[d][e], sc[e], sh[e]); y[d] = leaky01(zz[d]); }
        int arg = 0;
        if (y[1] > y[arg]) arg = 1;
        if (y[2] > y[arg]) arg = 2;
#pragma unroll
        for (int d = 0; d < 3; ++d) o[d][e] = d == arg ? g[e] * (zz[d] > 0.f ? 1.f : 0.1f) : 0.f;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) Vec16<T>::store(dz + (t + d * M * CV) * 8, o[d]);
}

// `Add` rows of the YAML (nn/extra_modules/block.py:3479-3484: torch.sum(torch.stack(x), 0)) for up to four NHWC inputs /
// channel slices: fp32 accumulation, one rounding -- what torch's reduction does for bf16 tensors
struct AddArgs { const T* src[4]; int ld[4]; int n; };
__global__ void __launch_bounds__(256)
add_nhwc_kernel(AddArgs a, T* __restrict__ out, int ldo, int CV, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    const long long pix = t / CV;
    float acc[8], v[8];
    Vec16<T>::load(a.src[0] + pix * a.ld[0] + cv * 8, acc);
    for (int k = 1; k < a.n; ++k) {
        Vec16<T>::load(a.src[k] + pix * a.ld[k] + cv * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    Vec16<T>::store(out + pix * ldo + cv * 8, acc);
}

__global__ void __launch_bounds__(256)
scalseq_tail_kernel(const T* __restrict__ z0, const T* __restrict__ z1, const T* __restrict__ z2, const T* __restrict__ add,
                    int ld_add, T* __restrict__ out, int ldo, int H, int W, int H1, int W1, int H2, int W2, int CV,
                    long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    const long long pix = t / CV;
    const int j = (int)(pix % W);
    const int i = (int)((pix / W) % H);
    const long long b = pix / ((long long)W * H);
    const int C = CV * 8;
    // torch nearest: src = floor(dst * in / out)
    const int i1 = (int)(((long long)i * H1) / H), j1 = (int)(((long long)j * W1) / W);
    const int i2 = (int)(((long long)i * H2) / H), j2 = (int)(((long long)j * W2) / W);
    float a[8], b1[8], c2[8];
    Vec16<T>::load(z0 + pix * C + cv * 8, a);
    Vec16<T>::load(z1 + ((b * H1 + i1) * W1 + j1) * C + cv * 8, b1);
    Vec16<T>::load(z2 + ((b * H2 + i2) * W2 + j2) * C + cv * 8, c2);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(fmaxf(a[e], b1[e]), c2[e]);
    if (add) {
        float r[8];
        Vec16<T>::load(add + pix * ld_add + cv * 8, r);
        // the reference adds two bf16 tensors: round the max first (it is already a bf16 value), then add and round once
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += r[e];
    }
    Vec16<T>::store(out + pix * ldo + cv * 8, a);
}

__global__ void __launch_bounds__(256)
sppf_pools_kernel(const T* __restrict__ x, T* __restrict__ o1, T* __restrict__ o2, T* __restrict__ o3, int ld, int H, int W,
                  int CV, int r, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    const long long pix = t / CV;
    const int j = (int)(pix % W);
    const int i = (int)((pix / W) % H);
    const long long b = pix / ((long long)W * H);
    float m1[8], m2[8], m3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m1[e] = m2[e] = m3[e] = -INFINITY;
    for (int di = -3 * r; di <= 3 * r; ++di) {
        const int ii = i + di;
        if (ii < 0 || ii >= H) continue;
        const int ai = di < 0 ? -di : di;
        for (int dj = -3 * r; dj <= 3 * r; ++dj) {
            const int jj = j + dj;
            if (jj < 0 || jj >= W) continue;
            const int aj = dj < 0 ? -dj : dj;
            const int ring = ai > aj ? ai : aj;
            float v[8];
            Vec16<T>::load(x + ((b * H + ii) * W + jj) * ld + cv * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                m3[e] = fmaxf(m3[e], v[e]);
                if (ring <= 2 * r) m2[e] = fmaxf(m2[e], v[e]);
                if (ring <= r) m1[e] = fmaxf(m1[e], v[e]);
            }
        }
    }
    Vec16<T>::store(o1 + pix * ld + cv * 8, m1);
    Vec16<T>::store(o2 + pix * ld + cv * 8, m2);
    Vec16<T>::store(o3 + pix * ld + cv * 8, m3);
}

// SPPF pools, one CTA per (image, 8-channel vector): the plane is staged in shared memory, a horizontal pass produces
// the 5/9/13-wide row maxima, a vertical pass the window maxima (40 loads per output instead of 169).
__global__ void __launch_bounds__(256)
sppf_pools_smem_kernel(const T* __restrict__ x, T* __restrict__ o1, T* __restrict__ o2, T* __restrict__ o3, int ld, int H, int W,
                       int CV, int r)
{
    extern __shared__ uint4 s_pl[];            // [4][H*W]: input, h(r), h(2r), h(3r)
    const int cv = blockIdx.x % CV;
    const long long b = blockIdx.x / CV;
    const int HW = H * W;
    const T* xb = x + b * HW * ld + cv * 8;
    for (int t = threadIdx.x; t < HW; t += blockDim.x) s_pl[t] = *reinterpret_cast<const uint4*>(xb + (long long)t * ld);
    __syncthreads();
    auto vmax = [](uint4 a, uint4 b2) {
        uint4 o;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b2);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) po[e] = __hmax2(pa[e], pb[e]);
        return o;
    };
    for (int t = threadIdx.x; t < HW; t += blockDim.x) {
        const int i = t / W, j = t % W;
        uint4 m = s_pl[t];
        uint4 m1 = m, m2 = m, m3 = m;
        for (int d = 1; d <= 3 * r; ++d) {
            if (j - d >= 0) m = vmax(m, s_pl[i * W + j - d]);
            if (j + d < W) m = vmax(m, s_pl[i * W + j + d]);
            if (d == r) m1 = m;
            if (d == 2 * r) m2 = m;
        }
        m3 = m;
        s_pl[HW + t] = m1; s_pl[2 * HW + t] = m2; s_pl[3 * HW + t] = m3;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < HW; t += blockDim.x) {
        const int i = t / W, j = t % W;
        uint4 a = s_pl[HW + t], b2 = s_pl[2 * HW + t], c = s_pl[3 * HW + t];
        for (int d = 1; d <= 3 * r; ++d) {
            const int up = i - d, dn = i + d;
            if (d <= r) {
                if (up >= 0) a = vmax(a, s_pl[HW + up * W + j]);
                if (dn < H) a = vmax(a, s_pl[HW + dn * W + j]);
            }
            if (d <= 2 * r) {
                if (up >= 0) b2 = vmax(b2, s_pl[2 * HW + up * W + j]);
                if (dn < H) b2 = vmax(b2, s_pl[2 * HW + dn * W + j]);
            }
            if (up >= 0) c = vmax(c, s_pl[3 * HW + up * W + j]);
            if (dn < H) c = vmax(c, s_pl[3 * HW + dn * W + j]);
        }
        const long long o = (b * HW + t) * ld + cv * 8;
        *reinterpret_cast<uint4*>(o1 + o) = a;
        *reinterpret_cast<uint4*>(o2 + o) = b2;
        *reinterpret_cast<uint4*>(o3 + o) = c;
    }
}

// SPPF pools as the reference computes them (nn/modules/block.py:166-171: y1 = m(x), y2 = m(y1), y3 = m(y2), m = MaxPool2d(k, 1, k/2)):
// three cascaded separable passes over TWO shared-memory planes (cur, horizontal maxima) instead of four, so that four CTAs fit an
// SM and the 512 (image, channel-vector) planes of the model's SPPF are resident in one wave (the four-plane kernel ran 1.7 waves
// of two CTAs per SM).  Padding is "ignore" (-inf) like MaxPool2d, so the cascade equals the 5 / 9 / 13 windows exactly.
__global__ void __launch_bounds__(256)
sppf_pools_cascade_kernel(const T* __restrict__ x, T* __restrict__ o1, T* __restrict__ o2, T* __restrict__ o3, int ld, int H, int W,
                          int CV, int r)
{
    extern __shared__ uint4 s_pl[];            // [2][H*W]: current map, its horizontal maxima
    const int cv = blockIdx.x % CV;
    const long long b = blockIdx.x / CV;
    const int HW = H * W;
    const T* xb = x + b * HW * ld + cv * 8;
    for (int t = threadIdx.x; t < HW; t += blockDim.x) s_pl[t] = *reinterpret_cast<const uint4*>(xb + (long long)t * ld);
    __syncthreads();
    auto vmax = [](uint4 a, uint4 b2) {
        uint4 o;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b2);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) po[e] = __hmax2(pa[e], pb[e]);
        return o;
    };
    T* outs[3] = {o1, o2, o3};
#pragma unroll 1
    for (int stage = 0; stage < 3; ++stage) {
        for (int t = threadIdx.x; t < HW; t += blockDim.x) {
            const int j = t % W;
            uint4 m = s_pl[t];
            for (int d = 1; d <= r; ++d) {
                if (j - d >= 0) m = vmax(m, s_pl[t - d]);
                if (j + d < W) m = vmax(m, s_pl[t + d]);
            }
            s_pl[HW + t] = m;
        }
        __syncthreads();
        T* op = outs[stage];
        for (int t = threadIdx.x; t < HW; t += blockDim.x) {
            const int i = t / W;
            uint4 m = s_pl[HW + t];
            for (int d = 1; d <= r; ++d) {
                if (i - d >= 0) m = vmax(m, s_pl[HW + t - d * W]);
                if (i + d < H) m = vmax(m, s_pl[HW + t + d * W]);
            }
            s_pl[t] = m;
            *reinterpret_cast<uint4*>(op + (b * HW + t) * ld + cv * 8) = m;
        }
        __syncthreads();
    }
}

// uint8 NCHW image batch (what the reference's predictor uploads, engine/predictor.py:120-131) -> bf16 NHWC in [0,1]:
// the `im.half(); im /= 255` of the reference plus the layout change, one pass.
__global__ void __launch_bounds__(256)
image_u8_to_nhwc_kernel(const uint8_t* __restrict__ x, T* __restrict__ out, int C, long long hw, float scale, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long b = t / hw, p = t % hw;
    const uint8_t* src = x + b * C * hw + p;
    T* dst = out + t * C;
    for (int c = 0; c < C; ++c) dst[c] = __float2bfloat16_rn((float)src[(long long)c * hw] * scale);
}

// C = 3, 8 pixels per thread: one 8-byte load per colour plane, 48 contiguous output bytes as three 16-byte stores (the per-pixel
// kernel above issues three 1-byte loads and three 2-byte stores per pixel)
__global__ void __launch_bounds__(256)
image_u8_to_nhwc3_x8_kernel(const uint8_t* __restrict__ x, T* __restrict__ out, long long hw, float scale, long long groups)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= groups) return;
    const long long gpi = hw >> 3;                       // groups of 8 pixels per image
    const long long b = t / gpi, p = (t - b * gpi) << 3;
    const uint8_t* src = x + b * 3 * hw + p;
    const uint2 r = *reinterpret_cast<const uint2*>(src), g = *reinterpret_cast<const uint2*>(src + hw),
                bl = *reinterpret_cast<const uint2*>(src + 2 * hw);
    const uint32_t rw[2] = {r.x, r.y}, gw[2] = {g.x, g.y}, bw[2] = {bl.x, bl.y};
    uint32_t w[12];                                       // 8 pixels x 3 channels bf16 = 24 values = 12 words
    float v[24];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[3 * k] = (float)((rw[k >> 2] >> (8 * (k & 3))) & 0xffu) * scale;
        v[3 * k + 1] = (float)((gw[k >> 2] >> (8 * (k & 3))) & 0xffu) * scale;
        v[3 * k + 2] = (float)((bw[k >> 2] >> (8 * (k & 3))) & 0xffu) * scale;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
    uint4* dst = reinterpret_cast<uint4*>(out + (b * hw + p) * 3);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
}

}  // namespace ldc

using namespace ldc;

LDC_API int ldconv_image_u8_to_nhwc(const void* x_u8, void* out, int B, int C, int H, int W, float scale, int dtype,
                                    void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_image_u8_to_nhwc: bf16 only");
    LDC_REQUIRE(x_u8 && out && C >= 1 && C <= 16, "ldconv_image_u8_to_nhwc: bad arguments");
    const long long total = (long long)B * H * W;
    if (total == 0) return LDCONV_OK;
    if (C == 3 && ((long long)H * W) % 8 == 0 && (reinterpret_cast<uintptr_t>(x_u8) & 7) == 0 && aligned16(out)) {
        const long long groups = total / 8;
        image_u8_to_nhwc3_x8_kernel<<<cdiv(groups, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)x_u8, (T*)out, (long long)H * W,
                                                                                       scale, groups);
        LDC_LAUNCH_CHECK("image_u8_to_nhwc3_x8_kernel");
        return LDCONV_OK;
    }
    image_u8_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)x_u8, (T*)out, C,
                                                                               (long long)H * W, scale, total);
    LDC_LAUNCH_CHECK("image_u8_to_nhwc_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_upsample_nearest(const void* x, int ldx, void* out, int ldo, int B, int H, int W, int C, int factor,
                                    int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_upsample_nearest: bf16 only");
    LDC_REQUIRE(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && factor >= 1 && aligned16(x) && aligned16(out),
                "ldconv_upsample_nearest: needs C, ldx, ldo multiples of 8 and 16-byte aligned pointers");
    const long long total = (long long)B * H * W * (C / 8);
    if (total == 0) return LDCONV_OK;
    upsample_nearest_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, ldx, (T*)out, ldo, H, W, C / 8,
                                                                               factor, total);
    LDC_LAUNCH_CHECK("upsample_nearest_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_ssff_max_fwd(const void* pre, const float* scale, const float* shift, void* out, long long M, int C, int dtype,
                                void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_ssff_max_fwd: bf16 only");
    LDC_REQUIRE(pre && scale && shift && out && M >= 0 && C >= 8 && C % 8 == 0 && aligned16(pre) && aligned16(out),
                "ldconv_ssff_max_fwd: needs C %% 8 == 0 and 16-byte aligned pointers");
    const long long total = M * (C / 8);
    if (total == 0) return LDCONV_OK;
    ssff_max_fwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)pre, scale, shift, (T*)out, M, C / 8, total);
    LDC_LAUNCH_CHECK("ssff_max_fwd_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_ssff_max_bwd(const void* pre, const float* scale, const float* shift, const void* grad_out, void* dz, long long M,
                                int C, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_ssff_max_bwd: bf16 only");
    LDC_REQUIRE(pre && scale && shift && grad_out && dz && M >= 0 && C >= 8 && C % 8 == 0 && aligned16(pre) && aligned16(grad_out) &&
                    aligned16(dz), "ldconv_ssff_max_bwd: needs C %% 8 == 0 and 16-byte aligned pointers");
    const long long total = M * (C / 8);
    if (total == 0) return LDCONV_OK;
    ssff_max_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)pre, scale, shift, (const T*)grad_out, (T*)dz, M,
                                                                           C / 8, total);
    LDC_LAUNCH_CHECK("ssff_max_bwd_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_upsample_nearest_bwd(const void* grad_out, int ldg, void* grad_x, int ldx, int B, int H, int W, int C, int factor,
                                        int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_upsample_nearest_bwd: bf16 only");
    LDC_REQUIRE(grad_out && grad_x && C % 8 == 0 && ldg % 8 == 0 && ldx % 8 == 0 && factor >= 1 && aligned16(grad_out) && aligned16(grad_x),
                "ldconv_upsample_nearest_bwd: needs C, ldg, ldx multiples of 8 and 16-byte aligned pointers");
    const long long total = (long long)B * H * W * (C / 8);     // (H, W): the SOURCE (low-resolution) size
    if (total == 0) return LDCONV_OK;
    upsample_nearest_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)grad_out, ldg, (T*)grad_x, ldx, H, W, C / 8,
                                                                                   factor, total);
    LDC_LAUNCH_CHECK("upsample_nearest_bwd_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_add_nhwc(const void* const* srcs, const int* lds, int n, void* out, int ldo, long long pixels, int C, int dtype,
                            void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_add_nhwc: bf16 only");
    LDC_REQUIRE(srcs && lds && out && n >= 1 && n <= 4 && C % 8 == 0 && ldo % 8 == 0 && aligned16(out), "ldconv_add_nhwc: bad arguments");
    AddArgs a;
    a.n = n;
    for (int k = 0; k < 4; ++k) {
        a.src[k] = (const T*)srcs[k < n ? k : 0];
        a.ld[k] = lds[k < n ? k : 0];
        LDC_REQUIRE(a.src[k] && aligned16(a.src[k]) && a.ld[k] % 8 == 0, "ldconv_add_nhwc: input %d must be 16-byte aligned, stride % 8 == 0", k);
    }
    const long long total = pixels * (C / 8);
    if (total == 0) return LDCONV_OK;
    add_nhwc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(a, (T*)out, ldo, C / 8, total);
    LDC_LAUNCH_CHECK("add_nhwc_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_scalseq_tail(const void* z0, const void* z1, const void* z2, const void* addend, int ld_add, void* out,
                                int ldo, int B, int H, int W, int H1, int W1, int H2, int W2, int C, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_scalseq_tail: bf16 only");
    LDC_REQUIRE(z0 && z1 && z2 && out && C % 8 == 0 && ldo % 8 == 0 && (!addend || ld_add % 8 == 0),
                "ldconv_scalseq_tail: bad arguments");
    LDC_REQUIRE(aligned16(z0) && aligned16(z1) && aligned16(z2) && aligned16(out) && (!addend || aligned16(addend)),
                "ldconv_scalseq_tail: pointers must be 16-byte aligned");
    const long long total = (long long)B * H * W * (C / 8);
    if (total == 0) return LDCONV_OK;
    scalseq_tail_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)z0, (const T*)z1, (const T*)z2,
                                                                           (const T*)addend, ld_add, (T*)out, ldo, H, W, H1,
                                                                           W1, H2, W2, C / 8, total);
    LDC_LAUNCH_CHECK("scalseq_tail_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_sppf_pools(const void* x, void* o1, void* o2, void* o3, int ld, int B, int H, int W, int C, int k,
                              int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_sppf_pools: bf16 only");
    LDC_REQUIRE(x && o1 && o2 && o3 && C % 8 == 0 && ld % 8 == 0 && k >= 1 && (k & 1), "ldconv_sppf_pools: bad arguments");
    LDC_REQUIRE(aligned16(x) && aligned16(o1) && aligned16(o2) && aligned16(o3), "ldconv_sppf_pools: alignment");
    const long long total = (long long)B * H * W * (C / 8);
    if (total == 0) return LDCONV_OK;
    constexpr int cascade = 1;      // three cascaded poolings over two planes: 58 -> 44 us
    const size_t plane2 = (size_t)2 * H * W * 16;
    if (cascade && plane2 <= 200 * 1024 && (long long)B * (C / 8) <= 0x7fffffffll) {
        LDC_CUDA(cudaFuncSetAttribute(sppf_pools_cascade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane2));
        sppf_pools_cascade_kernel<<<(unsigned)(B * (C / 8)), 256, plane2, (cudaStream_t)stream>>>((const T*)x, (T*)o1, (T*)o2, (T*)o3,
                                                                                                 ld, H, W, C / 8, k / 2);
        LDC_LAUNCH_CHECK("sppf_pools_cascade_kernel");
        return LDCONV_OK;
    }
    const size_t plane = (size_t)4 * H * W * 16;
    if (plane <= 200 * 1024 && (long long)B * (C / 8) <= 0x7fffffffll) {
        LDC_CUDA(cudaFuncSetAttribute(sppf_pools_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane));
        sppf_pools_smem_kernel<<<(unsigned)(B * (C / 8)), 256, plane, (cudaStream_t)stream>>>((const T*)x, (T*)o1, (T*)o2, (T*)o3,
                                                                                              ld, H, W, C / 8, k / 2);
        LDC_LAUNCH_CHECK("sppf_pools_smem_kernel");
        return LDCONV_OK;
    }
    sppf_pools_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)o1, (T*)o2, (T*)o3, ld, H, W, C / 8,
                                                                         k / 2, total);
    LDC_LAUNCH_CHECK("sppf_pools_kernel");
    return LDCONV_OK;
}
