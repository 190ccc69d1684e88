// ldconv_offconv_bwd.cu -- backward of LDConv's offset conv p_conv (/root/reference/ultralytics/nn/modules/conv.py:356,368:
// Conv2d(inc, 2N, 3, padding=1, stride=s)) for bf16 activations, sm_100a.
//
//   weight gradient   dW[o, (tap, c)] = sum_m grad_off[m, o] * x[b, i*s + ky - 1, j*s + kx - 1, c]
//     is a reduction over M = B*h*w output pixels with a tiny 2N x 9C result: a tensor-core job.  The 3x3 neighbourhoods are
//     laid out chunk by chunk as an (rows, 9C) bf16 matrix in a caller-provided workspace sized to stay in L2 (im2col3x3_kernel),
//     grad_off is rounded once to bf16 (column sums = the bias gradient come out of the same pass in fp32), and every chunk is
//     reduced by the MN-major tcgen05 kernel of ldconv_wgrad_umma.cu (both operands straight from their row-major layout,
//     fp32 accumulation in TMEM).  The CUDA-core kernel this replaces took 1.4-6.0 ms per layer call at batch 64
//     (profiles/r1_bwd_launches_L*.csv), 37-57 % of a whole LDConv forward + backward.
//   data gradient     grad_x[b, r, k, c] += sum_{taps, o} grad_off[b, i, j, o] * w[ky, kx, c, o]   (i*s + ky - 1 = r, ...)
//     gather form, one thread per (input pixel, 4 channels), weights transposed in shared memory so that one LDS.128 feeds
//     four FMAs; grad_x is read-modify-written with 16-byte accesses (it already holds the scatter of the bilinear gather).
#include "common.cuh"

namespace ldc {

int wgrad_umma_supported(int M, int K, int O, const void* g, const void* a);
int wgrad_umma(const void* g, const void* a, float* dW, int M, int K, int O, cudaStream_t st);

// ---- grad_off (M, O2) fp32 -> (M, O2P) bf16 (zero-padded columns) + column sums (bias gradient) ------------------------------
__global__ void __launch_bounds__(256)
goff_to_bf16_kernel(const float* __restrict__ goff, __nv_bfloat16* __restrict__ g16, float* __restrict__ grad_b, long long M,
                    int O2, int O2P, int rows_per_cta)
{
    __shared__ float s_sum[32];
    if (threadIdx.x < 32) s_sum[threadIdx.x] = 0.f;
    __syncthreads();
    const long long m0 = (long long)blockIdx.x * rows_per_cta;
    const long long m1 = min(M, m0 + rows_per_cta);
    // thread = (row within a group, column); groups of 256 / O2P rows
    const int o = threadIdx.x % O2P, rsub = threadIdx.x / O2P, rstep = blockDim.x / O2P;
    float acc = 0.f;
    if (rsub < rstep) {
        for (long long m = m0 + rsub; m < m1; m += rstep) {
            const float v = o < O2 ? goff[m * O2 + o] : 0.f;
            acc += v;
            g16[m * O2P + o] = __float2bfloat16_rn(v);
        }
        if (grad_b && o < O2) atomicAdd(&s_sum[o], acc);
    }
    __syncthreads();
    if (grad_b && threadIdx.x < O2) atomicAdd(grad_b + threadIdx.x, s_sum[threadIdx.x]);
}

// ---- im2col of the 3x3 / pad 1 / stride s neighbourhoods of output pixels [m0, m0 + rows) -> col (rows, Kp) bf16 -------------
// VEC:  thread = (row, tap), copies the C channels of that tap as 16-byte vectors (the pixel decode is paid once per tap)
// !VEC: thread = row (small C, e.g. the 3-channel image): 9*C scalar loads, Kp/8 16-byte stores
template <bool VEC>
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, int C, int H, int W, int h, int w, int s,
                 long long m0, int rows, int Kp)
{
    if (VEC) {
        const int CV = C / 8;
        const long long total = (long long)rows * 9;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(t / 9);
            const int tap = (int)(t - (long long)r * 9);
            const long long m = m0 + r;
            const int j = (int)(m % w);
            const int i = (int)((m / w) % h);
            const long long b = m / ((long long)w * h);
            const int rr = i * s + tap / 3 - 1, kk = j * s + tap % 3 - 1;
            uint4* dst = reinterpret_cast<uint4*>(col + (size_t)r * Kp + (size_t)tap * C);
            if (rr >= 0 && rr < H && kk >= 0 && kk < W) {
                const uint4* src = reinterpret_cast<const uint4*>(x + ((b * H + rr) * W + kk) * C);
                for (int cv = 0; cv < CV; ++cv) dst[cv] = __ldg(src + cv);
            } else {
                for (int cv = 0; cv < CV; ++cv) dst[cv] = make_uint4(0, 0, 0, 0);
            }
        }
    } else {
        for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
            const long long m = m0 + r;
            const int j = (int)(m % w);
            const int i = (int)((m / w) % h);
            const long long b = m / ((long long)w * h);
            uint4* dst = reinterpret_cast<uint4*>(col + (size_t)r * Kp);
            uint32_t pack[4];
            int filled = 0, k = 0;
            for (int k8 = 0; k8 < Kp; k8 += 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e, ++k) {
                    unsigned short v = 0;
                    if (k < 9 * C) {
                        const int tap = k / C, c = k - tap * C;
                        const int rr = i * s + tap / 3 - 1, kk = j * s + tap % 3 - 1;
                        if (rr >= 0 && rr < H && kk >= 0 && kk < W)
                            v = __bfloat16_as_ushort(x[((b * H + rr) * W + kk) * C + c]);
                    }
                    if (e & 1) pack[e >> 1] |= (uint32_t)v << 16; else pack[e >> 1] = v;
                }
                dst[k8 >> 3] = make_uint4(pack[0], pack[1], pack[2], pack[3]);
                ++filled;
            }
        }
    }
}

// ---- dW (O2P, Kp) fp32 [o][tap*C + c]  ->  grad_w (3,3,C,2N) fp32 [tap][c][o] (accumulated) -----------------------------------
__global__ void dw_scatter_kernel(const float* __restrict__ dW, float* __restrict__ grad_w, int C, int O2, int Kp)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 9 * C * O2) return;
    const int o = t % O2, kc = t / O2;
    grad_w[t] += dW[(size_t)o * Kp + kc];
}

// ---- data gradient ------------------------------------------------------------------------------------------------------
// grad_x[b, r, k, c] += sum over taps (ky, kx) with (r + 1 - ky, k + 1 - kx) = s * (i, j):  sum_o grad_off[b, i, j, o] * w[ky, kx, c, o]
// (conv_transpose of conv.py:356 in gather form).  Thread = (column k, CG channels), a CTA walks `rows` rows of one image: grid
// (column segments, row chunks, B), no index division per element.  The sum is ADDED with one 16-byte reduction per thread and row
// (red.global.add: v4.f32, or v4.bf16x2 for the 16-bit accumulator of ldconv_gather_bwd_acc16) instead of a read-modify-write:
// the first versions loaded the old value, one 8 / 16-byte request in flight per thread, and ran at 0.75-0.85 TB/s however the
// index arithmetic was written (532 / 467 us for 399 MB at layer 1, batch 64: latency-bound); a reduction returns nothing, so
// nothing waits for DRAM.  fp32: the same rounding as load-add-store.  bf16: the sum is rounded to bf16 before it is added.
// S = 1 / 2: compile-time stride (the model's values) so the tap validity test is a bit test; S = 0: any stride.
template <int NMAX, int S, typename ACC>
__global__ void __launch_bounds__(256)
offconv_bwd_data_kernel(const float* __restrict__ goff, const float* __restrict__ w, ACC* __restrict__ grad_x, int C, int H, int W,
                        int h, int wo, int N, int s_rt, int rows, int segs, int chunks, int items)
{
    constexpr bool V16 = sizeof(ACC) == 2;
    constexpr int CG = V16 ? 8 : 4;                         // channels per thread: one 16-byte reduction
    extern __shared__ __align__(16) float s_w[];            // [9][2N][Cp]: w (3,3,C,2N) transposed to channel-fastest, Cp = C rounded up to CG
    const int s = S ? S : s_rt;
    const int O2 = 2 * N;
    const int CGn = (C + CG - 1) / CG, Cp = CGn * CG;
    for (int t = threadIdx.x; t < 9 * O2 * Cp; t += blockDim.x) {
        const int c = t % Cp, o = (t / Cp) % O2, tap = t / (Cp * O2);
        s_w[t] = c < C ? w[((size_t)tap * C + c) * O2 + o] : 0.f;
    }
    __syncthreads();
    // work item = (image, row chunk, column segment); a CTA stages the weights once and walks its items (grid = a few CTAs per SM:
    // with num_param 9 the weights are 20 KB, re-staging them for every chunk of a small map cost more than the chunk itself)
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int seg = item % segs, chunk = (item / segs) % chunks, b = item / (segs * chunks);
    const unsigned t = (unsigned)seg * blockDim.x + threadIdx.x;
    if (t >= (unsigned)(W * CGn)) continue;
    const int k = (int)(t / (unsigned)CGn), cg = (int)(t - (unsigned)k * (unsigned)CGn);
    // the (at most three) taps of this column: kx valid iff k + 1 - kx = s * j with 0 <= j < wo
    int jx[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
        const int kj = k + 1 - kx;
        const bool ok = kj >= 0 && (S == 2 ? (kj & 1) : (S == 1 ? 0 : kj % s)) == 0;
        const int j = S == 2 ? (kj >> 1) : (S == 1 ? kj : kj / s);
        jx[kx] = (ok && j < wo) ? j : -1;
    }
    const int r_begin = chunk * rows, r_end = min(H, r_begin + rows);
    for (int r = r_begin; r < r_end; ++r) {
        float acc[CG];
#pragma unroll
        for (int e = 0; e < CG; ++e) acc[e] = 0.f;
        bool any = false;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ri = r + 1 - ky;
            if (ri < 0 || (S == 2 ? (ri & 1) : (S == 1 ? 0 : ri % s)) != 0) continue;
            const int i = S == 2 ? (ri >> 1) : (S == 1 ? ri : ri / s);
            if (i >= h) continue;
            const float* grow = goff + ((size_t)b * h + i) * wo * O2;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                if (jx[kx] < 0) continue;
                any = true;
                const float2* gp = reinterpret_cast<const float2*>(grow + (size_t)jx[kx] * O2);
                const float4* wp = reinterpret_cast<const float4*>(s_w + ((size_t)(ky * 3 + kx) * O2) * Cp + cg * CG);
#pragma unroll
                for (int n = 0; n < NMAX; ++n) {
                    if (n >= N) break;
                    const float2 g = __ldg(gp + n);
#pragma unroll
                    for (int q = 0; q < CG / 4; ++q) {
                        const float4 w0 = wp[(size_t)(2 * n) * (Cp / 4) + q], w1 = wp[(size_t)(2 * n + 1) * (Cp / 4) + q];
                        acc[4 * q + 0] = fmaf(g.x, w0.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(g.x, w0.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(g.x, w0.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(g.x, w0.w, acc[4 * q + 3]);
                        acc[4 * q + 0] = fmaf(g.y, w1.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(g.y, w1.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(g.y, w1.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(g.y, w1.w, acc[4 * q + 3]);
                    }
                }
            }
        }
        if (!any) continue;
        ACC* dst = grad_x + (((size_t)b * H + r) * W + k) * C + cg * CG;
        if constexpr (V16) {
            uint32_t p0, p1, p2, p3;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(acc[1]), "f"(acc[0]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(acc[3]), "f"(acc[2]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(acc[5]), "f"(acc[4]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(acc[7]), "f"(acc[6]));
            asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
        } else if ((C & 3) == 0) {
            atomicAdd(reinterpret_cast<float4*>(dst), make_float4(acc[0], acc[1], acc[2], acc[3]));
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (cg * 4 + e < C) atomicAdd(dst + e, acc[e]);
        }
    }
    }
}

// returns 1 when launched, 0 when the shape is outside the kernel's range (caller uses the generic kernel)
template <typename ACC>
static int offconv_bwd_data_fast_t(const float* goff, const float* w, ACC* grad_x, int B, int C, int H, int W, int N, int s,
                                   cudaStream_t st)
{
    constexpr int CG = sizeof(ACC) == 2 ? 8 : 4;
    const int CGn = (C + CG - 1) / CG;
    const size_t smem = (size_t)9 * 2 * N * CGn * CG * sizeof(float);
    if (smem > 96 * 1024 || N > 16 || (long long)W * CGn > 0x7fffffffll) return 0;
    if (sizeof(ACC) == 2 && C % 8 != 0) return 0;
    const int h = out_size(H, s), wo = out_size(W, s);
    const long long segs = ((long long)W * CGn + 255) / 256;
    // rows per item: enough items to fill the machine a few times over, at most 16 rows
    int rows = 16;
    while (rows > 1 && segs * ((H + rows - 1) / rows) * B < (long long)num_sms() * 16) rows >>= 1;
    const long long chunks = (H + rows - 1) / rows;
    const long long items = segs * chunks * B;
    if (items > 0x7fffffffll) return 0;
    const unsigned grid = (unsigned)min(items, (long long)num_sms() * 8);
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, 256, smem, st>>>(goff, w, grad_x, C, H, W, h, wo, N, s, rows, (int)segs, (int)chunks, (int)items);
    };
#define LDC_BWD_DATA(NM)                                                          \
    do {                                                                          \
        if (s == 1) launch(offconv_bwd_data_kernel<NM, 1, ACC>);                  \
        else if (s == 2) launch(offconv_bwd_data_kernel<NM, 2, ACC>);             \
        else launch(offconv_bwd_data_kernel<NM, 0, ACC>);                         \
    } while (0)
    if (N <= 1) LDC_BWD_DATA(1);
    else if (N <= 3) LDC_BWD_DATA(3);
    else if (N <= 5) LDC_BWD_DATA(5);
    else if (N <= 9) LDC_BWD_DATA(9);
    else LDC_BWD_DATA(16);
#undef LDC_BWD_DATA
    return 1;
}

int offconv_bwd_data_fast(const float* goff, const float* w, float* grad_x, int B, int C, int H, int W, int N, int s, cudaStream_t st)
{
    return offconv_bwd_data_fast_t<float>(goff, w, grad_x, B, C, H, W, N, s, st);
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static size_t chunk_rows_for(long long M, int Kp)
{
    // the im2col chunk should stay resident in the 126 MB L2 between its producer and the reduction that reads it
    long long rows = (48ll << 20) / ((long long)Kp * 2);
    rows = rows / 64 * 64;
    if (rows < 4096) rows = 4096;
    if (rows > M) rows = (M + 63) / 64 * 64;
    return (size_t)rows;
}

size_t offconv_bwd_tc_workspace(int B, int C, int H, int W, int N, int s)
{
    const long long M = (long long)B * out_size(H, s) * out_size(W, s);
    const int O2P = round_up(2 * N, 8), Kp = round_up(9 * C, 8);
    const size_t g16 = ((size_t)M * O2P * 2 + 255) & ~(size_t)255;
    const size_t col = (chunk_rows_for(M, Kp) * (size_t)Kp * 2 + 255) & ~(size_t)255;
    const size_t dw = ((size_t)O2P * Kp * 4 + 255) & ~(size_t)255;
    return g16 + col + dw + 256;
}

// grad_x (fp32) or grad_x16 (bf16 accumulator), at most one of them
int offconv_bwd_tc(const float* goff, const __nv_bfloat16* x, const float* w, float* grad_x, __nv_bfloat16* grad_x16, float* grad_w,
                   float* grad_b, void* workspace, size_t workspace_bytes, int B, int C, int H, int W, int N, int s, cudaStream_t st)
{
    const int h = out_size(H, s), wo = out_size(W, s);
    const long long M = (long long)B * h * wo;
    const int O2 = 2 * N, O2P = round_up(O2, 8), Kp = round_up(9 * C, 8);
    if (M > 0x7fffffffll) return fail(LDCONV_E_ARG, "offset conv backward: too many output pixels");
    if (workspace_bytes < offconv_bwd_tc_workspace(B, C, H, W, N, s))
        return fail(LDCONV_E_ARG, "offset conv backward: workspace of %zu bytes is too small", workspace_bytes);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    __nv_bfloat16* g16 = reinterpret_cast<__nv_bfloat16*>(ws);
    const size_t g16_bytes = ((size_t)M * O2P * 2 + 255) & ~(size_t)255;
    const size_t rows_chunk = chunk_rows_for(M, Kp);
    __nv_bfloat16* col = reinterpret_cast<__nv_bfloat16*>(ws + g16_bytes);
    const size_t col_bytes = (rows_chunk * (size_t)Kp * 2 + 255) & ~(size_t)255;
    float* dW = reinterpret_cast<float*>(ws + g16_bytes + col_bytes);

    if (grad_x16) {
        if (!offconv_bwd_data_fast_t<__nv_bfloat16>(goff, w, grad_x16, B, C, H, W, N, s, st))
            return fail(LDCONV_E_ARG, "ldconv_offset_conv_bwd_tc_acc16: shape outside the 16-bit accumulator kernel (C=%d, N=%d)", C, N);
        LDC_LAUNCH_CHECK("offconv_bwd_data_kernel");
    }
    if (grad_x) {       // data gradient first: it only needs grad_off
        if (offconv_bwd_data_fast_t<float>(goff, w, grad_x, B, C, H, W, N, s, st)) {
            LDC_LAUNCH_CHECK("offconv_bwd_data_kernel");
        } else {        // weights do not fit shared memory (e.g. C=256, N=9): the generic kernel of ldconv_core.cu
            if (int e = ldconv_offset_conv_bwd(goff, x, w, grad_x, nullptr, nullptr, B, C, H, W, N, s, LDCONV_BF16, (void*)st)) return e;
        }
    }
    if (!grad_w && !grad_b) return LDCONV_OK;
    {
        int ctas = num_sms() * 4;
        long long rows = (M + ctas - 1) / ctas;
        if (rows < 64) rows = 64;
        ctas = (int)((M + rows - 1) / rows);
        goff_to_bf16_kernel<<<ctas, 256, 0, st>>>(goff, g16, grad_b, M, O2, O2P, (int)rows);
        LDC_LAUNCH_CHECK("goff_to_bf16_kernel");
    }
    if (!grad_w) return LDCONV_OK;
    if (!wgrad_umma_supported((int)M, Kp, O2P, g16, col))
        return fail(LDCONV_E_ARG, "offset conv backward: tensor-core reduction does not cover K=%d O=%d", Kp, O2P);
    LDC_CUDA(cudaMemsetAsync(dW, 0, (size_t)O2P * Kp * 4, st));
    for (long long m0 = 0; m0 < M; m0 += (long long)rows_chunk) {
        const int rows = (int)min((long long)rows_chunk, M - m0);
        const long long items = (long long)rows * ((C % 8 == 0) ? 9 : 1);
        const unsigned blocks = (unsigned)min((long long)num_sms() * 16, (items + 255) / 256);
        if (C % 8 == 0)
            im2col3x3_kernel<true><<<blocks, 256, 0, st>>>(x, col, C, H, W, h, wo, s, m0, rows, Kp);
        else
            im2col3x3_kernel<false><<<blocks, 256, 0, st>>>(x, col, C, H, W, h, wo, s, m0, rows, Kp);
        LDC_LAUNCH_CHECK("im2col3x3_kernel");
        if (int e = wgrad_umma(g16 + (size_t)m0 * O2P, col, dW, rows, Kp, O2P, st)) return e;
    }
    dw_scatter_kernel<<<cdiv(9ll * C * O2, 256), 256, 0, st>>>(dW, grad_w, C, O2, Kp);
    LDC_LAUNCH_CHECK("dw_scatter_kernel");
    set_impl(LDCONV_IMPL_TCGEN05);
    return LDCONV_OK;
}

}  // namespace ldc

using namespace ldc;

LDC_API size_t ldconv_offset_conv_bwd_workspace_bytes(int B, int C, int H, int W, int N, int s, int dtype)
{
    if (dtype != LDCONV_BF16 || B < 1 || C < 1 || H < 1 || W < 1 || N < 1 || N > 16 || s < 1) return 0;
    return offconv_bwd_tc_workspace(B, C, H, W, N, s);
}

LDC_API int ldconv_offset_conv_bwd_tc(const float* grad_off, const void* x, const float* w, float* grad_x, float* grad_w,
                                      float* grad_b, void* workspace, size_t workspace_bytes, int B, int C, int H, int W, int N,
                                      int s, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_offset_conv_bwd_tc: bf16 activations only (fp32 uses ldconv_offset_conv_bwd)");
    LDC_REQUIRE(grad_off && x && w && workspace, "ldconv_offset_conv_bwd_tc: null pointer");
    LDC_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1 && N >= 1 && N <= 16 && s >= 1, "ldconv_offset_conv_bwd_tc: bad dims");
    if (B == 0) return LDCONV_OK;
    return offconv_bwd_tc(grad_off, (const __nv_bfloat16*)x, w, grad_x, nullptr, grad_w, grad_b, workspace, workspace_bytes, B, C, H,
                          W, N, s, (cudaStream_t)stream);
}

LDC_API int ldconv_bwd_acc16_supported(int B, int C, int H, int W, int N, int s)
{
    // gather_bwd: 16-byte bf16 vectors (C % 8 == 0); data gradient: the offset conv's weights must fit its shared-memory copy
    const size_t smem = (size_t)9 * 2 * N * ((C + 3) / 4) * 4 * sizeof(float);
    return B >= 1 && C >= 8 && C % 8 == 0 && H >= 1 && W >= 1 && N >= 1 && N <= 16 && s >= 1 && smem <= 96 * 1024;
}

LDC_API int ldconv_offset_conv_bwd_tc_acc16(const float* grad_off, const void* x, const float* w, void* grad_x, float* grad_w,
                                            float* grad_b, void* workspace, size_t workspace_bytes, int B, int C, int H, int W,
                                            int N, int s, void* stream)
{
    LDC_REQUIRE(grad_off && x && w && workspace, "ldconv_offset_conv_bwd_tc_acc16: null pointer");
    LDC_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1 && N >= 1 && N <= 16 && s >= 1, "ldconv_offset_conv_bwd_tc_acc16: bad dims");
    if (B == 0) return LDCONV_OK;
    return offconv_bwd_tc(grad_off, (const __nv_bfloat16*)x, w, nullptr, (__nv_bfloat16*)grad_x, grad_w, grad_b, workspace,
                          workspace_bytes, B, C, H, W, N, s, (cudaStream_t)stream);
}
