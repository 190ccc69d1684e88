// ldconv_fused.cu -- whole-module inference kernels (sm_100a): the resampled operand never reaches HBM.
//
// Replaces, in ONE launch, /root/reference/ultralytics/nn/modules/conv.py:366-410 for eval mode: offset conv (:368),
// sampling grid / clamps / bilinear weights (:369-393), four gathers + bilinear sum + rearrange (:396-407) and the
// (N,1) conv + folded BatchNorm + SiLU (:355,408).  Two kernels behind the one C-ABI entry point ldconv_fused_fwd:
//
//   * small-C kernel (C <= 4; the first layer, 3 -> 16 channels: 23 % of the model's gather bytes and far too narrow for
//     128-bit NHWC vectors or for an MMA): one thread per output pixel on CUDA cores, x read through L1.
//   * tcgen05 kernel (C % 16 == 0, K = N*C <= 512, O % 16 == 0, O <= 256), see below.
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

// =====================================================================================================================
// small-C kernel: thread = output pixel.  K = N*C <= 36 samples stay in registers, W (O x K) and the offset-conv weights
// sit in shared memory as fp32.
// =====================================================================================================================
// NFIX / OFIX: num_param and outc as compile-time constants (0 = run-time values N_rt / O_rt).  The model's first layer
// (N=3, O=16) gets the exact instantiation: with run-time bounds the unrolled loops executed ~3600 mostly predicated-off
// instructions per pixel and missed the instruction cache (profiles/r1_ncu_smallc_before.csv).
template <typename T, int C, int NMAX, int NFIX, int OFIX>
__global__ void __launch_bounds__(128)
smallc_fused_kernel(const T* __restrict__ x, const float* __restrict__ w_off, const float* __restrict__ b_off,
                    const int* __restrict__ pn, const T* __restrict__ wt, const float* __restrict__ scale,
                    const float* __restrict__ shift, T* __restrict__ out, float* __restrict__ off_out, int B, int H, int W,
                    int h, int w, int N_rt, int s, int O_rt, int act)
{
    const int N = NFIX ? NFIX : N_rt;
    const int O = OFIX ? OFIX : O_rt;
    constexpr int NLOOP = NFIX ? NFIX : NMAX;
    extern __shared__ __align__(16) float smem_f[];
    const int O2 = 2 * N, K = N * C;
    float* s_woff = smem_f;                                  // [9][C][O2]
    float* s_wt = s_woff + ((9 * C * O2 + 3) & ~3);          // [K][O]  (transposed; 16-byte aligned rows: O % 4 == 0)
    float* s_sc = s_wt + K * O;                              // [O] scale, [O] shift
    for (int t = threadIdx.x; t < 9 * C * O2; t += blockDim.x) s_woff[t] = w_off[t];
    for (int t = threadIdx.x; t < K * O; t += blockDim.x) {
        const int k = t / O, o = t % O;
        s_wt[t] = Elem<T>::to_f(wt[(size_t)o * K + k]);
    }
    for (int t = threadIdx.x; t < O; t += blockDim.x) {
        s_sc[t] = scale ? scale[t] : 1.f;
        s_sc[O + t] = shift ? shift[t] : 0.f;
    }
    __syncthreads();

    const long long M = (long long)B * h * w;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int j = (int)(m % w);
    const int i = (int)((m / w) % h);
    const int b = (int)(m / ((long long)w * h));
    const T* xb = x + (size_t)b * H * W * C;

    // ---- offset conv (conv.py:368): 3x3 / pad 1 / stride s, fp32 accumulation ------------------------------------------
    // accumulators as packed fp32 pairs (FFMA2: two FMAs per issued instruction, each lane rounds like the scalar fmaf; the
    // kernel is issue-bound, profiles/r1_ncu_smallcL0b.txt); pair o2 = outputs (2 o2, 2 o2 + 1), rows of s_woff are 8-byte aligned
    uint64_t offp[NLOOP];
#pragma unroll
    for (int o2 = 0; o2 < NLOOP; ++o2)
        offp[o2] = f2_pack((2 * o2 < O2 && b_off) ? b_off[2 * o2] : 0.f, (2 * o2 + 1 < O2 && b_off) ? b_off[2 * o2 + 1] : 0.f);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int r = i * s + tap / 3 - 1, k = j * s + tap % 3 - 1;
        if (r < 0 || r >= H || k < 0 || k >= W) continue;
        const T* xp = xb + ((size_t)r * W + k) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float xv = Elem<T>::to_f(xp[c]);
            const uint64_t xx = f2_pack(xv, xv);
            const float2* wp = reinterpret_cast<const float2*>(s_woff + (tap * C + c) * O2);
#pragma unroll
            for (int o2 = 0; o2 < NLOOP; ++o2)
                if (2 * o2 < O2) { const float2 wv = wp[o2]; offp[o2] = f2_fma(xx, f2_pack(wv.x, wv.y), offp[o2]); }
        }
    }
    float offv[2 * NLOOP];
#pragma unroll
    for (int o2 = 0; o2 < NLOOP; ++o2) f2_unpack(offp[o2], offv[2 * o2], offv[2 * o2 + 1]);
    if (off_out) {
        float* op = off_out + (size_t)m * O2;
#pragma unroll
        for (int o = 0; o < 2 * NLOOP; ++o)
            if (o < O2) op[o] = offv[o];
    }

    // ---- sampling + (N,1) conv, one sample at a time (conv.py:369-408) ----------------------------------------------------
    constexpr int OMAX = OFIX ? OFIX : 32;
    uint64_t accp[OMAX / 2];
#pragma unroll
    for (int o = 0; o < OMAX / 2; ++o) accp[o] = 0ull;
#pragma unroll
    for (int n = 0; n < NLOOP; ++n) {
        if (n >= N) break;
        const SamplePoint q = make_point(i, j, s, pn[n], pn[N + n], offv[n], offv[N + n], H, W);
        const float g_lt = __fmul_rn(q.ar0, q.ak0), g_rb = __fmul_rn(q.ar1, q.ak1);
        const float g_lb = __fmul_rn(q.ar0, q.ak1), g_rt = __fmul_rn(q.ar1, q.ak0);
        const T* p00 = xb + ((size_t)q.r0 * W + q.k0) * C;
        const T* p11 = xb + ((size_t)q.r1 * W + q.k1) * C;
        const T* p01 = xb + ((size_t)q.r0 * W + q.k1) * C;
        const T* p10 = xb + ((size_t)q.r1 * W + q.k0) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v = bilinear(g_lt, g_rb, g_lb, g_rt, Elem<T>::to_f(p00[c]), Elem<T>::to_f(p11[c]), Elem<T>::to_f(p01[c]),
                               Elem<T>::to_f(p10[c]));
            v = Elem<T>::to_f(Elem<T>::from_f(v));       // the operand is rounded to the activation dtype, as in the 3-kernel path
            const uint64_t vv = f2_pack(v, v);
            const float4* wrow = reinterpret_cast<const float4*>(s_wt + (n * C + c) * O);
#pragma unroll
            for (int o4 = 0; o4 < OMAX / 4; ++o4) {
                if (o4 * 4 < O) {
                    const float4 wv = wrow[o4];
                    accp[o4 * 2 + 0] = f2_fma(vv, f2_pack(wv.x, wv.y), accp[o4 * 2 + 0]);
                    accp[o4 * 2 + 1] = f2_fma(vv, f2_pack(wv.z, wv.w), accp[o4 * 2 + 1]);
                }
            }
        }
    }
    // ---- folded BatchNorm + SiLU, 16-byte stores ----------------------------------------------------------------------
    float acc[OMAX];
#pragma unroll
    for (int o = 0; o < OMAX / 2; ++o) f2_unpack(accp[o], acc[2 * o], acc[2 * o + 1]);
    T* dst = out + (size_t)m * O;
    constexpr int V = Vec16<T>::N;
#pragma unroll
    for (int o0 = 0; o0 < OMAX; o0 += V) {
        if (o0 < O) {
            float y[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float z = fmaf(acc[o0 + e], s_sc[o0 + e], s_sc[O + o0 + e]);
                y[e] = act == LDCONV_ACT_SILU ? (sizeof(T) == 2 ? silu_fast(z) : silu(z)) : z;   // bf16: one MUFU, error below bf16's ulp
            }
            Vec16<T>::store(dst + o0, y);
        }
    }
}

// =====================================================================================================================
// small-C kernel, tiled: a 256-thread CTA owns a 16 x 16 tile of output pixels; the input footprint of the tile (+2 pixel
// halo) is staged once in shared memory as fp32 (coalesced 2-byte loads, zero outside the image = the offset conv's
// padding), then every thread runs the offset conv, the N samples (corners from the staged tile, from global memory when an
// offset leaves the halo) and the K x O contraction out of shared memory / registers.  This is the first layer of the
// model (C=3 -> 16, 640^2 -> 320^2): 23 % of the model's LDConv bytes.
// =====================================================================================================================
template <typename T, int C, int NMAX>
__global__ void __launch_bounds__(256, 2)
smallc_tiled_kernel(const T* __restrict__ x, const float* __restrict__ w_off, const float* __restrict__ b_off,
                    const int* __restrict__ pn, const T* __restrict__ wt, const float* __restrict__ scale,
                    const float* __restrict__ shift, T* __restrict__ out, float* __restrict__ off_out, int H, int W, int h,
                    int w, int N, int s, int O, int act, int THin, int TWin, int tiles_h, int tiles_w)
{
    constexpr int TS = 16, HALO = 2;
    extern __shared__ __align__(16) float smem_f[];
    const int O2 = 2 * N, K = N * C;
    const int O2P = (O2 + 3) & ~3;                              // offset-conv weights padded to float4 rows
    float* s_x = smem_f;                                        // [THin][TWin][C] fp32
    float* s_woff = s_x + ((THin * TWin * C + 3) & ~3);          // [9*C][O2P]
    float* s_wt = s_woff + 9 * C * O2P;                         // [K][O]
    float* s_sc = s_wt + K * O;                                 // [O] scale, [O] shift
    const int tj = blockIdx.x % tiles_w;
    const int ti = (blockIdx.x / tiles_w) % tiles_h;
    const int b = blockIdx.x / (tiles_w * tiles_h);
    const int i0 = ti * TS, j0 = tj * TS;
    const int r_org = i0 * s - HALO, k_org = j0 * s - HALO;
    const T* xb = x + (size_t)b * H * W * C;

    for (int t = threadIdx.x; t < THin * TWin * C; t += blockDim.x) {
        const int c = t % C, kk = (t / C) % TWin, rr = t / (C * TWin);
        const int r = r_org + rr, k = k_org + kk;
        s_x[t] = (r >= 0 && r < H && k >= 0 && k < W) ? Elem<T>::to_f(xb[((size_t)r * W + k) * C + c]) : 0.f;
    }
    for (int t = threadIdx.x; t < 9 * C * O2P; t += blockDim.x) {
        const int o = t % O2P, tc = t / O2P;
        s_woff[t] = o < O2 ? w_off[tc * O2 + o] : 0.f;
    }
    for (int t = threadIdx.x; t < K * O; t += blockDim.x) {
        const int k = t / O, o = t % O;
        s_wt[t] = Elem<T>::to_f(wt[(size_t)o * K + k]);
    }
    for (int t = threadIdx.x; t < O; t += blockDim.x) {
        s_sc[t] = scale ? scale[t] : 1.f;
        s_sc[O + t] = shift ? shift[t] : 0.f;
    }
    __syncthreads();

    const int pi = threadIdx.x / TS, pj = threadIdx.x % TS;
    const int i = i0 + pi, j = j0 + pj;
    if (i >= h || j >= w) return;
    const size_t m = ((size_t)b * h + i) * w + j;

    // ---- offset conv (conv.py:368) from the staged tile -------------------------------------------------------------------
    float offv[2 * NMAX];
#pragma unroll
    for (int o = 0; o < 2 * NMAX; ++o) offv[o] = (o < O2 && b_off) ? b_off[o] : 0.f;
    {
        const float* xp = s_x + ((pi * s + HALO - 1) * TWin + (pj * s + HALO - 1)) * C;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float* xt = xp + ((tap / 3) * TWin + (tap % 3)) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float xv = xt[c];
                const float4* w4 = reinterpret_cast<const float4*>(s_woff + (tap * C + c) * O2P);
#pragma unroll
                for (int o4 = 0; o4 < (2 * NMAX + 3) / 4; ++o4) {
                    if (o4 * 4 < O2) {
                        const float4 wv = w4[o4];
                        offv[o4 * 4 + 0] = fmaf(xv, wv.x, offv[o4 * 4 + 0]);
                        if (o4 * 4 + 1 < 2 * NMAX) offv[o4 * 4 + 1] = fmaf(xv, wv.y, offv[o4 * 4 + 1]);
                        if (o4 * 4 + 2 < 2 * NMAX) offv[o4 * 4 + 2] = fmaf(xv, wv.z, offv[o4 * 4 + 2]);
                        if (o4 * 4 + 3 < 2 * NMAX) offv[o4 * 4 + 3] = fmaf(xv, wv.w, offv[o4 * 4 + 3]);
                    }
                }
            }
        }
    }
    if (off_out) {
        float* op = off_out + m * O2;
#pragma unroll
        for (int o = 0; o < 2 * NMAX; ++o)
            if (o < O2) op[o] = offv[o];
    }

    // ---- sampling + (N,1) conv (conv.py:369-408) ------------------------------------------------------------------------------
    constexpr int OMAX = 32;
    float acc[OMAX];
#pragma unroll
    for (int o = 0; o < OMAX; ++o) acc[o] = 0.f;
    const int r_end = r_org + THin, k_end = k_org + TWin;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        if (n >= N) break;
        const SamplePoint q = make_point(i, j, s, pn[n], pn[N + n], offv[n], offv[N + n], H, W);
        const float g_lt = __fmul_rn(q.ar0, q.ak0), g_rb = __fmul_rn(q.ar1, q.ak1);
        const float g_lb = __fmul_rn(q.ar0, q.ak1), g_rt = __fmul_rn(q.ar1, q.ak0);
        const bool inside = q.r0 >= r_org && q.r1 < r_end && q.k0 >= k_org && q.k1 < k_end;
        float x00[C], x11[C], x01[C], x10[C];
        if (inside) {
            const int ra = (q.r0 - r_org) * TWin, rb = (q.r1 - r_org) * TWin, ka = q.k0 - k_org, kb = q.k1 - k_org;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                x00[c] = s_x[(ra + ka) * C + c]; x11[c] = s_x[(rb + kb) * C + c];
                x01[c] = s_x[(ra + kb) * C + c]; x10[c] = s_x[(rb + ka) * C + c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                x00[c] = Elem<T>::to_f(xb[((size_t)q.r0 * W + q.k0) * C + c]); x11[c] = Elem<T>::to_f(xb[((size_t)q.r1 * W + q.k1) * C + c]);
                x01[c] = Elem<T>::to_f(xb[((size_t)q.r0 * W + q.k1) * C + c]); x10[c] = Elem<T>::to_f(xb[((size_t)q.r1 * W + q.k0) * C + c]);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v = bilinear(g_lt, g_rb, g_lb, g_rt, x00[c], x11[c], x01[c], x10[c]);
            v = Elem<T>::to_f(Elem<T>::from_f(v));       // the operand is rounded to the activation dtype, as in the 3-kernel path
            const float4* wrow = reinterpret_cast<const float4*>(s_wt + (n * C + c) * O);
#pragma unroll
            for (int o4 = 0; o4 < OMAX / 4; ++o4) {
                if (o4 * 4 < O) {
                    const float4 wv = wrow[o4];
                    acc[o4 * 4 + 0] = fmaf(v, wv.x, acc[o4 * 4 + 0]);
                    acc[o4 * 4 + 1] = fmaf(v, wv.y, acc[o4 * 4 + 1]);
                    acc[o4 * 4 + 2] = fmaf(v, wv.z, acc[o4 * 4 + 2]);
                    acc[o4 * 4 + 3] = fmaf(v, wv.w, acc[o4 * 4 + 3]);
                }
            }
        }
    }
    T* dst = out + m * O;
    constexpr int V = Vec16<T>::N;
    const bool fast = sizeof(T) == 2;
#pragma unroll
    for (int o0 = 0; o0 < OMAX; o0 += V) {
        if (o0 < O) {
            float y[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float z = fmaf(acc[o0 + e], s_sc[o0 + e], s_sc[O + o0 + e]);
                y[e] = act == LDCONV_ACT_SILU ? (fast ? silu_fast(z) : silu(z)) : z;
            }
            Vec16<T>::store(dst + o0, y);
        }
    }
}

template <typename T, int C>
static int launch_smallc(const T* x, const float* w_off, const float* b_off, const int* pn, const T* wt, const float* scale,
                         const float* shift, T* out, float* off_out, int B, int H, int W, int N, int s, int O, int act,
                         cudaStream_t st)
{
    const int h = out_size(H, s), w = out_size(W, s);
    int max_r = 0, max_k = 0;
    {
        int32_t table[64];
        if (ldconv_p_n(N, table) == LDCONV_OK)
            for (int n = 0; n < N; ++n) { max_r = table[n] > max_r ? table[n] : max_r; max_k = table[N + n] > max_k ? table[N + n] : max_k; }
    }
    const int THin = 15 * s + 2 + max_r + 4, TWin = 15 * s + 2 + max_k + 4;
    const int tiles_h = (h + 15) / 16, tiles_w = (w + 15) / 16;
    const int O2P = (2 * N + 3) & ~3;
    const size_t smem = (size_t)(((THin * TWin * C + 3) & ~3) + 9 * C * O2P + N * C * O + 2 * O) * sizeof(float);
    const long long ctas = (long long)B * tiles_h * tiles_w;
    static const int use_tiled = getenv("LDCONV_SMALLC_TILED") ? atoi(getenv("LDCONV_SMALLC_TILED")) : 0;   // measured slower (1448 vs 1124 us on layer 0)
    if (use_tiled && smem <= 96 * 1024 && ctas <= 0x7fffffffll) {
        auto kern = smallc_tiled_kernel<T, C, 9>;
        LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)ctas, 256, smem, st>>>(x, w_off, b_off, pn, wt, scale, shift, out, off_out, H, W, h, w, N, s, O, act, THin,
                                                TWin, tiles_h, tiles_w);
        LDC_LAUNCH_CHECK("smallc_tiled_kernel");
    } else {
        const long long M = (long long)B * h * w;
        const size_t smem1 = (size_t)(((9 * C * 2 * N + 3) & ~3) + N * C * O + 2 * O) * sizeof(float);
        if (N == 3 && O == 16) {
            auto kern = smallc_fused_kernel<T, C, 9, 3, 16>;
            kern<<<cdiv(M, 128), 128, smem1, st>>>(x, w_off, b_off, pn, wt, scale, shift, out, off_out, B, H, W, h, w, N, s, O, act);
        } else {
            auto kern = smallc_fused_kernel<T, C, 9, 0, 0>;
            kern<<<cdiv(M, 128), 128, smem1, st>>>(x, w_off, b_off, pn, wt, scale, shift, out, off_out, B, H, W, h, w, N, s, O, act);
        }
        LDC_LAUNCH_CHECK("smallc_fused_kernel");
    }
    set_impl(LDCONV_IMPL_FFMA);
    return LDCONV_OK;
}

template <typename T>
static int dispatch_smallc(const void* x, const float* w_off, const float* b_off, const int* pn, const void* wt,
                           const float* scale, const float* shift, void* out, float* off_out, int B, int C, int H, int W,
                           int N, int s, int O, int act, cudaStream_t st)
{
    const T* xx = (const T*)x;
    const T* ww = (const T*)wt;
    T* oo = (T*)out;
    switch (C) {
        case 1: return launch_smallc<T, 1>(xx, w_off, b_off, pn, ww, scale, shift, oo, off_out, B, H, W, N, s, O, act, st);
        case 2: return launch_smallc<T, 2>(xx, w_off, b_off, pn, ww, scale, shift, oo, off_out, B, H, W, N, s, O, act, st);
        case 3: return launch_smallc<T, 3>(xx, w_off, b_off, pn, ww, scale, shift, oo, off_out, B, H, W, N, s, O, act, st);
        case 4: return launch_smallc<T, 4>(xx, w_off, b_off, pn, ww, scale, shift, oo, off_out, B, H, W, N, s, O, act, st);
        default: return fail(LDCONV_E_ARG, "small-C fused kernel: C=%d", C);
    }
}

int umma_fused_supported(int B, int C, int H, int W, int N, int s, int O, int dtype);
int umma_fused_fwd(const void* x, const float* w_off, const float* b_off, const int* pn, const void* wt,
                   const float* scale, const float* shift, void* out, float* off_out, int B, int C, int H, int W, int N,
                   int s, int O, int act, cudaStream_t st);

}  // namespace ldc

using namespace ldc;

LDC_API int ldconv_fused_supported(int B, int C, int H, int W, int N, int s, int O, int dtype)
{
    if (B < 0 || C < 1 || H < 1 || W < 1 || N < 1 || s < 1 || O < 1) return 0;
    if (dtype != LDCONV_F32 && dtype != LDCONV_BF16) return 0;
    const int V = dtype == LDCONV_BF16 ? 8 : 4;
    if (C <= 4 && N <= 9 && O <= 32 && O % V == 0) return 1;
    return umma_fused_supported(B, C, H, W, N, s, O, dtype);
}

LDC_API int ldconv_fused_fwd(const void* x, const float* w_off, const float* b_off, const int32_t* p_n, const void* wt,
                             const float* scale, const float* shift, void* out, float* off_out, int B, int C, int H, int W,
                             int N, int s, int O, int act, int dtype, void* stream)
{
    LDC_REQUIRE(x && w_off && p_n && wt && out, "ldconv_fused_fwd: null pointer");
    LDC_REQUIRE(act == LDCONV_ACT_NONE || act == LDCONV_ACT_SILU, "ldconv_fused_fwd: unknown activation %d", act);
    if (!ldconv_fused_supported(B, C, H, W, N, s, O, dtype))
        return fail(LDCONV_E_ARG,
                    "ldconv_fused_fwd: shape C=%d num_param=%d O=%d dtype=%d is outside the fused kernels' range "
                    "(query ldconv_fused_supported and use the offset_conv / gather / gemm entry points)", C, N, O, dtype);
    if (B == 0) return LDCONV_OK;
    LDC_REQUIRE(aligned16(out), "ldconv_fused_fwd: out must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 4) {
        if (dtype == LDCONV_BF16)
            return dispatch_smallc<__nv_bfloat16>(x, w_off, b_off, p_n, wt, scale, shift, out, off_out, B, C, H, W, N, s, O,
                                                  act, st);
        return dispatch_smallc<float>(x, w_off, b_off, p_n, wt, scale, shift, out, off_out, B, C, H, W, N, s, O, act, st);
    }
    return umma_fused_fwd(x, w_off, b_off, p_n, wt, scale, shift, out, off_out, B, C, H, W, N, s, O, act, st);
}
