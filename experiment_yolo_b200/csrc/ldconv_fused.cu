// ldconv_fused.cu -- whole-module inference kernels (sm_100a): the resampled operand never reaches HBM.
//
// Replaces, in ONE launch, /root/reference/ultralytics/nn/modules/conv.py:366-410 for eval mode: offset conv (:368),
// sampling grid / clamps / bilinear weights (:369-393), four gathers + bilinear sum + rearrange (:396-407) and the
// (N,1) conv + folded BatchNorm + SiLU (:355,408).  Two kernels behind the one C-ABI entry point ldconv_fused_fwd:
//
//   * small-C kernel (C <= 4; the first layer, 3 -> 16 channels: 23 % of the model's gather bytes and far too narrow for
//     128-bit NHWC vectors or for an MMA): one thread per output pixel on CUDA cores, x read through L1.
//   * tcgen05 kernel (C % 16 == 0, K = N*C <= 512, O % 16 == 0, O <= 256), see below.
#include <mutex>

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

// =====================================================================================================================
// first-layer kernel (bf16, C = 3 -> O = 16, num_param = 3, stride 2, even H and W): rows of output pixels per CTA.
// The thread-per-pixel kernel above executed ~1570 instructions per pixel of which ~150 are the packed FMAs: a quarter went
// into the per-CTA weight staging and the 64-bit index decomposition, another quarter into the addressing of 63 two-byte
// loads (profiles/r1_ncu_smallcL0b.txt: 79 % of the issue slots, DRAM at 10 %).  Here
//   * a CTA stages the weights ONCE and then walks over whole output rows (grid = resident CTAs, row = blockIdx.x + k * grid),
//   * the 3 x 3 x 3 window of the offset conv (conv.py:368) is read as five aligned 4-byte words per input row (the 18
//     bytes of three NHWC pixels starting at byte 12 j - 6 sit inside the 20 bytes from 12 j - 8), the left / top padding
//     is a predicate on the words,
//   * the offset conv contracts PAIRS of window elements per packed FMA (the two halves of a word are the two lanes of the
//     FFMA2, weights pre-paired in shared memory, the two partial sums added once at the end): no broadcast moves,
//   * corner addresses are 32-bit element indices relative to the image.
// Sampling arithmetic is make_point / bilinear as everywhere else, the operand is rounded to bf16 like the 3-kernel path.
// =====================================================================================================================
constexpr int L0_KP = 14;      // pairs of the 27 window elements: 12 word pairs + (e1 row 0, e1 row 1) + (e1 row 2, 1 * bias)

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// the layer's weights in the order the kernel consumes them (1408 bytes)
struct L0Weights {
    float wo[L0_KP * 3 * 4];     // [pair][q] -> (wA[2q], wB[2q], wA[2q+1], wB[2q+1]) of the offset conv
    float wt[9 * 16];            // [k = n*C + c][o] of the (N,1) conv
    float sc[2 * 16];            // scale | shift (halved for SiLU: 0.5 z (1 + tanh(0.5 z)))
    int pn[6];
    int pad[2];
};
// Uniform weights read from shared memory cost one L1 wavefront per 8 bytes per warp: 174 of the rows kernel's 336 wavefronts
// per 32 pixels, and the L1 data pipe is what bounds it (profiles/r1_ncu_l0rows_v1.txt: 83.5 % of peak, issue slots 57 %).
// From the constant bank they are operands of the FFMA2s and cost nothing.  The library keeps no per-module state, so the
// bank has L0_SLOTS slots per device, a slot belongs to the first set of argument pointers (one layer's parameters) that asks for it for
// the life of the process, and its content is rewritten in stream order by l0_prep_kernel before EVERY launch (the same
// bytes unless the caller changed the weights); a fifth distinct layer on a device uses the shared-memory variant (SLOT -1).
constexpr int L0_SLOTS = 4;
__constant__ L0Weights c_l0[L0_SLOTS];

__device__ __forceinline__ void l0_fill(L0Weights* dst, const float* __restrict__ w_off, const float* __restrict__ b_off,
                                        const int* __restrict__ pn, const __nv_bfloat16* __restrict__ wt,
                                        const float* __restrict__ scale, const float* __restrict__ shift, int act)
{
    constexpr int O = 16, O2 = 6, K = 9;
    for (int t = threadIdx.x; t < L0_KP * 12; t += blockDim.x) {
        const int p = t / 12, q = (t % 12) / 4, e = t % 4;
        const int o = 2 * q + (e >> 1), second = e & 1;
        int k;                                               // window element tr*9 + tc*3 + c (w_off is [3][3][C][2N])
        if (p < 12) k = (p / 4) * 9 + 2 * (p % 4) + 1 + second;
        else if (p == 12) k = second ? 9 : 0;
        else k = second ? -1 : 18;                           // the pad element is the constant 1: its weight is the bias
        dst->wo[t] = k >= 0 ? w_off[k * O2 + o] : (b_off ? b_off[o] : 0.f);
    }
    for (int t = threadIdx.x; t < K * O; t += blockDim.x) dst->wt[t] = __bfloat162float(wt[(t % O) * K + t / O]);
    const float half = act == LDCONV_ACT_SILU ? 0.5f : 1.f;
    for (int t = threadIdx.x; t < O; t += blockDim.x) {
        dst->sc[t] = half * (scale ? scale[t] : 1.f);
        dst->sc[O + t] = half * (shift ? shift[t] : 0.f);
    }
    if (threadIdx.x < 6) dst->pn[threadIdx.x] = pn[threadIdx.x];
}

__global__ void l0_prep_kernel(L0Weights* dst, const float* __restrict__ w_off, const float* __restrict__ b_off,
                               const int* __restrict__ pn, const __nv_bfloat16* __restrict__ wt,
                               const float* __restrict__ scale, const float* __restrict__ shift, int act)
{
    l0_fill(dst, w_off, b_off, pn, wt, scale, shift, act);
}

template <int MINB, int SLOT>
__global__ void __launch_bounds__(160, MINB)
l0_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_off, const float* __restrict__ b_off,
               const int* __restrict__ pn, const __nv_bfloat16* __restrict__ wt, const float* __restrict__ scale,
               const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, float* __restrict__ off_out, int rows, int H,
               int W, int h, int w, int act)
{
    constexpr int C = 3, N = 3, O = 16, O2 = 6;
    __shared__ __align__(16) L0Weights s_w;
    if (SLOT < 0) {
        l0_fill(&s_w, w_off, b_off, pn, wt, scale, shift, act);
        __syncthreads();
    }
    const float* s_wo = SLOT < 0 ? s_w.wo : c_l0[SLOT < 0 ? 0 : SLOT].wo;
    const float* s_wt = SLOT < 0 ? s_w.wt : c_l0[SLOT < 0 ? 0 : SLOT].wt;
    const float* s_sc = SLOT < 0 ? s_w.sc : c_l0[SLOT < 0 ? 0 : SLOT].sc;
    const int* s_pn = SLOT < 0 ? s_w.pn : c_l0[SLOT < 0 ? 0 : SLOT].pn;

    const int j = blockIdx.x * blockDim.x + threadIdx.x;     // a thread keeps its column and walks down the rows
    if (j >= w) return;
    const float hm = (float)(H - 1), wm = (float)(W - 1);
    const uint32_t row_words = (uint32_t)W * C / 2;          // W even: rows are whole 4-byte words
    const uint32_t* xcol = reinterpret_cast<const uint32_t*>(x) + 3 * j;          // word of byte 12 j of a row
    const unsigned short* xs0 = reinterpret_cast<const unsigned short*>(x);
    int b = blockIdx.y / h, i = blockIdx.y - b * h;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        {
            // ---- offset conv ------------------------------------------------------------------------------------------
            uint64_t accp[O2];
#pragma unroll
            for (int o = 0; o < O2; ++o) accp[o] = 0ull;
            float e1[3];
#pragma unroll
            for (int tr = 0; tr < 3; ++tr) {
                uint32_t wd[5] = {0u, 0u, 0u, 0u, 0u};
                if (tr > 0 || i > 0) {                                            // input row 2 i - 1 + tr of image b (H = 2 h)
                    const uint32_t* rp = xcol + (size_t)(uint32_t)(2 * row - 1 + tr) * row_words;
                    if (j > 0) { wd[0] = rp[-2]; wd[1] = rp[-1]; }
                    wd[2] = rp[0]; wd[3] = rp[1]; wd[4] = rp[2];
                }
                e1[tr] = bf16_hi(wd[0]);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint64_t xx = f2_pack(bf16_lo(wd[q4 + 1]), bf16_hi(wd[q4 + 1]));
                    const float4* wp = reinterpret_cast<const float4*>(s_wo + (tr * 4 + q4) * 12);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 wv = wp[q];
                        accp[2 * q] = f2_fma(xx, f2_pack(wv.x, wv.y), accp[2 * q]);
                        accp[2 * q + 1] = f2_fma(xx, f2_pack(wv.z, wv.w), accp[2 * q + 1]);
                    }
                }
            }
#pragma unroll
            for (int p = 12; p < 14; ++p) {
                const uint64_t xx = p == 12 ? f2_pack(e1[0], e1[1]) : f2_pack(e1[2], 1.f);
                const float4* wp = reinterpret_cast<const float4*>(s_wo + p * 12);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 wv = wp[q];
                    accp[2 * q] = f2_fma(xx, f2_pack(wv.x, wv.y), accp[2 * q]);
                    accp[2 * q + 1] = f2_fma(xx, f2_pack(wv.z, wv.w), accp[2 * q + 1]);
                }
            }
            float offv[O2];
#pragma unroll
            for (int o = 0; o < O2; ++o) {
                float lo, hi;
                f2_unpack(accp[o], lo, hi);
                offv[o] = lo + hi;
            }
            const size_t m = (size_t)row * w + j;
            const unsigned short* xs = xs0 + (size_t)b * ((size_t)H * W * C);
            if (off_out) {
                float2* op = reinterpret_cast<float2*>(off_out + m * O2);
                op[0] = make_float2(offv[0], offv[1]);
                op[1] = make_float2(offv[2], offv[3]);
                op[2] = make_float2(offv[4], offv[5]);
            }

            // ---- sampling + (N,1) conv -----------------------------------------------------------------------------------
            uint64_t acc[O / 2];
#pragma unroll
            for (int o = 0; o < O / 2; ++o) acc[o] = 0ull;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const SamplePoint q = make_point_grid(2 * i + s_pn[n], 2 * j + s_pn[N + n], offv[n], offv[N + n], hm, wm);
                const float g_lt = __fmul_rn(q.ar0, q.ak0), g_rb = __fmul_rn(q.ar1, q.ak1);
                const float g_lb = __fmul_rn(q.ar0, q.ak1), g_rt = __fmul_rn(q.ar1, q.ak0);
                const uint32_t ra = (uint32_t)q.r0 * (uint32_t)(W * C), rb = (uint32_t)q.r1 * (uint32_t)(W * C);
                const uint32_t ka = (uint32_t)q.k0 * C, kb = (uint32_t)q.k1 * C;
                const unsigned short* p00 = xs + (ra + ka);
                const unsigned short* p11 = xs + (rb + kb);
                const unsigned short* p01 = xs + (ra + kb);
                const unsigned short* p10 = xs + (rb + ka);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float v = bilinear(g_lt, g_rb, g_lb, g_rt, bf16_lo(p00[c]), bf16_lo(p11[c]), bf16_lo(p01[c]), bf16_lo(p10[c]));
                    v = __bfloat162float(__float2bfloat16_rn(v));     // the operand is rounded to bf16, as in the 3-kernel path
                    const uint64_t vv = f2_pack(v, v);
                    const float4* wrow = reinterpret_cast<const float4*>(s_wt + (n * C + c) * O);
#pragma unroll
                    for (int o4 = 0; o4 < O / 4; ++o4) {
                        const float4 wv = wrow[o4];
                        acc[o4 * 2 + 0] = f2_fma(vv, f2_pack(wv.x, wv.y), acc[o4 * 2 + 0]);
                        acc[o4 * 2 + 1] = f2_fma(vv, f2_pack(wv.z, wv.w), acc[o4 * 2 + 1]);
                    }
                }
            }
            // ---- folded BatchNorm + SiLU, two 16-byte stores ----------------------------------------------------------------
            uint32_t pk[O / 2];
#pragma unroll
            for (int o4 = 0; o4 < O / 4; ++o4) {
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + o4 * 4);
                const float4 sh = *reinterpret_cast<const float4*>(s_sc + O + o4 * 4);
                float z[4];
                f2_unpack(f2_fma(acc[o4 * 2 + 0], f2_pack(sc.x, sc.y), f2_pack(sh.x, sh.y)), z[0], z[1]);
                f2_unpack(f2_fma(acc[o4 * 2 + 1], f2_pack(sc.z, sc.w), f2_pack(sh.z, sh.w)), z[2], z[3]);
                if (act == LDCONV_ACT_SILU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {                  // z holds 0.5 (scale acc + shift): y = z tanh(z) + z
                        float t;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z[e]));
                        z[e] = fmaf(z[e], t, z[e]);
                    }
                }
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[o4 * 2 + 0]) : "f"(z[1]), "f"(z[0]));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[o4 * 2 + 1]) : "f"(z[3]), "f"(z[2]));
            }
            uint4* dst = reinterpret_cast<uint4*>(out + m * O);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        i += gridDim.y;
        while (i >= h) { i -= h; ++b; }
    }
}

static bool l0_rows_applicable(const void* x, int B, int C, int H, int W, int N, int s, int O)
{
    return C == 3 && N == 3 && O == 16 && s == 2 && H % 2 == 0 && W % 2 == 0 && ((uintptr_t)x & 3) == 0 &&
           (long long)H * W * C < (1ll << 31) && (long long)B * (H / 2) < (1ll << 31);
}

struct L0Slot { int dev, act; const void *w_off, *b_off, *pn, *wt, *scale, *shift; L0Weights* addr; };

// slot of this (device, weights) pair in the constant bank, or -1 (bank full: shared-memory variant)
static int l0_slot(int dev, int act, const void* w_off, const void* b_off, const void* pn, const void* wt, const void* scale,
                   const void* shift, L0Weights** addr)
{
    static std::mutex mu;
    static L0Slot table[64];
    static int used = 0;
    std::lock_guard<std::mutex> lock(mu);
    int on_dev = 0;
    for (int t = 0; t < used; ++t) {
        if (table[t].dev != dev) continue;
        const L0Slot& e = table[t];
        if (e.act == act && e.w_off == w_off && e.b_off == b_off && e.pn == pn && e.wt == wt && e.scale == scale && e.shift == shift) {
            *addr = e.addr;
            return on_dev;
        }
        ++on_dev;
    }
    if (on_dev >= L0_SLOTS || used >= 64) return -1;
    L0Weights* base = nullptr;
    if (cudaGetSymbolAddress((void**)&base, c_l0) != cudaSuccess) { cudaGetLastError(); return -1; }
    table[used] = L0Slot{dev, act, w_off, b_off, pn, wt, scale, shift, base + on_dev};
    *addr = table[used].addr;
    ++used;
    return on_dev;
}

typedef void (*l0_kernel_t)(const __nv_bfloat16*, const float*, const float*, const int*, const __nv_bfloat16*, const float*,
                            const float*, __nv_bfloat16*, float*, int, int, int, int, int, int);

static int launch_l0_rows(const __nv_bfloat16* x, const float* w_off, const float* b_off, const int* pn, const __nv_bfloat16* wt,
                          const float* scale, const float* shift, __nv_bfloat16* out, float* off_out, int B, int H, int W,
                          int act, cudaStream_t st)
{
    const int h = H / 2, w = W / 2, rows = B * h;
    const int threads = w % 160 == 0 ? 160 : (w >= 128 ? 128 : ((w + 31) / 32) * 32);
    // CTAs per SM the register budget is compiled for: 6 (64 registers) or 5 (80 registers)
    constexpr int minb = 6;      // measured: the 64-register build with six CTAs per SM (profiles/r1_l0_ab_s4.jsonl)
    int dev = 0;
    LDC_CUDA(cudaGetDevice(&dev));
    L0Weights* slot_addr = nullptr;
    const int slot = l0_slot(dev, act, w_off, b_off, pn, wt, scale, shift, &slot_addr);
    static const l0_kernel_t kerns[2][L0_SLOTS + 1] = {
        {l0_rows_kernel<6, -1>, l0_rows_kernel<6, 0>, l0_rows_kernel<6, 1>, l0_rows_kernel<6, 2>, l0_rows_kernel<6, 3>},
        {l0_rows_kernel<5, -1>, l0_rows_kernel<5, 0>, l0_rows_kernel<5, 1>, l0_rows_kernel<5, 2>, l0_rows_kernel<5, 3>}};
    const int v = minb == 5 ? 1 : 0;
    l0_kernel_t kern = kerns[v][slot + 1];
    static int sms = 0, per_sm[2][2][3] = {};
    const int tslot = threads == 160 ? 0 : (threads == 128 ? 1 : 2), cslot = slot >= 0 ? 1 : 0;
    if (!sms) LDC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!per_sm[v][cslot][tslot]) {
        int n = 0;
        LDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, tslot == 2 ? 128 : threads, 0));
        per_sm[v][cslot][tslot] = n < 1 ? 1 : n;
    }
    const int segs = (w + threads - 1) / threads;            // column segments: grid.x; grid.y CTAs share the rows of a segment
    long long gy = (long long)sms * per_sm[v][cslot][tslot] / segs;
    if (gy < 1) gy = 1;
    if (gy > rows) gy = rows;
    if (gy > 65535) gy = 65535;
    if (slot >= 0) {
        l0_prep_kernel<<<1, 192, 0, st>>>(slot_addr, w_off, b_off, pn, wt, scale, shift, act);
        LDC_LAUNCH_CHECK("l0_prep_kernel");
    }
    kern<<<dim3((unsigned)segs, (unsigned)gy), threads, 0, st>>>(x, w_off, b_off, pn, wt, scale, shift, out, off_out, rows, H, W, h,
                                                                  w, act);
    LDC_LAUNCH_CHECK("l0_rows_kernel");
    set_impl(LDCONV_IMPL_FFMA);
    return LDCONV_OK;
}

// =====================================================================================================================
// first-layer kernel on the tensor cores (bf16, C = 3 -> O = 16, num_param = 3, stride 2, even H and W).
// l0_rows_kernel above is issue-bound (73 % of the slots, 832 warp instructions per 32 pixels, profiles/r1_ncu_l0rows_v2_constbank.txt):
// 164 packed FMAs, 92 uniform weight loads and ~50 bf16 unpacks per pixel belong to the two small contractions (offset conv:
// 27 window elements -> 6 offsets; (N,1) conv: 9 samples -> 16 channels).  Both are MMAs with the pixel as the M row:
//   * a thread owns one output pixel = one row of the CTA's 128-row operand tiles = one TMEM lane of the accumulators;
//   * offset conv: the 3 x 3 x 3 window is read as five aligned 4-byte words per input row exactly as above; words 1..4 of the
//     three rows ARE already eight consecutive bf16 of the operand row (K = 8 tr + e - 1), the three leading half words are
//     packed into K = 24..26: four 16-byte shared-memory stores build the 64-byte row (SWIZZLE_64B, K = 32), no unpacking;
//     two tcgen05.mma (M 128, N 16, K 16) against the pre-arranged bf16 weights leave the six offsets in TMEM;
//   * sampling arithmetic and the 36 two-byte corner loads are unchanged (make_point_grid / bilinear, operand rounded to bf16);
//   * (N,1) conv: the nine bf16 samples are the 32-byte operand row (SWIZZLE_32B, K = 16), one tcgen05.mma, 16 accumulators per
//     thread back through tcgen05.ld, folded BatchNorm + SiLU, two 16-byte stores.
// A tile is 64 columns x 2 rows of output pixels (w = 320 = 5 x 64); CTAs are persistent over tiles; 4 warps = the 4 TMEM lane
// quarters.  The offset conv's weights reach the MMA as bf16 (the module's are bf16 already: exact), its bias is added in fp32.
// =====================================================================================================================
__device__ __forceinline__ uint64_t l0_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;                       // 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
    return d;
}
__device__ __forceinline__ uint32_t l0_swz(uint32_t lin, uint32_t mask) { return lin ^ (((lin >> 7) & mask) << 4); }

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

template <int OFS> __device__ __forceinline__ uint32_t lds_u16_at(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFS));
    return v;
}
template <int OFS> __device__ __forceinline__ uint32_t lds_u32_at(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFS));
    return v;
}

constexpr int L0T_THREADS = 128, L0T_COLS = 64;
// staged input tile of a 64 x 2 output tile: 10 input rows (2 i0 - 3 ...) x 216 four-byte words = 144 pixels (2 j0 - 8 ...); the image
// rows are mapped as rows of 4-byte words (W * 3 / 2 per row, box dims are limited to 256 elements), out-of-image parts are
// zero-filled by the TMA unit = the offset conv's zero padding.  The box must START on a 16-byte boundary of the row (a start at
// 8 mod 16 bytes raised "illegal instruction"): the column halo is 8 pixels = 48 bytes.
constexpr int L0T_TROWS = 10, L0T_TWORDS = 216, L0T_TPIX = 144, L0T_RHALO = 3, L0T_KHALO = 8;
static int g_l0_variant = 0;
static long long* g_l0_trace = nullptr;          // debug: clock64 stamps of CTA 0 / thread 0 (benchmarks/l0_ab.py --trace)

template <bool TRACE>
__global__ void __launch_bounds__(L0T_THREADS, 8)
l0_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_off, const float* __restrict__ b_off,
             const int* __restrict__ pn, const __nv_bfloat16* __restrict__ wt, const float* __restrict__ scale,
             const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, float* __restrict__ off_out, int rows, int H, int W,
             int h, int w, int act, int tiles_x, int num_tiles, long long* trace)
{
    constexpr int C = 3, N = 3, O = 16, O2 = 6;
    int tn = 0;
    auto stamp = [&](int it) {
        if constexpr (TRACE) {
            if (trace && blockIdx.x == 0 && threadIdx.x == 0 && it >= 2 && it < 6 && tn < 63) trace[tn++] = clock64();
        }
    };
    // shared memory: operand tiles (swizzle atoms: 512 B / 256 B; the arrays are 1024-byte aligned), weights, constants
    __shared__ __align__(1024) uint8_t s_a1[128 * 64];       // offset-conv operand: 128 pixels x K = 32 bf16, SWIZZLE_64B
    __shared__ __align__(1024) uint8_t s_a2[128 * 32];       // (N,1)-conv operand: 128 pixels x K = 16 bf16, SWIZZLE_32B
    __shared__ __align__(1024) uint8_t s_w1[16 * 64];        // offset-conv weights: 16 rows (6 used) x K = 32
    __shared__ __align__(1024) uint8_t s_w2[16 * 32];        // (N,1)-conv weights: 16 rows x K = 16
    __shared__ __align__(128) uint32_t s_x[L0T_TROWS * L0T_TWORDS];      // TMA-staged input tile
    __shared__ __align__(16) float s_sc[2 * O];              // scale | shift (halved for SiLU)
    __shared__ float s_b[8];
    __shared__ int s_pn[8];
    __shared__ __align__(8) uint64_t bar[3];                 // offsets ready, accumulator ready, input tile landed
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup ------------------------------------------------------------------------------------------------
    for (int t = tid; t < 16 * 32; t += L0T_THREADS) {        // W1[n][k]
        const int n = t >> 5, k = t & 31;
        float v = 0.f;
        if (n < O2) {
            int tr = -1, e = 0;
            if (k < 24) { tr = k >> 3; e = (k & 7) + 1; }
            else if (k < 27) { tr = k - 24; e = 0; }
            if (tr >= 0) v = w_off[(tr * 9 + e) * O2 + n];
        }
        *reinterpret_cast<__nv_bfloat16*>(s_w1 + l0_swz((uint32_t)(n * 64 + k * 2), 3u)) = __float2bfloat16_rn(v);
    }
    for (int t = tid; t < 16 * 16; t += L0T_THREADS) {        // W2[o][k]
        const int o = t >> 4, k = t & 15;
        *reinterpret_cast<__nv_bfloat16*>(s_w2 + l0_swz((uint32_t)(o * 32 + k * 2), 1u)) = k < N * C ? wt[o * (N * C) + k] : __float2bfloat16_rn(0.f);
    }
    {
        const float half = act == LDCONV_ACT_SILU ? 0.5f : 1.f;
        if (tid < O) {
            s_sc[tid] = half * (scale ? scale[tid] : 1.f);
            s_sc[O + tid] = half * (shift ? shift[tid] : 0.f);
        }
        if (tid < 8) {
            s_b[tid] = (tid < O2 && b_off) ? b_off[tid] : 0.f;
            s_pn[tid] = tid < 2 * N ? pn[tid] : 0;
        }
    }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    fence_proxy_async_smem();                                  // the weight tiles are read by the tensor core (async proxy)
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc = make_idesc_bf16(128, 16);
    const uint64_t dA1 = l0_desc(smem_u32(s_a1), 512, 4), dW1 = l0_desc(smem_u32(s_w1), 512, 4);
    const uint64_t dA2 = l0_desc(smem_u32(s_a2), 256, 6), dW2 = l0_desc(smem_u32(s_w2), 256, 6);
    const uint32_t a1_row = smem_u32(s_a1) + (uint32_t)tid * 64u, a2_row = smem_u32(s_a2) + (uint32_t)tid * 32u;
    const uint32_t sw1 = ((uint32_t)(tid >> 1) & 3u) << 4, sw2 = ((uint32_t)(tid >> 2) & 1u) << 4;     // XOR of this row's chunks
    const float hm = (float)(H - 1), wm = (float)(W - 1);
    const unsigned short* xs0 = reinterpret_cast<const unsigned short*>(x);
    const bool leader = warp == 0 && elect_one();
    const uint32_t tile_s = smem_u32(s_x);
    constexpr uint32_t kTileBytes = L0T_TROWS * L0T_TWORDS * 4, kRowB = L0T_TWORDS * 4;
    // tile -> (column segment, image, first output row of the pair); h is even (host), so a pair of output rows lies in one image.
    // The decomposition is kept incrementally (tile advances by gridDim.x): no division per tile.
    struct TilePos { int seg, b, i0; };
    const int step_pairs = (int)gridDim.x / tiles_x, step_segs = (int)gridDim.x - step_pairs * tiles_x;
    auto advance = [&](TilePos t) {
        t.seg += step_segs;
        t.i0 += 2 * step_pairs;
        if (t.seg >= tiles_x) { t.seg -= tiles_x; t.i0 += 2; }
        while (t.i0 >= h) { t.i0 -= h; ++t.b; }
        return t;
    };
    auto issue_tile = [&](const TilePos& t) {
        mbar_arrive_expect_tx(&bar[2], kTileBytes);
        tma_load_4d(s_x, &tmX, &bar[2], (t.seg * L0T_COLS * 2 - L0T_KHALO) * 3 / 2, 2 * t.i0 - L0T_RHALO, t.b, 0);
    };
    TilePos cur;
    {
        const int pair = (int)blockIdx.x / tiles_x;
        cur.seg = (int)blockIdx.x - pair * tiles_x;
        cur.b = (2 * pair) / h;
        cur.i0 = 2 * pair - cur.b * h;
    }
    if (leader && (int)blockIdx.x < num_tiles) issue_tile(cur);

    if (TRACE && trace && blockIdx.x == 0 && threadIdx.x == 0) trace[63] = gridDim.x;
    uint32_t ph = 0;
    int it = 0;
    const int jl = tid & (L0T_COLS - 1), rsel = tid >> 6;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ph ^= 1, ++it) {
        stamp(it);
        const TilePos nxt = advance(cur);
        const int j = cur.seg * L0T_COLS + jl;
        const int b = cur.b, i = cur.i0 + rsel;
        const int row = b * h + i;
        const bool valid = j < w && row < rows;
        const int r_org = 2 * cur.i0 - L0T_RHALO, k_org = cur.seg * L0T_COLS * 2 - L0T_KHALO;
        mbar_wait(&bar[2], ph);
        // ---- offset-conv operand row: the window's five words per input row from the staged tile (zero fill = padding) ----
        {
            // pixel 2 j - 1 starts at byte 6 (2 jl + L0T_KHALO - 1) of the tile row = 2 mod 4: the words start 2 bytes earlier
            const uint32_t wbase = tile_s + (uint32_t)(2 * rsel - 1 + L0T_RHALO) * kRowB + (uint32_t)(6 * (2 * jl + L0T_KHALO - 1) - 2);
            uint32_t e0[3];
#define L0T_WINDOW_ROW(TR)                                                                                                        \
    {                                                                                                                             \
        constexpr int RB_ = TR * (int)kRowB;                                                                                      \
        e0[TR] = lds_u32_at<RB_>(wbase);                                                                                          \
        const uint32_t w1_ = lds_u32_at<RB_ + 4>(wbase), w2_ = lds_u32_at<RB_ + 8>(wbase), w3_ = lds_u32_at<RB_ + 12>(wbase),     \
                       w4_ = lds_u32_at<RB_ + 16>(wbase);                                                                         \
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a1_row + (((uint32_t)TR << 4) ^ sw1)), "r"(w1_), "r"(w2_),  \
                     "r"(w3_), "r"(w4_) : "memory");                                                                              \
    }
            L0T_WINDOW_ROW(0)
            L0T_WINDOW_ROW(1)
            L0T_WINDOW_ROW(2)
#undef L0T_WINDOW_ROW
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a1_row + ((3u << 4) ^ sw1)), "r"(__byte_perm(e0[0], e0[1], 0x7632)),
                         "r"(e0[2] >> 16), "r"(0u), "r"(0u) : "memory");
        }
        stamp(it);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        stamp(it);
        if (leader) {
            tc_fence_after_sync();
            mma_bf16_ss(tmem_base, dA1, dW1, idesc, 0u);
            mma_bf16_ss(tmem_base, dA1 + 2, dW1 + 2, idesc, 1u);          // K 16..31: +32 bytes inside the swizzled rows
            mma_commit(&bar[0]);
        }
        __syncwarp();
        mbar_wait(&bar[0], ph);
        stamp(it);
        tc_fence_after_sync();
        float offv[O2];
        {
            uint32_t v[8];
            tmem_ld_32x32b_x8(lane_addr, v);
            tmem_ld_wait();
#pragma unroll
            for (int o = 0; o < O2; ++o) offv[o] = __uint_as_float(v[o]) + s_b[o];
        }
        stamp(it);
        const size_t m = (size_t)row * w + j;
        // ---- sampling -> (N,1)-conv operand row --------------------------------------------------------------------------
        if (valid) {
            if (off_out) {
                float2* op = reinterpret_cast<float2*>(off_out + m * O2);
                op[0] = make_float2(offv[0], offv[1]);
                op[1] = make_float2(offv[2], offv[3]);
                op[2] = make_float2(offv[4], offv[5]);
            }
            const unsigned short* xs = xs0 + (size_t)b * ((size_t)H * W * C);
            float sv[N * C + 1];
            sv[N * C] = 0.f;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const SamplePoint q = make_point_grid(2 * i + s_pn[n], 2 * j + s_pn[N + n], offv[n], offv[N + n], hm, wm);
                const float g_lt = __fmul_rn(q.ar0, q.ak0), g_rb = __fmul_rn(q.ar1, q.ak1);
                const float g_lb = __fmul_rn(q.ar0, q.ak1), g_rt = __fmul_rn(q.ar1, q.ak0);
                const int t0 = q.r0 - r_org, t1 = q.r1 - r_org, u0 = q.k0 - k_org, u1 = q.k1 - k_org;
                const bool inside = (unsigned)t0 < (unsigned)L0T_TROWS && (unsigned)t1 < (unsigned)L0T_TROWS &&
                                    (unsigned)u0 < (unsigned)L0T_TPIX && (unsigned)u1 < (unsigned)L0T_TPIX;
                if (inside) {          // the four corners from the staged tile (two-byte shared-memory loads)
                    const uint32_t a0 = tile_s + (uint32_t)t0 * kRowB, a1 = tile_s + (uint32_t)t1 * kRowB;
                    const uint32_t b0 = (uint32_t)u0 * 6u, b1 = (uint32_t)u1 * 6u;
                    const uint32_t p00 = a0 + b0, p11 = a1 + b1, p01 = a0 + b1, p10 = a1 + b0;
                    // a pixel = 6 bytes at 0 or 2 mod 4: always inside two aligned words (two loads per corner instead of three
                    // two-byte ones: a third fewer shared-memory wavefronts, the stride-12-byte lanes cost three per instruction)
                    float c00[3], c11[3], c01[3], c10[3];
                    auto corner = [&](uint32_t p, float (&e)[3]) {
                        const uint32_t pa = p & ~3u, sh = (p & 2u) << 3;
                        const uint32_t w0 = lds_u32_at<0>(pa), w1 = lds_u32_at<4>(pa);
                        const uint32_t xw = __funnelshift_r(w0, w1, sh), yw = w1 >> sh;
                        e[0] = __uint_as_float(xw << 16);
                        e[1] = __uint_as_float(xw & 0xffff0000u);
                        e[2] = __uint_as_float(yw << 16);
                    };
                    corner(p00, c00);
                    corner(p11, c11);
                    corner(p01, c01);
                    corner(p10, c10);
#pragma unroll
                    for (int c = 0; c < C; ++c) sv[n * C + c] = bilinear(g_lt, g_rb, g_lb, g_rt, c00[c], c11[c], c01[c], c10[c]);
                } else {               // offset beyond the halo: from the image (L2), as in l0_rows_kernel
                    const uint32_t ra = (uint32_t)q.r0 * (uint32_t)(W * C), rb = (uint32_t)q.r1 * (uint32_t)(W * C);
                    const uint32_t ka = (uint32_t)q.k0 * C, kb = (uint32_t)q.k1 * C;
                    const unsigned short* p00 = xs + (ra + ka);
                    const unsigned short* p11 = xs + (rb + kb);
                    const unsigned short* p01 = xs + (ra + kb);
                    const unsigned short* p10 = xs + (rb + ka);
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        sv[n * C + c] = bilinear(g_lt, g_rb, g_lb, g_rt, bf16_lo(p00[c]), bf16_lo(p11[c]), bf16_lo(p01[c]), bf16_lo(p10[c]));
                }
            }
            uint32_t pk[5];
#pragma unroll
            for (int q2 = 0; q2 < 5; ++q2) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[q2]) : "f"(sv[2 * q2 + 1]), "f"(sv[2 * q2]));
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a2_row + (0u ^ sw2)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a2_row + (16u ^ sw2)), "r"(pk[4]), "r"(0u), "r"(0u), "r"(0u) : "memory");
        }
        stamp(it);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        stamp(it);
        if (leader) {
            tc_fence_after_sync();
            mma_bf16_ss(tmem_base + 16, dA2, dW2, idesc, 0u);
            mma_commit(&bar[1]);
            // every thread is past its last read of the staged tile: fetch the next one under the MMA and the epilogue
            if (tile + (int)gridDim.x < num_tiles) issue_tile(nxt);
        }
        __syncwarp();
        mbar_wait(&bar[1], ph);
        stamp(it);
        tc_fence_after_sync();
        {
            uint32_t v[16];
            tmem_ld_32x32b_x16(lane_addr + 16, v);
            tmem_ld_wait();
            if (valid) {
                uint32_t pk[O / 2];
#pragma unroll
                for (int o4 = 0; o4 < O / 4; ++o4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_sc + o4 * 4);
                    const float4 sh = *reinterpret_cast<const float4*>(s_sc + O + o4 * 4);
                    float z[4];
                    z[0] = fmaf(__uint_as_float(v[o4 * 4 + 0]), sc.x, sh.x);
                    z[1] = fmaf(__uint_as_float(v[o4 * 4 + 1]), sc.y, sh.y);
                    z[2] = fmaf(__uint_as_float(v[o4 * 4 + 2]), sc.z, sh.z);
                    z[3] = fmaf(__uint_as_float(v[o4 * 4 + 3]), sc.w, sh.w);
                    if (act == LDCONV_ACT_SILU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {                  // z holds 0.5 (scale acc + shift): y = z tanh(z) + z
                            float t;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z[e]));
                            z[e] = fmaf(z[e], t, z[e]);
                        }
                    }
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[o4 * 2 + 0]) : "f"(z[1]), "f"(z[0]));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[o4 * 2 + 1]) : "f"(z[3]), "f"(z[2]));
                }
                uint4* dst = reinterpret_cast<uint4*>(out + m * O);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        cur = nxt;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 32);
}

static int launch_l0_tc(const __nv_bfloat16* x, const float* w_off, const float* b_off, const int* pn, const __nv_bfloat16* wt,
                        const float* scale, const float* shift, __nv_bfloat16* out, float* off_out, int B, int H, int W, int act,
                        cudaStream_t st)
{
    const int h = H / 2, w = W / 2, rows = B * h;
    const int tiles_x = (w + L0T_COLS - 1) / L0T_COLS;
    const long long nt = (long long)tiles_x * ((rows + 1) / 2);
    if (nt > 0x7fffffffll) return fail(LDCONV_E_ARG, "first-layer kernel: too many tiles");
    static int sms = 0, per_sm = 0;
    if (!sms) {
        int dev = 0;
        LDC_CUDA(cudaGetDevice(&dev));
        LDC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        // CTAs per SM from the kernel's own resource counts: cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for this kernel
        // (23 KB of static shared memory against the default carveout) although eight CTAs do run side by side -- measured: 861 us
        // with 148 CTAs, 489 / 293 / 213 us with 2 / 4 / 8 CTAs per SM (benchmarks/l0_ab.py)
        cudaFuncAttributes fa;
        LDC_CUDA(cudaFuncGetAttributes(&fa, l0_tc_kernel<false>));
        LDC_CUDA(cudaFuncSetAttribute(l0_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        LDC_CUDA(cudaFuncSetAttribute(l0_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        const int by_regs = 65536 / ((fa.numRegs > 0 ? fa.numRegs : 64) * L0T_THREADS);
        const int by_smem = (int)((220u * 1024u) / (fa.sharedSizeBytes + 1024u));
        per_sm = by_regs < by_smem ? by_regs : by_smem;
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 16) per_sm = 16;                      // 32 TMEM columns per CTA: 16 CTAs own the SM's 512
    }
    long long grid = (long long)sms * per_sm;
    if (grid > nt) grid = nt;
    CUtensorMap tm;
    {   // the image as rows of 4-byte words: (W * 3 / 2, H, B, 1)
        cuuint64_t gdim[4] = {(cuuint64_t)W * 3 / 2, (cuuint64_t)H, (cuuint64_t)B, 1};
        cuuint64_t gstr[3] = {(cuuint64_t)W * 6, (cuuint64_t)H * W * 6, (cuuint64_t)B * H * W * 6};
        cuuint32_t box[4] = {(cuuint32_t)L0T_TWORDS, (cuuint32_t)L0T_TROWS, 1, 1};
        if (int e = encode_map(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, x, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    }
    auto kern = g_l0_trace ? l0_tc_kernel<true> : l0_tc_kernel<false>;
    kern<<<(unsigned)grid, L0T_THREADS, 0, st>>>(tm, x, w_off, b_off, pn, wt, scale, shift, out, off_out, rows, H, W, h, w, act,
                                                        tiles_x, (int)nt, g_l0_trace);
    LDC_LAUNCH_CHECK("l0_tc_kernel");
    set_impl(LDCONV_IMPL_TCGEN05);
    return LDCONV_OK;
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
    constexpr int use_tiled = 0;      // the tiled variant measured slower (1448 vs 1124 us on layer 0); kept for C = 4 experiments
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

}  // namespace ldc

using namespace ldc;

// debug / A-B: 0 = first layer on the tensor cores (l0_tc_kernel), 1 = the CUDA-core rows kernel
LDC_API int ldconv_debug_l0_variant(int v) { ldc::g_l0_variant = v; return LDCONV_OK; }
// device buffer of 64 long long (zero-filled by the caller) that the next first-layer launches stamp, or NULL to stop
LDC_API int ldconv_debug_l0_trace(void* device_buf) { ldc::g_l0_trace = (long long*)device_buf; return LDCONV_OK; }

LDC_API int ldconv_fused_supported(int B, int C, int H, int W, int N, int s, int O, int dtype)
{
    if (B < 0 || C < 1 || H < 1 || W < 1 || N < 1 || s < 1 || O < 1) return 0;
    if (dtype != LDCONV_F32 && dtype != LDCONV_BF16) return 0;
    const int V = dtype == LDCONV_BF16 ? 8 : 4;
    // small-C layers only (the model's first row); wider layers take ldconv_onepass_fwd (offset conv on the tensor cores)
    return (C <= 4 && N <= 9 && O <= 32 && O % V == 0) ? 1 : 0;
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
    if (dtype == LDCONV_BF16 && l0_rows_applicable(x, B, C, H, W, N, s, O)) {
        // tensor-core kernel: row pairs inside one image (H % 4 == 0), image rows that are whole 16-byte units for the tensor map
        const bool tc_ok = H % 4 == 0 && W % 8 == 0 && ((uintptr_t)x & 15) == 0 && B <= 65535;
        if (g_l0_variant == 1 || !tc_ok)
            return launch_l0_rows((const __nv_bfloat16*)x, w_off, b_off, p_n, (const __nv_bfloat16*)wt, scale, shift,
                                  (__nv_bfloat16*)out, off_out, B, H, W, act, st);
        return launch_l0_tc((const __nv_bfloat16*)x, w_off, b_off, p_n, (const __nv_bfloat16*)wt, scale, shift,
                            (__nv_bfloat16*)out, off_out, B, H, W, act, st);
    }
    if (dtype == LDCONV_BF16)
        return dispatch_smallc<__nv_bfloat16>(x, w_off, b_off, p_n, wt, scale, shift, out, off_out, B, C, H, W, N, s, O, act, st);
    return dispatch_smallc<float>(x, w_off, b_off, p_n, wt, scale, shift, out, off_out, B, C, H, W, N, s, O, act, st);
}
