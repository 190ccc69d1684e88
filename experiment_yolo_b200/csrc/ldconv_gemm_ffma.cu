// ldconv_gemm_ffma.cu -- CUDA-core (FFMA) GEMMs of the LDConv (N,1) conv, fp32 accumulation.
//
// These are the fp32-precision path (the 1e-4 max-abs parity bound rules out plain TF32 at K >~ 100) and
// the general-shape path for bf16 (any M / K / O, any alignment).  bf16 shapes that fit the tensor-core kernel are
// routed to ldconv_umma.cu (tcgen05 / TMEM) by ldconv_gemm_fwd below.
//
//   forward / data gradient:  C(M,O) = A(M,K) . Wt(O,K)^T  (+ folded-BN affine, SiLU, BatchNorm batch sums)
//                             replaces nn.Conv2d(inc, outc, (N,1), (N,1)) + BatchNorm2d + SiLU,
//                             /root/reference/ultralytics/nn/modules/conv.py:355,408
//   weight gradient:          dWt(O,K) += G(M,O)^T . A(M,K)   (long reduction over M, split across CTAs)
#include "common.cuh"

namespace ldc {

int umma_gemm_supported(int M, int K, int O, int dtype, const void* a, const void* wt, const void* out, const void* pre);
int wgrad_umma_supported(int M, int K, int O, const void* g, const void* a);
int wgrad_umma(const void* g, const void* a, float* dW, int M, int K, int O, cudaStream_t st);
int umma_gemm_fwd(const void* a, const void* wt, const float* scale, const float* shift, void* out, void* pre,
                  double* stat_sum, double* stat_sqsum, int M, int K, int O, int act, cudaStream_t st);

constexpr int BM = 128, BN = 64, BK = 16, PAD = 4;

template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4], bool vec_ok, int valid);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4], bool vec_ok, int valid)
{
    if (vec_ok && valid == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        for (int e = 0; e < valid; ++e) p[e] = v[e];
    }
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4], bool vec_ok, int valid)
{
    if (vec_ok && valid == 4) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        *reinterpret_cast<uint2*>(p) = u;
    } else {
        for (int e = 0; e < valid; ++e) p[e] = __float2bfloat16_rn(v[e]);
    }
}

// 256 threads, CTA tile 128 x 64, thread tile 8 x 4, K step 16, operands staged k-major in shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_nt_kernel(const T* __restrict__ A, const T* __restrict__ Wt, const float* __restrict__ scale,
               const float* __restrict__ shift, T* __restrict__ out, T* __restrict__ pre, double* __restrict__ stat_sum,
               double* __restrict__ stat_sqsum, int M, int K, int O, int act, int vec_in, int vec_out)
{
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    const int a_row = tid / 2, a_kh = (tid % 2) * 8;   // A tile: 128 rows x 16 k, 8 elements per thread
    const int b_row = tid / 4, b_kq = (tid % 4) * 4;   // W tile:  64 rows x 16 k, 4 elements per thread
    for (int k0 = 0; k0 < K; k0 += BK) {
        {   // stage A
            const long long m = m0 + a_row;
            float v[8];
            if (m < M && vec_in && k0 + a_kh + 8 <= K) {
                if constexpr (sizeof(T) == 2) {
                    Vec16<T>::load(A + m * K + k0 + a_kh, v);
                } else {
                    float lo[4], hi[4];
                    Vec16<T>::load(A + m * K + k0 + a_kh, lo);
                    Vec16<T>::load(A + m * K + k0 + a_kh + 4, hi);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { v[e] = lo[e]; v[4 + e] = hi[e]; }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = k0 + a_kh + e;
                    v[e] = (m < M && k < K) ? Elem<T>::to_f(A[m * K + k]) : 0.f;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) As[a_kh + e][a_row] = v[e];
        }
        {   // stage W
            const int n = n0 + b_row;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + b_kq + e;
                Bs[b_kq + e][b_row] = (n < O && k < K) ? Elem<T>::to_f(Wt[(size_t)n * K + k]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a_lo = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a_hi = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }

    // ---- epilogue: raw accumulator, folded BN + activation, BatchNorm batch sums ------------------------------------
    const int nb = n0 + tx * 4;
    const int ncols = nb >= O ? 0 : (O - nb < 4 ? O - nb : 4);
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < ncols; ++c) {
        if (scale) sc[c] = scale[nb + c];
        if (shift) sh[c] = shift[nb + c];
    }
    float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long m = m0 + ty * 8 + r;
        if (m >= M || ncols == 0) continue;
        if (pre) store4<T>(pre + m * O + nb, acc[r], vec_out, ncols);
        if (stat_sum) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { csum[c] += acc[r][c]; csq[c] = fmaf(acc[r][c], acc[r][c], csq[c]); }
        }
        if (out) {
            float y[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float z = fmaf(acc[r][c], sc[c], sh[c]);
                y[c] = act == LDCONV_ACT_SILU ? silu(z) : z;
            }
            store4<T>(out + m * O + nb, y, vec_out, ncols);
        }
    }
    if (stat_sum) {
        // fold the 16 row groups of the CTA through shared memory (reuse As), then one fp64 atomic per column
        __syncthreads();
        float* red = &As[0][0];  // needs 2 * 16 * 64 floats = 2048 <= 16 * 132
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            red[ty * 64 + tx * 4 + c] = csum[c];
            red[1024 + ty * 64 + tx * 4 + c] = csq[c];
        }
        __syncthreads();
        if (tid < 64 && n0 + tid < O) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int g = 0; g < 16; ++g) { a += (double)red[g * 64 + tid]; b += (double)red[1024 + g * 64 + tid]; }
            atomicAdd(stat_sum + n0 + tid, a);
            atomicAdd(stat_sqsum + n0 + tid, b);
        }
    }
}

// dWt(O,K) += G(M,O)^T . A(M,K).  CTA tile 64 (o) x 64 (k), thread tile 4 x 4, 16 rows of M per step; blockIdx.z splits M.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const T* __restrict__ G, const T* __restrict__ A, float* __restrict__ dW, int M, int K, int O,
               int rows_per_cta, int vec_g, int vec_a)
{
    __shared__ __align__(16) float Gs[16][64];
    __shared__ __align__(16) float As[16][64];
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    const int o0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
    const long long m_begin = (long long)blockIdx.z * rows_per_cta;
    const long long m_end = m_begin + rows_per_cta < M ? m_begin + rows_per_cta : M;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const int lrow = tid / 16, lcol = (tid % 16) * 4;
    for (long long m0 = m_begin; m0 < m_end; m0 += 16) {
        const long long m = m0 + lrow;
        float gv[4], av[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = o0 + lcol + e, k = k0 + lcol + e;
            gv[e] = (m < m_end && o < O) ? Elem<T>::to_f(G[m * O + o]) : 0.f;
            av[e] = (m < m_end && k < K) ? Elem<T>::to_f(A[m * K + k]) : 0.f;
        }
        (void)vec_g; (void)vec_a;
        *reinterpret_cast<float4*>(&Gs[lrow][lcol]) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        *reinterpret_cast<float4*>(&As[lrow][lcol]) = make_float4(av[0], av[1], av[2], av[3]);
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < 16; ++mm) {
            const float4 g4 = *reinterpret_cast<const float4*>(&Gs[mm][ty * 4]);
            const float4 a4 = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(g[r], a[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int o = o0 + ty * 4 + r;
        if (o >= O) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = k0 + tx * 4 + c;
            if (k < K) atomicAdd(dW + (size_t)o * K + k, acc[r][c]);
        }
    }
}


// ---- thin shapes (the model's first layer: K = N*C = 9, O = 16, M = B*320*320) ---------------------------------------------------
// The tiled kernels above waste > 90 % of their 128 x 64 / 64 x 64 tiles on such shapes (benchmarks/profile_train.py: 1.7 ms
// forward, 3.1 ms weight gradient at batch 128 for ~0.1 ms of HBM traffic).  Thin kernels: one thread per row of A, the O x K
// weight in shared memory as fp32, accumulators in registers; rows are walked grid-stride so that the BatchNorm sums / the weight
// gradient are reduced per thread first, then per warp (shuffles), per CTA (shared memory) and once per CTA in global memory.
constexpr int kThinO = 16, kThinK = 16;

template <typename T>
__global__ void __launch_bounds__(256, 3)
thin_nt_kernel(const T* __restrict__ A, const T* __restrict__ Wt, const float* __restrict__ scale, const float* __restrict__ shift,
               T* __restrict__ out, T* __restrict__ pre, double* __restrict__ stat_sum, double* __restrict__ stat_sqsum, int M, int K,
               int O, int act)
{
    __shared__ __align__(16) float sW[kThinK][kThinO];       // [k][o], zero padded
    __shared__ float sAff[2][kThinO];
    __shared__ double sStat[2][kThinO];
    const int tid = threadIdx.x;
    for (int t = tid; t < kThinK * kThinO; t += blockDim.x) {
        const int k = t / kThinO, o = t % kThinO;
        sW[k][o] = (k < K && o < O) ? Elem<T>::to_f(Wt[(size_t)o * K + k]) : 0.f;
    }
    if (tid < kThinO) {
        sAff[0][tid] = (scale && tid < O) ? scale[tid] : 1.f;
        sAff[1][tid] = (shift && tid < O) ? shift[tid] : 0.f;
        sStat[0][tid] = 0.0; sStat[1][tid] = 0.0;
    }
    __syncthreads();
    float csum[kThinO], csq[kThinO];
#pragma unroll
    for (int o = 0; o < kThinO; ++o) csum[o] = csq[o] = 0.f;
    for (long long m = (long long)blockIdx.x * blockDim.x + tid; m < M; m += (long long)gridDim.x * blockDim.x) {
        const T* a = A + m * K;
        float acc[kThinO];
#pragma unroll
        for (int o = 0; o < kThinO; ++o) acc[o] = 0.f;
#pragma unroll
        for (int k = 0; k < kThinK; ++k) {
            if (k < K) {
                const float av = Elem<T>::to_f(a[k]);
                // volatile shared loads: the compiler otherwise hoists all 256 weights out of the row loop and spills them
                const uint32_t wrow = (uint32_t)__cvta_generic_to_shared(&sW[k][0]);
#pragma unroll
                for (int o4 = 0; o4 < kThinO / 4; ++o4) {
                    float4 wv;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(wv.x), "=f"(wv.y), "=f"(wv.z), "=f"(wv.w)
                                 : "r"(wrow + o4 * 16));
                    acc[o4 * 4 + 0] = fmaf(av, wv.x, acc[o4 * 4 + 0]);
                    acc[o4 * 4 + 1] = fmaf(av, wv.y, acc[o4 * 4 + 1]);
                    acc[o4 * 4 + 2] = fmaf(av, wv.z, acc[o4 * 4 + 2]);
                    acc[o4 * 4 + 3] = fmaf(av, wv.w, acc[o4 * 4 + 3]);
                }
            }
        }
        if (stat_sum) {
#pragma unroll
            for (int o = 0; o < kThinO; ++o) { csum[o] += acc[o]; csq[o] = fmaf(acc[o], acc[o], csq[o]); }
        }
        if (pre) {
            T* d = pre + m * O;
#pragma unroll
            for (int o = 0; o < kThinO; ++o)
                if (o < O) d[o] = Elem<T>::from_f(acc[o]);
        }
        if (out) {
            T* d = out + m * O;
#pragma unroll
            for (int o = 0; o < kThinO; ++o) {
                if (o < O) {
                    const float z = fmaf(acc[o], sAff[0][o], sAff[1][o]);
                    d[o] = Elem<T>::from_f(act == LDCONV_ACT_SILU ? silu(z) : z);
                }
            }
        }
    }
    if (stat_sum) {
#pragma unroll
        for (int o = 0; o < kThinO; ++o) {
            const float a = warp_sum(csum[o]), b = warp_sum(csq[o]);
            if ((tid & 31) == 0 && o < O) { atomicAdd(&sStat[0][o], (double)a); atomicAdd(&sStat[1][o], (double)b); }
        }
        __syncthreads();
        if (tid < O) { atomicAdd(stat_sum + tid, sStat[0][tid]); atomicAdd(stat_sqsum + tid, sStat[1][tid]); }
    }
}

// dWt(O,K) += G(M,O)^T . A(M,K) for O <= 16, K <= 12: two threads per row (each owns 8 of the output channels).
constexpr int kThinWK = 12;
template <typename T>
__global__ void __launch_bounds__(256)
thin_tn_kernel(const T* __restrict__ G, const T* __restrict__ A, float* __restrict__ dW, int M, int K, int O)
{
    __shared__ float sAcc[kThinO][kThinWK];
    const int tid = threadIdx.x, half = tid & 1;
    for (int t = tid; t < kThinO * kThinWK; t += blockDim.x) (&sAcc[0][0])[t] = 0.f;
    __syncthreads();
    float acc[8][kThinWK];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int k = 0; k < kThinWK; ++k) acc[o][k] = 0.f;
    const long long pairs = ((long long)gridDim.x * blockDim.x) >> 1;
    for (long long m = ((long long)blockIdx.x * blockDim.x + tid) >> 1; m < M; m += pairs) {
        float g[8], a[kThinWK];
#pragma unroll
        for (int o = 0; o < 8; ++o) g[o] = half * 8 + o < O ? Elem<T>::to_f(G[m * O + half * 8 + o]) : 0.f;
#pragma unroll
        for (int k = 0; k < kThinWK; ++k) a[k] = k < K ? Elem<T>::to_f(A[m * K + k]) : 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o)
#pragma unroll
            for (int k = 0; k < kThinWK; ++k) acc[o][k] = fmaf(g[o], a[k], acc[o][k]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int k = 0; k < kThinWK; ++k) {
            float v = acc[o][k];                      // lanes of equal parity own the same output channels
#pragma unroll
            for (int d = 16; d >= 2; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((tid & 31) < 2 && k < K && half * 8 + o < O) atomicAdd(&sAcc[half * 8 + o][k], v);
        }
    __syncthreads();
    for (int t = tid; t < O * K; t += blockDim.x) atomicAdd(dW + t, sAcc[t / K][t % K]);
}

static constexpr int thin_enabled() { return 1; }      // K <= 16, O <= 16: thread-per-row kernels (145.8 vs 149.8 ms per training step)

template <typename T>
static int gemm_nt_t(const T* a, const T* wt, const float* scale, const float* shift, T* out, T* pre, double* stat_sum,
                     double* stat_sqsum, int M, int K, int O, int act, cudaStream_t st)
{
    if (thin_enabled() && K <= kThinK && O <= kThinO && M >= 4096) {
        unsigned blocks = cdiv(M, 256 * 8);
        if (blocks > (unsigned)num_sms() * 8) blocks = (unsigned)num_sms() * 8;
        thin_nt_kernel<T><<<blocks, 256, 0, st>>>(a, wt, scale, shift, out, pre, stat_sum, stat_sqsum, M, K, O, act);
        LDC_LAUNCH_CHECK("thin_nt_kernel");
        set_impl(LDCONV_IMPL_FFMA);
        return LDCONV_OK;
    }
    const int vec_in = (K % 8 == 0) && aligned16(a);
    const int vec_out = (O % 4 == 0) && (!out || aligned16(out)) && (!pre || aligned16(pre));
    dim3 grid(cdiv(M, BM), cdiv(O, BN));
    gemm_nt_kernel<T><<<grid, 256, 0, st>>>(a, wt, scale, shift, out, pre, stat_sum, stat_sqsum, M, K, O, act, vec_in,
                                            vec_out);
    LDC_LAUNCH_CHECK("gemm_nt_kernel");
    set_impl(LDCONV_IMPL_FFMA);
    return LDCONV_OK;
}

template <typename T>
static int gemm_tn_t(const T* g, const T* a, float* dw, int M, int K, int O, cudaStream_t st)
{
    if (thin_enabled() && K <= kThinWK && O <= kThinO && M >= 4096) {
        unsigned blocks = cdiv(M, 128 * 16);
        if (blocks > (unsigned)num_sms() * 4) blocks = (unsigned)num_sms() * 4;
        thin_tn_kernel<T><<<blocks, 256, 0, st>>>(g, a, dw, M, K, O);
        LDC_LAUNCH_CHECK("thin_tn_kernel");
        return LDCONV_OK;
    }
    const int tiles = (int)(cdiv(K, 64) * cdiv(O, 64));
    long long splits = (long long)num_sms() * 4 / tiles;
    if (splits < 1) splits = 1;
    long long rows = (M + splits - 1) / splits;
    rows = (rows + 15) / 16 * 16;
    if (rows < 64) rows = 64;
    splits = (M + rows - 1) / rows;
    dim3 grid(cdiv(K, 64), cdiv(O, 64), (unsigned)splits);
    gemm_tn_kernel<T><<<grid, 256, 0, st>>>(g, a, dw, M, K, O, (int)rows, 0, 0);
    LDC_LAUNCH_CHECK("gemm_tn_kernel");
    return LDCONV_OK;
}

}  // namespace ldc

using namespace ldc;

LDC_API int ldconv_gemm_fwd(const void* a, const void* wt, const float* scale, const float* shift, void* out, void* pre,
                            double* stat_sum, double* stat_sqsum, int M, int K, int O, int act, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_F32 || dtype == LDCONV_BF16, "ldconv_gemm_fwd: unsupported dtype %d", dtype);
    LDC_REQUIRE(a && wt && (out || pre) && M >= 0 && K >= 1 && O >= 1, "ldconv_gemm_fwd: bad arguments");
    LDC_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), "ldconv_gemm_fwd: pass both stat buffers or neither");
    LDC_REQUIRE(act == LDCONV_ACT_NONE || act == LDCONV_ACT_SILU, "ldconv_gemm_fwd: unknown activation %d", act);
    if (M == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (umma_gemm_supported(M, K, O, dtype, a, wt, out, pre))
        return umma_gemm_fwd(a, wt, scale, shift, out, pre, stat_sum, stat_sqsum, M, K, O, act, st);
    if (dtype == LDCONV_F32)
        return gemm_nt_t<float>((const float*)a, (const float*)wt, scale, shift, (float*)out, (float*)pre, stat_sum,
                                stat_sqsum, M, K, O, act, st);
    return gemm_nt_t<__nv_bfloat16>((const __nv_bfloat16*)a, (const __nv_bfloat16*)wt, scale, shift, (__nv_bfloat16*)out,
                                    (__nv_bfloat16*)pre, stat_sum, stat_sqsum, M, K, O, act, st);
}

LDC_API int ldconv_gemm_bwd_weight(const void* grad_pre, const void* operand, float* grad_wt, int M, int K, int O,
                                   int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_F32 || dtype == LDCONV_BF16, "ldconv_gemm_bwd_weight: unsupported dtype %d", dtype);
    LDC_REQUIRE(grad_pre && operand && grad_wt && M >= 0 && K >= 1 && O >= 1, "ldconv_gemm_bwd_weight: bad arguments");
    if (M == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_BF16 && wgrad_umma_supported(M, K, O, grad_pre, operand)) {      // tensor-core reduction over M
        set_impl(LDCONV_IMPL_TCGEN05);
        return wgrad_umma(grad_pre, operand, grad_wt, M, K, O, st);
    }
    set_impl(LDCONV_IMPL_FFMA);
    if (dtype == LDCONV_F32) return gemm_tn_t<float>((const float*)grad_pre, (const float*)operand, grad_wt, M, K, O, st);
    return gemm_tn_t<__nv_bfloat16>((const __nv_bfloat16*)grad_pre, (const __nv_bfloat16*)operand, grad_wt, M, K, O, st);
}
