// ldconv_core.cu -- CUDA-core kernels of the LDConv hot path (sm_100a): offset conv, fused grid + bilinear gather,
// BatchNorm/SiLU passes, and the backward scatter.  All of them are HBM-bound byte movers, so the rules
// that matter are coalesced 128-bit NHWC accesses and enough CTAs to cover 148 SMs -- not tensor cores.
//
// Reference being replaced: /root/reference/ultralytics/nn/modules/conv.py:350-503 (class LDConv); each kernel cites
// its lines.  The C ABI is declared in include/ldconv_b200.h.
#include "common.cuh"

namespace ldc {

// =====================================================================================================================
// error plumbing
// =====================================================================================================================
static thread_local char g_err[768] = "";
static thread_local int g_impl = 0;
char* err_buf() { return g_err; }
void set_impl(int impl) { g_impl = impl; }
int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int pdl_enabled()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("LDCONV_PDL"); v = e ? atoi(e) : 1; }
    return v;
}

int num_sms()
{
    static thread_local int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

// =====================================================================================================================
// offset conv forward -- conv.py:356,368   offset = p_conv(x), 3x3 / pad 1 / stride s, C -> 2N, + bias
// One thread per output pixel keeps all 2N accumulators in registers; the weights of a channel chunk sit in shared
// memory as [tap][c][ON] (ON = 2N padded to a multiple of 4) and are read as broadcast float4; x is read as 16-byte
// NHWC vectors (neighbouring pixels of a warp overlap in L1).
// =====================================================================================================================
template <typename T, int ON, bool VECX>
__global__ void __launch_bounds__(128)
offset_conv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ off, int B, int C, int H, int W, int h, int wo, int N, int s, int cchunk)
{
    extern __shared__ __align__(16) float s_w[];  // [9][cchunk][ON]
    const int O2 = 2 * N;
    const long long M = (long long)B * h * wo;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = m < M;
    int b = 0, i = 0, j = 0;
    if (valid) {
        j = (int)(m % wo);
        i = (int)((m / wo) % h);
        b = (int)(m / ((long long)wo * h));
    }
    float acc[ON];
#pragma unroll
    for (int o = 0; o < ON; ++o) acc[o] = (bias != nullptr && o < O2) ? bias[o] : 0.f;

    constexpr int V = VECX ? Vec16<T>::N : 1;
    for (int c0 = 0; c0 < C; c0 += cchunk) {
        const int cc_n = min(cchunk, C - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < 9 * cc_n * ON; t += blockDim.x) {
            const int o = t % ON;
            const int cc = (t / ON) % cc_n;
            const int tap = t / (ON * cc_n);
            s_w[(tap * cchunk + cc) * ON + o] = (o < O2) ? w[((size_t)tap * C + c0 + cc) * O2 + o] : 0.f;
        }
        __syncthreads();
        if (!valid) continue;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int r = i * s + tap / 3 - 1;
            const int k = j * s + tap % 3 - 1;
            if (r < 0 || r >= H || k < 0 || k >= W) continue;
            const T* xp = x + (((size_t)b * H + r) * W + k) * C + c0;
            const float* wp = s_w + (size_t)tap * cchunk * ON;
            for (int cc = 0; cc < cc_n; cc += V) {
                float xv[V];
                if constexpr (VECX) {
                    Vec16<T>::load(xp + cc, xv);
                } else {
                    xv[0] = Elem<T>::to_f(xp[cc]);
                }
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float4* w4 = reinterpret_cast<const float4*>(wp + (cc + v) * ON);
#pragma unroll
                    for (int o4 = 0; o4 < ON / 4; ++o4) {
                        const float4 wv = w4[o4];
                        acc[o4 * 4 + 0] = fmaf(xv[v], wv.x, acc[o4 * 4 + 0]);
                        acc[o4 * 4 + 1] = fmaf(xv[v], wv.y, acc[o4 * 4 + 1]);
                        acc[o4 * 4 + 2] = fmaf(xv[v], wv.z, acc[o4 * 4 + 2]);
                        acc[o4 * 4 + 3] = fmaf(xv[v], wv.w, acc[o4 * 4 + 3]);
                    }
                }
            }
        }
    }
    if (valid) {
        float* op = off + (size_t)m * O2;
#pragma unroll
        for (int o = 0; o < ON; ++o)
            if (o < O2) op[o] = acc[o];
    }
}

template <typename T, int ON>
static int launch_offset_conv_fwd(const T* x, const float* w, const float* bias, float* off, int B, int C, int H, int W,
                                  int N, int s, cudaStream_t st)
{
    const int h = out_size(H, s), wo = out_size(W, s);
    const long long M = (long long)B * h * wo;
    constexpr int V = Vec16<T>::N;
    const bool vec = (C % V == 0) && aligned16(x);
    // channel chunk: keep the staged weights under ~40 KB
    int cchunk = (40 * 1024) / (9 * ON * 4);
    cchunk = cchunk / V * V;
    if (cchunk < V) cchunk = V;
    if (cchunk > C) cchunk = C;
    const size_t smem = (size_t)9 * cchunk * ON * sizeof(float);
    const int threads = 128;
    const unsigned blocks = cdiv(M, threads);
    if (vec) {
        auto kern = offset_conv_fwd_kernel<T, ON, true>;
        LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
        kern<<<blocks, threads, smem, st>>>(x, w, bias, off, B, C, H, W, h, wo, N, s, cchunk);
    } else {
        auto kern = offset_conv_fwd_kernel<T, ON, false>;
        LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
        kern<<<blocks, threads, smem, st>>>(x, w, bias, off, B, C, H, W, h, wo, N, s, cchunk);
    }
    LDC_LAUNCH_CHECK("offset_conv_fwd_kernel");
    return LDCONV_OK;
}

template <typename T>
static int dispatch_offset_conv_fwd(const T* x, const float* w, const float* bias, float* off, int B, int C, int H,
                                    int W, int N, int s, cudaStream_t st)
{
    const int on = (2 * N + 3) / 4 * 4;
    switch (on) {
        case 4: return launch_offset_conv_fwd<T, 4>(x, w, bias, off, B, C, H, W, N, s, st);
        case 8: return launch_offset_conv_fwd<T, 8>(x, w, bias, off, B, C, H, W, N, s, st);
        case 12: return launch_offset_conv_fwd<T, 12>(x, w, bias, off, B, C, H, W, N, s, st);
        case 16: return launch_offset_conv_fwd<T, 16>(x, w, bias, off, B, C, H, W, N, s, st);
        case 20: return launch_offset_conv_fwd<T, 20>(x, w, bias, off, B, C, H, W, N, s, st);
        case 24: return launch_offset_conv_fwd<T, 24>(x, w, bias, off, B, C, H, W, N, s, st);
        case 28: return launch_offset_conv_fwd<T, 28>(x, w, bias, off, B, C, H, W, N, s, st);
        case 32: return launch_offset_conv_fwd<T, 32>(x, w, bias, off, B, C, H, W, N, s, st);
        default: return fail(LDCONV_E_ARG, "offset conv: num_param %d not supported (1..16)", N);
    }
}

// =====================================================================================================================
// fused grid + bilinear gather forward -- conv.py:369-407, 413-503
// One thread per (output pixel m, sample n, 16-byte channel vector cv): consecutive threads write consecutive 16-byte
// chunks of operand row m (k = n*C + c), so stores are fully coalesced, and the threads of one sample read the C
// contiguous channels of each corner pixel (coalesced NHWC).  Neighbouring samples share corner pixels through L1/L2.
// =====================================================================================================================
template <typename T, bool VECX>
__global__ void __launch_bounds__(256)
gather_fwd_kernel(const T* __restrict__ x, const float* __restrict__ off, const int* __restrict__ pn,
                  T* __restrict__ operand, int* __restrict__ dbg_idx, float* __restrict__ dbg_coord, int C, int H,
                  int W, int h, int w, int N, int s, int CV, long long total)
{
    constexpr int V = VECX ? Vec16<T>::N : 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cv = (int)(t % CV);
    const long long sn = t / CV;  // sample index m*N + n
    const int n = (int)(sn % N);
    const long long m = sn / N;
    const int j = (int)(m % w);
    const int i = (int)((m / w) % h);
    const int b = (int)(m / ((long long)w * h));

    const float* op = off + (size_t)m * 2 * N;
    const SamplePoint q = make_point(i, j, s, pn[n], pn[N + n], op[n], op[N + n], H, W);
    if (cv == 0) {
        if (dbg_idx) {
            int4 v = make_int4(q.r0, q.r1, q.k0, q.k1);
            *reinterpret_cast<int4*>(dbg_idx + (size_t)sn * 4) = v;
        }
        if (dbg_coord) {
            dbg_coord[(size_t)sn * 2 + 0] = q.pcr;
            dbg_coord[(size_t)sn * 2 + 1] = q.pck;
        }
    }
    const float g_lt = __fmul_rn(q.ar0, q.ak0), g_rb = __fmul_rn(q.ar1, q.ak1);
    const float g_lb = __fmul_rn(q.ar0, q.ak1), g_rt = __fmul_rn(q.ar1, q.ak0);

    const T* xb = x + (size_t)b * H * W * C + (size_t)cv * V;
    const T* p00 = xb + ((size_t)q.r0 * W + q.k0) * C;
    const T* p11 = xb + ((size_t)q.r1 * W + q.k1) * C;
    const T* p01 = xb + ((size_t)q.r0 * W + q.k1) * C;
    const T* p10 = xb + ((size_t)q.r1 * W + q.k0) * C;
    T* dst = operand + (size_t)m * N * C + (size_t)n * C + (size_t)cv * V;
    if constexpr (VECX) {
        float x00[V], x11[V], x01[V], x10[V], r[V];
        Vec16<T>::load(p00, x00);
        Vec16<T>::load(p11, x11);
        Vec16<T>::load(p01, x01);
        Vec16<T>::load(p10, x10);
#pragma unroll
        for (int v = 0; v < V; ++v) r[v] = bilinear(g_lt, g_rb, g_lb, g_rt, x00[v], x11[v], x01[v], x10[v]);
        Vec16<T>::store(dst, r);
    } else {
        const float r = bilinear(g_lt, g_rb, g_lb, g_rt, Elem<T>::to_f(*p00), Elem<T>::to_f(*p11), Elem<T>::to_f(*p01),
                                 Elem<T>::to_f(*p10));
        *dst = Elem<T>::from_f(r);
    }
}

template <typename T>
static int launch_gather_fwd(const T* x, const float* off, const int* pn, T* operand, int* dbg_idx, float* dbg_coord,
                             int B, int C, int H, int W, int N, int s, cudaStream_t st)
{
    const int h = out_size(H, s), w = out_size(W, s);
    constexpr int V = Vec16<T>::N;
    const bool vec = (C % V == 0) && aligned16(x) && aligned16(operand);
    const int CV = vec ? C / V : C;
    const long long total = (long long)B * h * w * N * CV;
    if (total == 0) return LDCONV_OK;
    const unsigned blocks = cdiv(total, 256);
    if (vec)
        gather_fwd_kernel<T, true><<<blocks, 256, 0, st>>>(x, off, pn, operand, dbg_idx, dbg_coord, C, H, W, h, w, N, s,
                                                           CV, total);
    else
        gather_fwd_kernel<T, false><<<blocks, 256, 0, st>>>(x, off, pn, operand, dbg_idx, dbg_coord, C, H, W, h, w, N,
                                                            s, CV, total);
    LDC_LAUNCH_CHECK("gather_fwd_kernel");
    return LDCONV_OK;
}

// =====================================================================================================================
// BatchNorm bookkeeping + activation passes -- the nn.BatchNorm2d / nn.SiLU tail of conv.py:355
// =====================================================================================================================
__global__ void bn_finalize_kernel(const double* __restrict__ stat_sum, const double* __restrict__ stat_sqsum,
                                   long long count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, float eps, float momentum, int training,
                                   float* scale, float* shift, float* save_mean, float* save_invstd, int O)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    double mean, var;
    if (training) {
        const double cnt = (double)count;
        mean = stat_sum[o] / cnt;
        var = stat_sqsum[o] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        if (running_mean) running_mean[o] = (float)((1.0 - (double)momentum) * running_mean[o] + (double)momentum * mean);
        if (running_var) {
            const double unbiased = count > 1 ? var * cnt / (cnt - 1.0) : var;
            running_var[o] = (float)((1.0 - (double)momentum) * running_var[o] + (double)momentum * unbiased);
        }
    } else {
        mean = running_mean[o];
        var = running_var[o];
    }
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double sc = (gamma ? (double)gamma[o] : 1.0) * invstd;
    scale[o] = (float)sc;
    shift[o] = (float)((beta ? (double)beta[o] : 0.0) - mean * sc);
    if (save_mean) save_mean[o] = (float)mean;
    if (save_invstd) save_invstd[o] = (float)invstd;
}

// ---- column-invariant streaming kernels (the vector path of the four BatchNorm / activation passes) -----------------------------
// The generic kernels below redo `(t * V) % O` in 64 bits and reload scale / shift / mean / invstd / the pass-1 sums for every
// element (up to six global loads per value): they were instruction / LSU-bound at 1.5-2.7 TB/s (profiles/r2_bwd_launches_L1_before.csv)
// and, with the training-mode Conv blocks going through them too, 21 % of the config-4 step.  Here the CTA size and the grid stride
// are multiples of the column-vector count CVn = O / V, so a thread keeps ONE column vector for its whole grid-stride loop: the
// per-column constants live in registers, the loop body is loads -> arithmetic -> store, unrolled UN times with the loads issued
// first (UN independent 16-byte requests in flight per thread).
constexpr int kColUnroll = 4;

__device__ __forceinline__ float sigmoid_t(float z, float)            // fp32 tensors: the reference's own formula
{
    return 1.f / (1.f + __expf(-z));
}
__device__ __forceinline__ float sigmoid_t(float z, __nv_bfloat16)    // bf16 tensors: MUFU.EX2 + MUFU.RCP (2 ulp, far below 2^-9)
{
    return __fdividef(1.f, 1.f + __expf(-z));
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_apply_cols_kernel(const T* __restrict__ pre, const float* __restrict__ scale, const float* __restrict__ shift,
                         T* __restrict__ out, long long nvec, int CVn, int act)
{
    constexpr int V = Vec16<T>::N;
    const unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int o0 = (int)(t0 % (unsigned)CVn) * V;
    float sc[V], sh[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { sc[e] = scale[o0 + e]; sh[e] = shift[o0 + e]; }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = t0; t < nvec; t += kColUnroll * stride) {
        uint4 raw[kColUnroll];
#pragma unroll
        for (int u = 0; u < kColUnroll; ++u)
            if (t + u * stride < nvec) raw[u] = *reinterpret_cast<const uint4*>(pre + (t + u * stride) * V);
#pragma unroll
        for (int u = 0; u < kColUnroll; ++u) {
            if (t + u * stride >= nvec) break;
            float v[V];
            Vec16<T>::unpack(raw[u], v);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float z = fmaf(v[e], sc[e], sh[e]);
                v[e] = act == LDCONV_ACT_SILU ? z * sigmoid_t(z, T()) : z;
            }
            *reinterpret_cast<uint4*>(out + (t + u * stride) * V) = Vec16<T>::pack(v);
        }
    }
}

// d(pre) = scale * (dz - (sum dz + xhat * sum dz xhat) / M) with dz = g * act'(z): folded per column into
//   scale * dz + c0 + c1 * (pre - mean),   c1 = -scale * invstd * red1 / M,   c0 = -scale * red0 / M      (training)
template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_cols_kernel(const T* __restrict__ pre, const T* __restrict__ gout, const float* __restrict__ scale,
                             const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                             const double* __restrict__ red, T* __restrict__ gpre, long long nvec, long long M, int O, int CVn,
                             int act, int training)
{
    constexpr int V = Vec16<T>::N;
    const unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int o0 = (int)(t0 % (unsigned)CVn) * V;
    float sc[V], sh[V], mu[V], c0[V], c1[V];
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        sc[e] = scale[o0 + e];
        sh[e] = shift[o0 + e];
        mu[e] = mean[o0 + e];
        if (training) {
            c1[e] = -sc[e] * invstd[o0 + e] * (float)red[O + o0 + e] * invM;
            c0[e] = -sc[e] * (float)red[o0 + e] * invM;
        } else {
            c1[e] = 0.f;
            c0[e] = 0.f;
        }
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = t0; t < nvec; t += kColUnroll * stride) {
        uint4 rp[kColUnroll], rg[kColUnroll];
#pragma unroll
        for (int u = 0; u < kColUnroll; ++u)
            if (t + u * stride < nvec) {
                rp[u] = *reinterpret_cast<const uint4*>(pre + (t + u * stride) * V);
                rg[u] = *reinterpret_cast<const uint4*>(gout + (t + u * stride) * V);
            }
#pragma unroll
        for (int u = 0; u < kColUnroll; ++u) {
            if (t + u * stride >= nvec) break;
            float pv[V], g[V];
            Vec16<T>::unpack(rp[u], pv);
            Vec16<T>::unpack(rg[u], g);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                float dz = g[e];
                if (act == LDCONV_ACT_SILU) {
                    const float z = fmaf(pv[e], sc[e], sh[e]);
                    const float sg = sigmoid_t(z, T());
                    dz *= sg * (1.f + z * (1.f - sg));
                }
                pv[e] = fmaf(sc[e], dz, fmaf(c1[e], pv[e] - mu[e], c0[e]));
            }
            *reinterpret_cast<uint4*>(gpre + (t + u * stride) * V) = Vec16<T>::pack(pv);
        }
    }
}

// fold the per-thread partial sums of the two column reductions below: tree over the row phases in shared memory, then one fp64
// atomic per column and quantity (the first version let `tile` threads walk all phases serially in fp64: up to 1024 dependent
// additions at O = 32, the tail of every small launch)
template <int V>
__device__ __forceinline__ void fold_phases(float (&a)[V], float (&b)[V], float* s_a, float* s_b, int tile, int phases,
                                            double* dst_a, double* dst_b, bool live)
{
    const int L = threadIdx.x;
#pragma unroll
    for (int e = 0; e < V; ++e) { s_a[L * V + e] = a[e]; s_b[L * V + e] = b[e]; }
    __syncthreads();
    for (int half = phases >> 1; half >= 1; half >>= 1) {
        if (L < half * tile) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                s_a[L * V + e] += s_a[(L + half * tile) * V + e];
                s_b[L * V + e] += s_b[(L + half * tile) * V + e];
            }
        }
        __syncthreads();
    }
    if (L < tile && live) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            atomicAdd(dst_a + e, (double)s_a[L * V + e]);
            atomicAdd(dst_b + e, (double)s_b[L * V + e]);
        }
    }
}

template <typename T, bool VECX>
__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const T* __restrict__ pre, const float* __restrict__ scale, const float* __restrict__ shift,
                    T* __restrict__ out, long long nvec, int O, int act)
{
    constexpr int V = VECX ? Vec16<T>::N : 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nvec;
         t += (long long)gridDim.x * blockDim.x) {
        const int o0 = (int)((t * V) % O);
        if constexpr (VECX) {
            float v[V];
            Vec16<T>::load(pre + t * V, v);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                float z = fmaf(v[e], scale[o0 + e], shift[o0 + e]);
                v[e] = act == LDCONV_ACT_SILU ? silu(z) : z;
            }
            Vec16<T>::store(out + t * V, v);
        } else {
            float z = fmaf(Elem<T>::to_f(pre[t]), scale[o0], shift[o0]);
            out[t] = Elem<T>::from_f(act == LDCONV_ACT_SILU ? silu(z) : z);
        }
    }
}

// Column reductions over an (M,O) row-major matrix.  Thread L of a 256-thread CTA owns column vector cv = L % tile of
// the rows whose phase is L / tile (tile = power of two >= O/V, at most 256); it accumulates V columns in fp32 over a
// grid-strided set of rows, the CTA folds the row phases through shared memory and issues one fp64 atomicAdd per
// column.  PASS2 = false: BatchNorm backward sums (sum dz, sum dz*xhat).
template <typename T, bool VECX>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const T* __restrict__ pre, const T* __restrict__ gout, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, double* __restrict__ red, long long M, int O, int act,
                         int tile, int col0)
{
    constexpr int V = VECX ? Vec16<T>::N : 1;
    __shared__ float s_a[256 * V];
    __shared__ float s_b[256 * V];
    const int L = threadIdx.x;
    const int cv = L % tile;
    const int phase = L / tile;
    const int phases = 256 / tile;
    const int o0 = col0 + cv * V;
    float a[V], bsum[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { a[e] = 0.f; bsum[e] = 0.f; }
    if (o0 < O) {
        float sc[V], sh[V], mu[V], is[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            sc[e] = scale[o0 + e]; sh[e] = shift[o0 + e]; mu[e] = mean[o0 + e]; is[e] = invstd[o0 + e];
        }
        const long long rstride = (long long)gridDim.x * phases;
        if constexpr (VECX) {
            for (long long r = (long long)blockIdx.x * phases + phase; r < M; r += kColUnroll * rstride) {
                uint4 rp[kColUnroll], rg[kColUnroll];
#pragma unroll
                for (int u = 0; u < kColUnroll; ++u)
                    if (r + u * rstride < M) {
                        rp[u] = *reinterpret_cast<const uint4*>(pre + (r + u * rstride) * O + o0);
                        rg[u] = *reinterpret_cast<const uint4*>(gout + (r + u * rstride) * O + o0);
                    }
#pragma unroll
                for (int u = 0; u < kColUnroll; ++u) {
                    if (r + u * rstride >= M) break;
                    float p[V], g[V];
                    Vec16<T>::unpack(rp[u], p);
                    Vec16<T>::unpack(rg[u], g);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        float dz = g[e];
                        if (act == LDCONV_ACT_SILU) {
                            const float z = fmaf(p[e], sc[e], sh[e]);
                            const float sg = sigmoid_t(z, T());
                            dz *= sg * (1.f + z * (1.f - sg));
                        }
                        a[e] += dz;
                        bsum[e] = fmaf(dz, (p[e] - mu[e]) * is[e], bsum[e]);
                    }
                }
            }
        } else {
            for (long long r = (long long)blockIdx.x * phases + phase; r < M; r += rstride) {
                const float p = Elem<T>::to_f(pre[r * O + o0]), g = Elem<T>::to_f(gout[r * O + o0]);
                const float z = fmaf(p, sc[0], sh[0]);
                const float dz = act == LDCONV_ACT_SILU ? g * silu_grad(z) : g;
                a[0] += dz;
                bsum[0] = fmaf(dz, (p - mu[0]) * is[0], bsum[0]);
            }
        }
    }
    fold_phases<V>(a, bsum, s_a, s_b, tile, phases, red + o0, red + O + o0, o0 < O);
}

template <typename T, bool VECX>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const T* __restrict__ pre, const T* __restrict__ gout, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const double* __restrict__ red, T* __restrict__ gpre,
                        long long nvec, long long M, int O, int act, int training)
{
    constexpr int V = VECX ? Vec16<T>::N : 1;
    const float invM = 1.f / (float)M;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nvec;
         t += (long long)gridDim.x * blockDim.x) {
        const int o0 = (int)((t * V) % O);
        float p[V], g[V];
        if constexpr (VECX) {
            Vec16<T>::load(pre + t * V, p);
            Vec16<T>::load(gout + t * V, g);
        } else {
            p[0] = Elem<T>::to_f(pre[t]);
            g[0] = Elem<T>::to_f(gout[t]);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int o = o0 + e;
            const float z = fmaf(p[e], scale[o], shift[o]);
            float d = act == LDCONV_ACT_SILU ? g[e] * silu_grad(z) : g[e];
            if (training) {
                const float xh = (p[e] - mean[o]) * invstd[o];
                d = d - ((float)red[o] + xh * (float)red[O + o]) * invM;
            }
            p[e] = scale[o] * d;
        }
        if constexpr (VECX) {
            Vec16<T>::store(gpre + t * V, p);
        } else {
            gpre[t] = Elem<T>::from_f(p[0]);
        }
    }
}

// Column sums / sums of squares of an (M,O) matrix (BatchNorm batch statistics of a stored pre-activation): same thread
// mapping as bn_act_bwd_reduce_kernel.
template <typename T, bool VECX>
__global__ void __launch_bounds__(256)
col_stats_kernel(const T* __restrict__ pre, double* __restrict__ sum, double* __restrict__ sqsum, long long M, int O,
                 int tile, int col0)
{
    constexpr int V = VECX ? Vec16<T>::N : 1;
    __shared__ float s_a[256 * V];
    __shared__ float s_b[256 * V];
    const int L = threadIdx.x;
    const int cv = L % tile;
    const int phase = L / tile;
    const int phases = 256 / tile;
    const int o0 = col0 + cv * V;
    float a[V], b[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { a[e] = 0.f; b[e] = 0.f; }
    if (o0 < O) {
        const long long rstride = (long long)gridDim.x * phases;
        if constexpr (VECX) {
            for (long long r = (long long)blockIdx.x * phases + phase; r < M; r += kColUnroll * rstride) {
                uint4 rp[kColUnroll];
#pragma unroll
                for (int u = 0; u < kColUnroll; ++u)
                    if (r + u * rstride < M) rp[u] = *reinterpret_cast<const uint4*>(pre + (r + u * rstride) * O + o0);
#pragma unroll
                for (int u = 0; u < kColUnroll; ++u) {
                    if (r + u * rstride >= M) break;
                    float p[V];
                    Vec16<T>::unpack(rp[u], p);
#pragma unroll
                    for (int e = 0; e < V; ++e) { a[e] += p[e]; b[e] = fmaf(p[e], p[e], b[e]); }
                }
            }
        } else {
            for (long long r = (long long)blockIdx.x * phases + phase; r < M; r += rstride) {
                const float p = Elem<T>::to_f(pre[r * O + o0]);
                a[0] += p;
                b[0] = fmaf(p, p, b[0]);
            }
        }
    }
    fold_phases<V>(a, b, s_a, s_b, tile, phases, sum + o0, sqsum + o0, o0 < O);
}

static int pow2_at_least(int v);

int col_stats_bf16(const __nv_bfloat16* pre, long long M, int O, double* sum, double* sqsum, cudaStream_t st)
{
    using T = __nv_bfloat16;
    constexpr int V = Vec16<T>::N;
    const bool vec = (O % V == 0) && aligned16(pre);
    const int cvn = vec ? O / V : O;
    const int tile = pow2_at_least(cvn < 256 ? cvn : 256);
    const int phases = 256 / tile;
    long long want = (M + (long long)phases * 8 - 1) / ((long long)phases * 8);
    // two CTAs per SM: every CTA ends with one fp64 atomic per column and quantity on the SAME addresses, which the L2 serialises;
    // measured at layer 1, batch 64 (105 MB): 49.6 us with 8 CTAs per SM, 37.0 with 4, 33.3 with 2 (the unrolled loop keeps
    // four 16-byte loads in flight per thread, so two CTAs per SM still cover the memory latency)
    const long long cap = (long long)num_sms() * 2;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
    const int vper = vec ? V : 1;
    for (int col0 = 0; col0 < O; col0 += tile * vper) {
        if (vec)
            col_stats_kernel<T, true><<<blocks, 256, 0, st>>>(pre, sum, sqsum, M, O, tile, col0);
        else
            col_stats_kernel<T, false><<<blocks, 256, 0, st>>>(pre, sum, sqsum, M, O, tile, col0);
        LDC_LAUNCH_CHECK("col_stats_kernel");
    }
    return LDCONV_OK;
}

static int pow2_at_least(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// =====================================================================================================================
// backward scatter -- autograd of conv.py:386-405: grad_x via atomics, grad_off reduced over C
// Same (m, n, cv) thread mapping as the forward gather, so grad_operand is read as coalesced 16-byte vectors.  The
// lanes that share a sample first reduce their grad_off partials with warp shuffles and issue ONE atomic per sample
// and axis (warp-aggregated); grad_x uses 128-bit vector reductions (red.global.add.v4.f32) per corner and lane.
// =====================================================================================================================
__device__ __forceinline__ void red_add_vec(float* dst, const float* v, int n)
{
#if __CUDA_ARCH__ >= 900
    if (n == 4) {
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
        return;
    }
#endif
    for (int e = 0; e < n; ++e) atomicAdd(dst + e, v[e]);
}

// bf16 accumulator (ldconv_gather_bwd_acc16): eight channels per 16-byte reduction -- half the L2 reduction requests of the fp32
// accumulator, which is what bounds the scatter (1.3-1.7 cycles per lane request, profiles/r1_ncu_scatterL1.txt).  The product
// g * weight is formed in fp32 and rounded to bf16 once; the accumulator rounds after every addition (what autograd's own
// scatter_add_ does for a reduced-precision model), so its tolerance is stated separately (tests/test_gpu_parity.py).
__device__ __forceinline__ void red_add_vec(__nv_bfloat16* dst, const float* v, int n)
{
    if (n == 8) {
        uint32_t p0, p1, p2, p3;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(v[1]), "f"(v[0]));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(v[3]), "f"(v[2]));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(v[5]), "f"(v[4]));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(v[7]), "f"(v[6]));
        asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
        return;
    }
    for (int e = 0; e < n; ++e) atomicAdd(dst + e, __float2bfloat16_rn(v[e]));
}

// IDX: the integer type of the item decomposition -- unsigned 32-bit whenever the item count fits (always at the benchmark
// sizes); the 64-bit divisions of the general case cost several hundred instructions per item
template <typename T, bool VECX, typename IDX, typename ACC = float>
__global__ void __launch_bounds__(256, 5)      // <= 51 registers: 40 warps per SM instead of 24
gather_bwd_kernel(const T* __restrict__ gop, const T* __restrict__ x, const float* __restrict__ off,
                  const int* __restrict__ pn, ACC* __restrict__ grad_x, float* __restrict__ grad_off, int C, int H,
                  int W, int h, int w, int N, int s, int CV, int group, long long total)
{
    constexpr int V = VECX ? Vec16<T>::N : 1;
    const IDX t = (IDX)blockIdx.x * (IDX)blockDim.x + (IDX)threadIdx.x;
    const bool valid = (long long)t < total;
    float acc_r = 0.f, acc_k = 0.f;
    IDX sn = 0, m = 0;
    int n = 0;
    bool in_r = false, in_k = false;
    if (valid) {
        const int cv = (int)(t % (IDX)CV);
        sn = t / (IDX)CV;
        n = (int)(sn % (IDX)N);
        m = sn / (IDX)N;
        const int j = (int)(m % (IDX)w);
        const IDX mi = m / (IDX)w;
        const int i = (int)(mi % (IDX)h);
        const int b = (int)(mi / (IDX)h);
        const float* op = off + (size_t)m * 2 * N;
        const SamplePoint q = make_point(i, j, s, pn[n], pn[N + n], op[n], op[N + n], H, W);
        in_r = q.in_r;
        in_k = q.in_k;
        const float g_lt = q.ar0 * q.ak0, g_rb = q.ar1 * q.ak1, g_lb = q.ar0 * q.ak1, g_rt = q.ar1 * q.ak0;
        const size_t cofs = (size_t)cv * V;
        const size_t base = (size_t)b * H * W * C;
        const size_t o00 = base + ((size_t)q.r0 * W + q.k0) * C + cofs;
        const size_t o11 = base + ((size_t)q.r1 * W + q.k1) * C + cofs;
        const size_t o01 = base + ((size_t)q.r0 * W + q.k1) * C + cofs;
        const size_t o10 = base + ((size_t)q.r1 * W + q.k0) * C + cofs;
        // Registers decide this kernel: it is load-latency-bound (profiles/r1_ncu_scatterL1.txt: 7.7 cycles of long scoreboard per
        // issued instruction at 24 warps per SM), so the four corner vectors of x live only while the offset gradient is formed,
        // and coincident corners are merged by adding their WEIGHTS -- one 8-value vector per emitted corner instead of four.
        float g[V];
        const T* gp = gop + (size_t)m * N * C + (size_t)n * C + cofs;
        {
            float x00[V], x11[V], x01[V], x10[V];
            if constexpr (VECX) {
                Vec16<T>::load(gp, g);
                Vec16<T>::load(x + o00, x00);
                Vec16<T>::load(x + o11, x11);
                Vec16<T>::load(x + o01, x01);
                Vec16<T>::load(x + o10, x10);
            } else {
                g[0] = Elem<T>::to_f(*gp);
                x00[0] = Elem<T>::to_f(x[o00]); x11[0] = Elem<T>::to_f(x[o11]);
                x01[0] = Elem<T>::to_f(x[o01]); x10[0] = Elem<T>::to_f(x[o10]);
            }
#pragma unroll
            for (int e = 0; e < V; ++e) {
                acc_r += g[e] * (-q.ak0 * x00[e] + q.ak1 * x11[e] - q.ak1 * x01[e] + q.ak0 * x10[e]);
                acc_k += g[e] * (-q.ar0 * x00[e] + q.ar1 * x11[e] + q.ar0 * x01[e] - q.ar1 * x10[e]);
            }
        }
        // scatter_add_ of the four GatherBackward nodes; V floats per corner as 128-bit vector reductions.
        // Corners that coincide (a clamped axis: r0 == r1 and / or k0 == k1, the reference's border doubling) are merged in
        // the thread first; lanes of the warp that still hit the SAME address (many samples clamped onto one border pixel --
        // the regime a diverging offset conv drives the layer into) are aggregated with match.any + shuffles so that one lane
        // issues the reduction for all of them.
        if (grad_x != nullptr) {
            const bool same_r = q.r0 == q.r1, same_k = q.k0 == q.k1;
            auto emit = [&](size_t o, float (&v)[V], bool clamped) {
                if (clamped) {                                  // lanes here: the warp's clamped corners only
                    const unsigned act = __activemask();
                    const unsigned peers = __match_any_sync(act, (unsigned long long)o);
                    if (__any_sync(act, (peers & (peers - 1)) != 0)) {          // some address is shared: aggregate
                        const int lane = threadIdx.x & 31;
                        const int leader = __ffs(peers) - 1;
                        float acc[V];
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[e] = 0.f;
                        for (unsigned m = act; m; m &= m - 1) {
                            const int src = __ffs(m) - 1;
                            const bool take = (peers >> src) & 1u;
#pragma unroll
                            for (int e = 0; e < V; ++e) {
                                const float t = __shfl_sync(act, v[e], src);
                                if (take) acc[e] += t;
                            }
                        }
                        if (lane != leader) return;
#pragma unroll
                        for (int e = 0; e < V; ++e) v[e] = acc[e];
                    }
                }
                constexpr int PER = sizeof(ACC) == 2 ? 8 : 4;     // values per 16-byte reduction
#pragma unroll
                for (int e0 = 0; e0 < V; e0 += PER) red_add_vec(grad_x + o + e0, v + e0, V - e0 < PER ? V - e0 : PER);
            };
            auto corner = [&](size_t o, float wgt, bool clamped) {
                float v[V];
#pragma unroll
                for (int e = 0; e < V; ++e) v[e] = g[e] * wgt;
                emit(o, v, clamped);
            };
            if (same_r && same_k) {
                corner(o00, (g_lt + g_rb) + (g_lb + g_rt), true);
            } else if (same_r) {           // (r0,k0) == (r1,k0) and (r0,k1) == (r1,k1)
                corner(o00, g_lt + g_rt, true);
                corner(o01, g_lb + g_rb, true);
            } else if (same_k) {           // (r0,k0) == (r0,k1) and (r1,k0) == (r1,k1)
                corner(o00, g_lt + g_lb, true);
                corner(o10, g_rt + g_rb, true);
            } else {
                corner(o00, g_lt, false);
                corner(o11, g_rb, false);
                corner(o01, g_lb, false);
                corner(o10, g_rt, false);
            }
        }
    }
    // grad_off: reduce over the channel lanes of the sample.  group = CV when CV is a power of two <= 32 (the lanes of
    // a sample are then an aligned sub-warp segment), else 1 (every lane adds its partial on its own).
    for (int o = group >> 1; o > 0; o >>= 1) {
        acc_r += __shfl_xor_sync(0xffffffffu, acc_r, o);
        acc_k += __shfl_xor_sync(0xffffffffu, acc_k, o);
    }
    if (valid && (t & (IDX)(group - 1)) == 0) {      // group is 1 or a power of two
        float* gp = grad_off + (size_t)m * 2 * N;
        if (in_r) atomicAdd(gp + n, acc_r);
        if (in_k) atomicAdd(gp + N + n, acc_k);
    }
}

// =====================================================================================================================
// offset conv backward -- conv2d backward of conv.py:356
//   (a) grad_x += conv_transpose(grad_off, w): gather form, one thread per (input pixel, 4 channels), plain RMW
//       (runs after the scatter kernel on the same stream; each element has exactly one owner).
//   (b) grad_w[tap][c][o] += sum_m grad_off[m,o] * x[tap-shifted m, c]; grad_b[o] += sum_m grad_off[m,o]
//       CTA (chunk of pixels, 16-channel slab): thread = (tap, channel), ON accumulators, grad_off rows broadcast via smem.
// =====================================================================================================================
template <int ON>
__global__ void __launch_bounds__(256)
offset_conv_bwd_data_kernel(const float* __restrict__ goff, const float* __restrict__ w, float* __restrict__ grad_x,
                            int B, int C, int H, int W, int h, int wo, int N, int s, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int O2 = 2 * N;
    const int cgn = (C + 3) / 4;
    const int cg = (int)(t % cgn);
    const long long pix = t / cgn;
    const int k = (int)(pix % W);
    const int r = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const int c0 = cg * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky) {
        const int ri = r + 1 - ky;
        if (ri < 0 || ri % s != 0) continue;
        const int i = ri / s;
        if (i >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int kj = k + 1 - kx;
            if (kj < 0 || kj % s != 0) continue;
            const int j = kj / s;
            if (j >= wo) continue;
            const float* gp = goff + (((size_t)b * h + i) * wo + j) * O2;
            const float* wp = w + ((size_t)(ky * 3 + kx) * C + c0) * O2;
            for (int o = 0; o < O2; ++o) {
                const float g = gp[o];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c0 + e < C) acc[e] = fmaf(g, wp[(size_t)e * O2 + o], acc[e]);
            }
        }
    }
    float* dst = grad_x + (size_t)pix * C + c0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (c0 + e < C) dst[e] += acc[e];
}

template <typename T, int ON>
__global__ void __launch_bounds__(160)
offset_conv_bwd_weight_kernel(const float* __restrict__ goff, const T* __restrict__ x, float* __restrict__ grad_w,
                              float* __restrict__ grad_b, int B, int C, int H, int W, int h, int wo, int N, int s,
                              long long M, int rows_per_cta)
{
    constexpr int PIX = 8;  // pixels staged per step
    __shared__ __align__(16) float s_g[PIX][ON];
    const int O2 = 2 * N;
    const int c_slab = blockIdx.y * 16;
    const int tap = threadIdx.x / 16;           // 0..9 (tap 9 = the bias lane group)
    const int c = c_slab + (threadIdx.x % 16);
    const bool is_w = tap < 9 && c < C;
    const bool is_b = tap == 9 && blockIdx.y == 0 && (threadIdx.x % 16) == 0 && grad_b != nullptr;
    float acc[ON];
#pragma unroll
    for (int o = 0; o < ON; ++o) acc[o] = 0.f;
    const long long m_begin = (long long)blockIdx.x * rows_per_cta;
    const long long m_end = min(M, m_begin + rows_per_cta);
    for (long long m0 = m_begin; m0 < m_end; m0 += PIX) {
        __syncthreads();
        for (int t = threadIdx.x; t < PIX * ON; t += blockDim.x) {
            const int p = t / ON, o = t % ON;
            s_g[p][o] = (m0 + p < m_end && o < O2) ? goff[(size_t)(m0 + p) * O2 + o] : 0.f;
        }
        __syncthreads();
        if (!(is_w || is_b)) continue;
        for (int p = 0; p < PIX; ++p) {
            const long long m = m0 + p;
            if (m >= m_end) break;
            float xv = 1.f;
            if (is_w) {
                const int j = (int)(m % wo);
                const int i = (int)((m / wo) % h);
                const int b = (int)(m / ((long long)wo * h));
                const int r = i * s + tap / 3 - 1;
                const int k = j * s + tap % 3 - 1;
                if (r < 0 || r >= H || k < 0 || k >= W) continue;
                xv = Elem<T>::to_f(x[(((size_t)b * H + r) * W + k) * C + c]);
            }
#pragma unroll
            for (int o = 0; o < ON; ++o) acc[o] = fmaf(xv, s_g[p][o], acc[o]);
        }
    }
    if (is_w) {
        float* dst = grad_w + ((size_t)tap * C + c) * O2;
#pragma unroll
        for (int o = 0; o < ON; ++o)
            if (o < O2) atomicAdd(dst + o, acc[o]);
    } else if (is_b) {
#pragma unroll
        for (int o = 0; o < ON; ++o)
            if (o < O2) atomicAdd(grad_b + o, acc[o]);
    }
}

int offconv_bwd_data_fast(const float* goff, const float* w, float* grad_x, int B, int C, int H, int W, int N, int s,
                          cudaStream_t st);

template <typename T, int ON>
static int launch_offset_conv_bwd(const float* goff, const T* x, const float* w, float* grad_x, float* grad_w,
                                  float* grad_b, int B, int C, int H, int W, int N, int s, cudaStream_t st)
{
    const int h = out_size(H, s), wo = out_size(W, s);
    const long long M = (long long)B * h * wo;
    if (grad_x) {
        if (offconv_bwd_data_fast(goff, w, grad_x, B, C, H, W, N, s, st)) {      // weights in smem, 16-byte RMW (ldconv_offconv_bwd.cu)
            LDC_LAUNCH_CHECK("offconv_bwd_data_kernel");
        } else {
            const long long total = (long long)B * H * W * ((C + 3) / 4);
            offset_conv_bwd_data_kernel<ON><<<cdiv(total, 256), 256, 0, st>>>(goff, w, grad_x, B, C, H, W, h, wo, N, s, total);
            LDC_LAUNCH_CHECK("offset_conv_bwd_data_kernel");
        }
    }
    if (grad_w || grad_b) {
        const int slabs = (C + 15) / 16;
        long long ctas_x = (long long)num_sms() * 8 / slabs;
        if (ctas_x < 1) ctas_x = 1;
        long long rows = (M + ctas_x - 1) / ctas_x;
        rows = (rows + 7) / 8 * 8;
        if (rows < 8) rows = 8;
        ctas_x = (M + rows - 1) / rows;
        dim3 grid((unsigned)ctas_x, (unsigned)slabs);
        offset_conv_bwd_weight_kernel<T, ON><<<grid, 160, 0, st>>>(goff, x, grad_w, grad_b, B, C, H, W, h, wo, N, s, M,
                                                                   (int)rows);
        LDC_LAUNCH_CHECK("offset_conv_bwd_weight_kernel");
    }
    return LDCONV_OK;
}

template <typename T>
static int dispatch_offset_conv_bwd(const float* goff, const T* x, const float* w, float* grad_x, float* grad_w,
                                    float* grad_b, int B, int C, int H, int W, int N, int s, cudaStream_t st)
{
    const int on = (2 * N + 3) / 4 * 4;
    switch (on) {
        case 4: return launch_offset_conv_bwd<T, 4>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 8: return launch_offset_conv_bwd<T, 8>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 12: return launch_offset_conv_bwd<T, 12>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 16: return launch_offset_conv_bwd<T, 16>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 20: return launch_offset_conv_bwd<T, 20>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 24: return launch_offset_conv_bwd<T, 24>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 28: return launch_offset_conv_bwd<T, 28>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        case 32: return launch_offset_conv_bwd<T, 32>(goff, x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
        default: return fail(LDCONV_E_ARG, "offset conv bwd: num_param %d not supported (1..16)", N);
    }
}

// =====================================================================================================================
// Detect head decode (reference nn/modules/head.py:55-77 + DFL nn/modules/block.py:37-56 + dist2bbox utils/tal.py:309-319):
// per anchor, softmax-expectation over the reg_max bins of each of the 4 sides, distances -> xywh box around the
// cell-centre anchor, times the level stride; class logits -> sigmoid.  One thread per anchor, fp32 math.
// =====================================================================================================================
template <int REG>
__global__ void __launch_bounds__(128)
detect_decode_kernel(const __nv_bfloat16* __restrict__ box, const __nv_bfloat16* __restrict__ cls,
                     __nv_bfloat16* __restrict__ y, int B, int H, int W, int nc, float stride, int a0, int total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_img = (long long)H * W;
    if (t >= (long long)B * per_img) return;
    const int b = (int)(t / per_img);
    const int a = (int)(t % per_img);
    const float ax = (float)(a % W) + 0.5f, ay = (float)(a / W) + 0.5f;
    const __nv_bfloat16* bp = box + t * (4 * REG);
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        float v[REG];
#pragma unroll
        for (int k0 = 0; k0 < REG; k0 += 8) {
            float tmp[8];
            Vec16<__nv_bfloat16>::load(bp + side * REG + k0, tmp);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[k0 + e] = tmp[e];
        }
        float mx = v[0];
#pragma unroll
        for (int k = 1; k < REG; ++k) mx = fmaxf(mx, v[k]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int k = 0; k < REG; ++k) {
            const float e = __expf(v[k] - mx);
            den += e;
            num = fmaf(e, (float)k, num);
        }
        d[side] = num / den;
    }
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    __nv_bfloat16* yp = y + (size_t)b * (4 + nc) * total + a0 + a;
    yp[0] = __float2bfloat16_rn(0.5f * (x1 + x2) * stride);
    yp[(size_t)total] = __float2bfloat16_rn(0.5f * (y1 + y2) * stride);
    yp[2 * (size_t)total] = __float2bfloat16_rn((x2 - x1) * stride);
    yp[3 * (size_t)total] = __float2bfloat16_rn((y2 - y1) * stride);
    const __nv_bfloat16* cp = cls + t * nc;
    for (int c = 0; c < nc; ++c) {
        const float z = __bfloat162float(cp[c]);
        yp[(size_t)(4 + c) * total] = __float2bfloat16_rn(1.f / (1.f + __expf(-z)));
    }
}

// Same decode with the box logits staged through shared memory: a warp copies the 32 x 128 bytes of its anchors with coalesced
// 16-byte loads (8 anchors = 8 lines per instruction; a thread reading its own 128-byte row touches 32 lines per instruction) into
// rows of 144 bytes (conflict-free for both the warp's stores and the per-thread reads), then every thread decodes its anchor.
template <int REG>
__global__ void __launch_bounds__(128)
detect_decode_staged_kernel(const __nv_bfloat16* __restrict__ box, const __nv_bfloat16* __restrict__ cls,
                            __nv_bfloat16* __restrict__ y, int B, int H, int W, int nc, float stride, int a0, int total)
{
    static_assert(REG == 16, "rows of 4 x 16 bins = 128 bytes");
    __shared__ uint4 s_row[128][9];
    const long long per_img = (long long)H * W;
    const long long n = (long long)B * per_img;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const long long w0 = (long long)blockIdx.x * blockDim.x + wrp * 32;      // first anchor of this warp
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = k * 32 + lane, al = idx >> 3, ch = idx & 7;
        if (w0 + al < n) s_row[wrp * 32 + al][ch] = __ldg(reinterpret_cast<const uint4*>(box + (w0 + al) * (4 * REG)) + ch);
    }
    __syncwarp();
    const long long t = w0 + lane;
    if (t >= n) return;
    const int b = (int)(t / per_img);
    const int a = (int)(t - (long long)b * per_img);
    const int ai = a / W, aj = a - ai * W;
    const float ax = (float)aj + 0.5f, ay = (float)ai + 0.5f;
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        const uint4 q0 = s_row[threadIdx.x][2 * side], q1 = s_row[threadIdx.x][2 * side + 1];
        const uint32_t wv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        float v[REG];
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[2 * e] = __uint_as_float(wv[e] << 16); v[2 * e + 1] = __uint_as_float(wv[e] & 0xffff0000u); }
        float mx = v[0];
#pragma unroll
        for (int k = 1; k < REG; ++k) mx = fmaxf(mx, v[k]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int k = 0; k < REG; ++k) {
            const float e = __expf(v[k] - mx);
            den += e;
            num = fmaf(e, (float)k, num);
        }
        d[side] = num / den;
    }
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    __nv_bfloat16* yp = y + (size_t)b * (4 + nc) * total + a0 + a;
    yp[0] = __float2bfloat16_rn(0.5f * (x1 + x2) * stride);
    yp[(size_t)total] = __float2bfloat16_rn(0.5f * (y1 + y2) * stride);
    yp[2 * (size_t)total] = __float2bfloat16_rn((x2 - x1) * stride);
    yp[3 * (size_t)total] = __float2bfloat16_rn((y2 - y1) * stride);
    const __nv_bfloat16* cp = cls + t * nc;
    for (int c = 0; c < nc; ++c) {
        const float z = __bfloat162float(cp[c]);
        yp[(size_t)(4 + c) * total] = __float2bfloat16_rn(1.f / (1.f + __expf(-z)));
    }
}

static int check_dims(const char* fn, int B, int C, int H, int W, int N, int s)
{
    if (B < 0 || C < 1 || H < 1 || W < 1 || N < 1 || N > 16 || s < 1)
        return fail(LDCONV_E_ARG, "%s: bad dims B=%d C=%d H=%d W=%d num_param=%d stride=%d", fn, B, C, H, W, N, s);
    if ((long long)H * W * C >= (1ll << 31))
        return fail(LDCONV_E_ARG, "%s: one image exceeds 2^31 elements", fn);
    return LDCONV_OK;
}

static int check_dtype(const char* fn, int dtype)
{
    if (dtype != LDCONV_F32 && dtype != LDCONV_BF16) return fail(LDCONV_E_ARG, "%s: unsupported dtype %d", fn, dtype);
    return LDCONV_OK;
}

}  // namespace ldc

using namespace ldc;

// =====================================================================================================================
// C ABI
// =====================================================================================================================
LDC_API int ldconv_version(void) { return LDCONV_ABI_VERSION; }
LDC_API const char* ldconv_last_error(void) { return ldc::err_buf(); }
LDC_API int ldconv_last_impl(void) { return ldc::g_impl; }

namespace ldc {
int umma_set_force_ffma(int v);
int gather_set_direct(int v);
int gather_set_miss_counter(void* p);
int gather_fwd_tiled(const void* x, const float* off, const int* pn, void* operand, int* dbg_idx, float* dbg_coord, int B,
                     int C, int H, int W, int N, int s, int dtype, cudaStream_t st);
int scatter_bwd_tiled(const __nv_bfloat16* x, const float* off, const int* pn, const __nv_bfloat16* gop, __nv_bfloat16* grad_x,
                      float* grad_off, int B, int C, int H, int W, int N, int s, cudaStream_t st);
}

LDC_API int ldconv_set_flag(int flag, int value)
{
    if (flag == LDCONV_FLAG_FORCE_FFMA) return ldc::umma_set_force_ffma(value);
    if (flag == LDCONV_FLAG_GATHER_DIRECT) return ldc::gather_set_direct(value);
    return fail(LDCONV_E_ARG, "ldconv_set_flag: unknown flag %d", flag);
}

LDC_API int ldconv_set_gather_miss_counter(void* device_u64) { return ldc::gather_set_miss_counter(device_u64); }

LDC_API int ldconv_device_check(void)
{
    int dev = 0, major = 0;
    LDC_CUDA(cudaGetDevice(&dev));
    LDC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        return fail(LDCONV_E_DEVICE, "libldconv_b200 is built for sm_100a only; device %d has compute capability %d.x", dev,
                    major);
    return LDCONV_OK;
}

LDC_API int ldconv_p_n(int N, int32_t* out)
{
    LDC_REQUIRE(N >= 1 && out != nullptr, "ldconv_p_n: bad arguments");
    // conv.py:413-432: base = round(sqrt(N)); N // base full rows, then a partial row of N % base
    int base = 1;
    while ((base + 1) * (base + 1) - (base + 1) < N) ++base;  // round(sqrt(N)) == base  <=>  base^2-base < N <= base^2+base
    const int rows = N / base, mod = N % base;
    int idx = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < base; ++c, ++idx) { out[idx] = r; out[N + idx] = c; }
    for (int c = 0; c < mod; ++c, ++idx) { out[idx] = rows; out[N + idx] = c; }
    return LDCONV_OK;
}

LDC_API int ldconv_offset_conv_fwd(const void* x, const float* w, const float* bias, float* off, int B, int C, int H,
                                   int W, int N, int s, int dtype, void* stream)
{
    if (int e = check_dims("ldconv_offset_conv_fwd", B, C, H, W, N, s)) return e;
    if (int e = check_dtype("ldconv_offset_conv_fwd", dtype)) return e;
    LDC_REQUIRE(x && w && off, "ldconv_offset_conv_fwd: null pointer");
    if (B == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_F32) return dispatch_offset_conv_fwd<float>((const float*)x, w, bias, off, B, C, H, W, N, s, st);
    return dispatch_offset_conv_fwd<__nv_bfloat16>((const __nv_bfloat16*)x, w, bias, off, B, C, H, W, N, s, st);
}

LDC_API int ldconv_gather_fwd(const void* x, const float* off, const int32_t* p_n, void* operand, int32_t* dbg_idx,
                              float* dbg_coord, int B, int C, int H, int W, int N, int s, int dtype, void* stream)
{
    if (int e = check_dims("ldconv_gather_fwd", B, C, H, W, N, s)) return e;
    if (int e = check_dtype("ldconv_gather_fwd", dtype)) return e;
    LDC_REQUIRE(x && off && p_n && operand, "ldconv_gather_fwd: null pointer");
    if (dbg_idx && !aligned16(dbg_idx)) return fail(LDCONV_E_ALIGN, "ldconv_gather_fwd: dbg_idx must be 16-byte aligned");
    if (B == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {   // TMA-staged tile kernel when the shape is eligible (C*elem % 16 == 0), else the direct-load kernel below
        const int rc = gather_fwd_tiled(x, off, p_n, operand, dbg_idx, dbg_coord, B, C, H, W, N, s, dtype, st);
        if (rc <= 0) return rc;
    }
    if (dtype == LDCONV_F32)
        return launch_gather_fwd<float>((const float*)x, off, p_n, (float*)operand, dbg_idx, dbg_coord, B, C, H, W, N, s, st);
    return launch_gather_fwd<__nv_bfloat16>((const __nv_bfloat16*)x, off, p_n, (__nv_bfloat16*)operand, dbg_idx,
                                            dbg_coord, B, C, H, W, N, s, st);
}

// Column sums / sums of squares of an existing (M, O) bf16 tensor (fp64 accumulators, caller zero-inits): the batch statistics of a
// Conv block's pre-activation when the conv itself ran elsewhere (cuDNN 3x3); the GEMM entry points produce them in their epilogue.
LDC_API int ldconv_col_stats(const void* pre, double* stat_sum, double* stat_sqsum, long long M, int O, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_col_stats: bf16 only");
    LDC_REQUIRE(pre && stat_sum && stat_sqsum && M >= 0 && O >= 1, "ldconv_col_stats: bad arguments");
    if (M == 0) return LDCONV_OK;
    return col_stats_bf16((const __nv_bfloat16*)pre, M, O, stat_sum, stat_sqsum, (cudaStream_t)stream);
}

LDC_API int ldconv_bn_finalize(const double* stat_sum, const double* stat_sqsum, long long count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float eps, float momentum,
                               int training, float* scale, float* shift, float* save_mean, float* save_invstd, int O,
                               void* stream)
{
    LDC_REQUIRE(O >= 1 && scale && shift, "ldconv_bn_finalize: bad arguments");
    if (training)
        LDC_REQUIRE(stat_sum && stat_sqsum && count >= 1, "ldconv_bn_finalize: training needs batch sums and count >= 1");
    else
        LDC_REQUIRE(running_mean && running_var, "ldconv_bn_finalize: eval needs running statistics");
    bn_finalize_kernel<<<cdiv(O, 128), 128, 0, (cudaStream_t)stream>>>(stat_sum, stat_sqsum, count, gamma, beta,
                                                                       running_mean, running_var, eps, momentum,
                                                                       training, scale, shift, save_mean, save_invstd, O);
    LDC_LAUNCH_CHECK("bn_finalize_kernel");
    return LDCONV_OK;
}

template <typename T>
static int bn_act_apply_t(const T* pre, const float* scale, const float* shift, T* out, long long M, int O, int act,
                          cudaStream_t st)
{
    constexpr int V = Vec16<T>::N;
    const bool vec = (O % V == 0) && aligned16(pre) && aligned16(out);
    const long long nvec = vec ? M * O / V : M * O;
    if (nvec == 0) return LDCONV_OK;
    long long want = (nvec + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    const int CVn = vec ? O / V : 0;
    if (vec && CVn <= 256) {      // threads and grid stride are multiples of CVn: a thread keeps its column vector
        const int threads = 256 / CVn * CVn;
        long long nb = (nvec + (long long)threads * kColUnroll - 1) / ((long long)threads * kColUnroll);
        if (nb > cap) nb = cap;
        bn_act_apply_cols_kernel<T><<<(unsigned)nb, threads, 0, st>>>(pre, scale, shift, out, nvec, CVn, act);
    } else if (vec)
        bn_act_apply_kernel<T, true><<<blocks, 256, 0, st>>>(pre, scale, shift, out, nvec, O, act);
    else
        bn_act_apply_kernel<T, false><<<blocks, 256, 0, st>>>(pre, scale, shift, out, nvec, O, act);
    LDC_LAUNCH_CHECK("bn_act_apply_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_bn_act_apply(const void* pre, const float* scale, const float* shift, void* out, long long M, int O,
                                int act, int dtype, void* stream)
{
    if (int e = check_dtype("ldconv_bn_act_apply", dtype)) return e;
    LDC_REQUIRE(pre && scale && shift && out && M >= 0 && O >= 1, "ldconv_bn_act_apply: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_F32) return bn_act_apply_t<float>((const float*)pre, scale, shift, (float*)out, M, O, act, st);
    return bn_act_apply_t<__nv_bfloat16>((const __nv_bfloat16*)pre, scale, shift, (__nv_bfloat16*)out, M, O, act, st);
}

template <typename T>
static int bn_act_bwd_reduce_t(const T* pre, const T* gout, const float* scale, const float* shift, const float* mean,
                               const float* invstd, double* red, long long M, int O, int act, cudaStream_t st)
{
    constexpr int V = Vec16<T>::N;
    const bool vec = (O % V == 0) && aligned16(pre) && aligned16(gout);
    const int cvn = vec ? O / V : O;
    const int tile = pow2_at_least(cvn < 256 ? cvn : 256);
    const int phases = 256 / tile;
    long long want = (M + (long long)phases * 8 - 1) / ((long long)phases * 8);
    const long long cap = (long long)num_sms() * 2;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
    const int vper = vec ? V : 1;
    for (int col0 = 0; col0 < O; col0 += tile * vper) {
        if (vec)
            bn_act_bwd_reduce_kernel<T, true><<<blocks, 256, 0, st>>>(pre, gout, scale, shift, mean, invstd, red, M, O, act,
                                                                      tile, col0);
        else
            bn_act_bwd_reduce_kernel<T, false><<<blocks, 256, 0, st>>>(pre, gout, scale, shift, mean, invstd, red, M, O,
                                                                       act, tile, col0);
        LDC_LAUNCH_CHECK("bn_act_bwd_reduce_kernel");
    }
    return LDCONV_OK;
}

LDC_API int ldconv_bn_act_bwd_reduce(const void* pre, const void* grad_out, const float* scale, const float* shift,
                                     const float* mean, const float* invstd, double* red, long long M, int O, int act,
                                     int dtype, void* stream)
{
    if (int e = check_dtype("ldconv_bn_act_bwd_reduce", dtype)) return e;
    LDC_REQUIRE(pre && grad_out && scale && shift && mean && invstd && red && M >= 1 && O >= 1,
                "ldconv_bn_act_bwd_reduce: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_F32)
        return bn_act_bwd_reduce_t<float>((const float*)pre, (const float*)grad_out, scale, shift, mean, invstd, red, M, O,
                                          act, st);
    return bn_act_bwd_reduce_t<__nv_bfloat16>((const __nv_bfloat16*)pre, (const __nv_bfloat16*)grad_out, scale, shift,
                                              mean, invstd, red, M, O, act, st);
}

template <typename T>
static int bn_act_bwd_apply_t(const T* pre, const T* gout, const float* scale, const float* shift, const float* mean,
                              const float* invstd, const double* red, T* gpre, long long M, int O, int act, int training,
                              cudaStream_t st)
{
    constexpr int V = Vec16<T>::N;
    const bool vec = (O % V == 0) && aligned16(pre) && aligned16(gout) && aligned16(gpre);
    const long long nvec = vec ? M * O / V : M * O;
    long long want = (nvec + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    const int CVn = vec ? O / V : 0;
    if (vec && CVn <= 256) {
        const int threads = 256 / CVn * CVn;
        long long nb = (nvec + (long long)threads * kColUnroll - 1) / ((long long)threads * kColUnroll);
        if (nb > cap) nb = cap;
        bn_act_bwd_apply_cols_kernel<T><<<(unsigned)nb, threads, 0, st>>>(pre, gout, scale, shift, mean, invstd, red, gpre, nvec, M, O,
                                                                           CVn, act, training);
    } else if (vec)
        bn_act_bwd_apply_kernel<T, true><<<blocks, 256, 0, st>>>(pre, gout, scale, shift, mean, invstd, red, gpre, nvec, M,
                                                                 O, act, training);
    else
        bn_act_bwd_apply_kernel<T, false><<<blocks, 256, 0, st>>>(pre, gout, scale, shift, mean, invstd, red, gpre, nvec,
                                                                  M, O, act, training);
    LDC_LAUNCH_CHECK("bn_act_bwd_apply_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_bn_act_bwd_apply(const void* pre, const void* grad_out, const float* scale, const float* shift,
                                    const float* mean, const float* invstd, const double* red, void* grad_pre,
                                    long long M, int O, int act, int training, int dtype, void* stream)
{
    if (int e = check_dtype("ldconv_bn_act_bwd_apply", dtype)) return e;
    LDC_REQUIRE(pre && grad_out && scale && shift && mean && invstd && grad_pre && M >= 1 && O >= 1,
                "ldconv_bn_act_bwd_apply: bad arguments");
    LDC_REQUIRE(!training || red, "ldconv_bn_act_bwd_apply: training needs the pass-1 sums");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_F32)
        return bn_act_bwd_apply_t<float>((const float*)pre, (const float*)grad_out, scale, shift, mean, invstd, red,
                                         (float*)grad_pre, M, O, act, training, st);
    return bn_act_bwd_apply_t<__nv_bfloat16>((const __nv_bfloat16*)pre, (const __nv_bfloat16*)grad_out, scale, shift,
                                             mean, invstd, red, (__nv_bfloat16*)grad_pre, M, O, act, training, st);
}


template <typename T, typename ACC>
static int gather_bwd_t(const T* gop, const T* x, const float* off, const int* pn, ACC* grad_x, float* grad_off, int B,
                        int C, int H, int W, int N, int s, cudaStream_t st)
{
    const int h = out_size(H, s), w = out_size(W, s);
    constexpr int V = Vec16<T>::N;
    const bool vec = (C % V == 0) && aligned16(x) && aligned16(gop) && (!grad_x || aligned16(grad_x));
    if (sizeof(ACC) == 2 && !vec)
        return fail(LDCONV_E_ARG, "ldconv_gather_bwd_acc16: needs C %% 8 == 0 and 16-byte aligned tensors (C=%d)", C);
    const int CV = vec ? C / V : C;
    const long long total = (long long)B * h * w * N * CV;
    LDC_CUDA(cudaMemsetAsync(grad_off, 0, (size_t)B * h * w * 2 * N * sizeof(float), st));
    if (total == 0) return LDCONV_OK;
    const int group = (CV <= 32 && (CV & (CV - 1)) == 0) ? CV : 1;
    const unsigned blocks = cdiv(total, 256);
    const bool small = total + 256 < 0xffffffffll;      // the last block's thread ids must not wrap
    if (vec && small)
        gather_bwd_kernel<T, true, unsigned, ACC><<<blocks, 256, 0, st>>>(gop, x, off, pn, grad_x, grad_off, C, H, W, h, w, N, s, CV,
                                                                          group, total);
    else if (vec)
        gather_bwd_kernel<T, true, long long, ACC><<<blocks, 256, 0, st>>>(gop, x, off, pn, grad_x, grad_off, C, H, W, h, w, N, s, CV,
                                                                           group, total);
    else if (small)
        gather_bwd_kernel<T, false, unsigned, ACC><<<blocks, 256, 0, st>>>(gop, x, off, pn, grad_x, grad_off, C, H, W, h, w, N, s, CV,
                                                                           group, total);
    else
        gather_bwd_kernel<T, false, long long, ACC><<<blocks, 256, 0, st>>>(gop, x, off, pn, grad_x, grad_off, C, H, W, h, w, N, s, CV,
                                                                            group, total);
    LDC_LAUNCH_CHECK("gather_bwd_kernel");
    return LDCONV_OK;
}

LDC_API int ldconv_gather_bwd(const void* grad_operand, const void* x, const float* off, const int32_t* p_n,
                              float* grad_x, float* grad_off, int B, int C, int H, int W, int N, int s, int dtype,
                              void* stream)
{
    if (int e = check_dims("ldconv_gather_bwd", B, C, H, W, N, s)) return e;
    if (int e = check_dtype("ldconv_gather_bwd", dtype)) return e;
    LDC_REQUIRE(grad_operand && x && off && p_n && grad_off, "ldconv_gather_bwd: null pointer");
    if (B == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_F32)
        return gather_bwd_t<float, float>((const float*)grad_operand, (const float*)x, off, p_n, grad_x, grad_off, B, C, H, W, N, s,
                                          st);
    return gather_bwd_t<__nv_bfloat16, float>((const __nv_bfloat16*)grad_operand, (const __nv_bfloat16*)x, off, p_n, grad_x,
                                              grad_off, B, C, H, W, N, s, st);
}

LDC_API int ldconv_gather_bwd_acc16(const void* grad_operand, const void* x, const float* off, const int32_t* p_n, void* grad_x,
                                    float* grad_off, int B, int C, int H, int W, int N, int s, void* stream)
{
    if (int e = check_dims("ldconv_gather_bwd_acc16", B, C, H, W, N, s)) return e;
    LDC_REQUIRE(grad_operand && x && off && p_n && grad_off, "ldconv_gather_bwd_acc16: null pointer");
    if (B == 0) return LDCONV_OK;
    {   // TMA-tiled kernel for the model's shapes (ldconv_gather_tma.cu); 1 = no instance for this shape
        const int e = scatter_bwd_tiled((const __nv_bfloat16*)x, off, p_n, (const __nv_bfloat16*)grad_operand, (__nv_bfloat16*)grad_x,
                                        grad_off, B, C, H, W, N, s, (cudaStream_t)stream);
        if (e != 1) return e;
    }
    return gather_bwd_t<__nv_bfloat16, __nv_bfloat16>((const __nv_bfloat16*)grad_operand, (const __nv_bfloat16*)x, off, p_n,
                                                      (__nv_bfloat16*)grad_x, grad_off, B, C, H, W, N, s, (cudaStream_t)stream);
}

LDC_API int ldconv_offset_conv_bwd(const float* grad_off, const void* x, const float* w, float* grad_x, float* grad_w,
                                   float* grad_b, int B, int C, int H, int W, int N, int s, int dtype, void* stream)
{
    if (int e = check_dims("ldconv_offset_conv_bwd", B, C, H, W, N, s)) return e;
    if (int e = check_dtype("ldconv_offset_conv_bwd", dtype)) return e;
    LDC_REQUIRE(grad_off && x && w, "ldconv_offset_conv_bwd: null pointer");
    if (B == 0) return LDCONV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LDCONV_F32)
        return dispatch_offset_conv_bwd<float>(grad_off, (const float*)x, w, grad_x, grad_w, grad_b, B, C, H, W, N, s, st);
    return dispatch_offset_conv_bwd<__nv_bfloat16>(grad_off, (const __nv_bfloat16*)x, w, grad_x, grad_w, grad_b, B, C, H,
                                                   W, N, s, st);
}

LDC_API int ldconv_detect_decode(const void* box, const void* cls, void* y, int B, int H, int W, int nc, int reg_max,
                                 float stride, int anchor_offset, int total_anchors, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_detect_decode: bf16 only");
    LDC_REQUIRE(box && cls && y && B >= 0 && H >= 1 && W >= 1 && nc >= 1, "ldconv_detect_decode: bad arguments");
    LDC_REQUIRE(reg_max == 16, "ldconv_detect_decode: reg_max %d not supported (16)", reg_max);
    LDC_REQUIRE(aligned16(box), "ldconv_detect_decode: box must be 16-byte aligned");
    LDC_REQUIRE(anchor_offset >= 0 && anchor_offset + H * W <= total_anchors, "ldconv_detect_decode: anchor range");
    if (B == 0) return LDCONV_OK;
    const long long n = (long long)B * H * W;
    constexpr int staged = 1;      // logits staged through shared memory: 98 instead of 101 us per step
    if (staged)
        detect_decode_staged_kernel<16><<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)box, (const __nv_bfloat16*)cls, (__nv_bfloat16*)y, B, H, W, nc, stride, anchor_offset, total_anchors);
    else
        detect_decode_kernel<16><<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)box, (const __nv_bfloat16*)cls, (__nv_bfloat16*)y, B, H, W, nc, stride, anchor_offset, total_anchors);
    LDC_LAUNCH_CHECK("detect_decode_kernel");
    return LDCONV_OK;
}

