// common.cuh -- shared device / host helpers of libldconv_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/ldconv_b200.h"

#define LDC_API extern "C" __attribute__((visibility("default")))

namespace ldc {

// ---- error reporting (thread-local message, negative return codes) ---------------------------------------------------
char* err_buf();
int fail(int code, const char* fmt, ...);
void set_impl(int impl);

#define LDC_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return ::ldc::fail(LDCONV_E_ARG, __VA_ARGS__); \
    } while (0)

#define LDC_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return ::ldc::fail(LDCONV_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define LDC_LAUNCH_CHECK(name)                                                                          \
    do {                                                                                                \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess)                                                                         \
            return ::ldc::fail(LDCONV_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));  \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int out_size(int H, int s) { return (H - 1) / s + 1; }
inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }
int num_sms();

// ---- programmatic dependent launch ------------------------------------------------------------------------------------------
// The inference step is ~85 short kernels (20-150 us each); with plain stream order every kernel pays its prologue (barrier
// init, TMEM allocation, tensor-map fetch, pipeline fill) after the previous one has drained.  Kernels launched through
// launch_pdl() may start while the previous kernel is still running: they execute pdl_launch_dependents() first (so the next
// launch can be scheduled as early as possible), run their prologue, and block in pdl_wait() until every prerequisite grid has
// completed and flushed -- BEFORE the first access to any global memory a predecessor may have written (parameters included:
// in training the folded BatchNorm scale / shift come from the preceding kernel).  LDCONV_PDL=0 launches without the attribute.
int pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- element types ---------------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kVec = 4;  // elements per 16-byte vector
    __device__ __forceinline__ static float to_f(float v) { return v; }
    __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of T, unpacked to / packed from fp32 lanes
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __device__ __forceinline__ static void unpack(const uint4& t, float (&v)[4]) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
    __device__ __forceinline__ static uint4 pack(const float (&v)[4]) {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static void unpack(const uint4& t, float (&v)[8]) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static uint4 pack(const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ __forceinline__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// ---- the sampling point: the ONE place the reference's coordinate arithmetic is restated on the device ---------------
// conv.py:435-454 (p = (p_0 + p_n) + offset, one rounding), :375-387 (floor, +1, independent clamps, clamped p),
// :390-393 (per-axis weights; both are 1 once p leaves [0, H-1): the reference's border doubling).
// Uses explicit round-to-nearest intrinsics so the compiler cannot contract anything into an FMA: indices, clamped
// coordinates and the fp32 weights are bit-exact with the reference (tests/test_gpu_parity.py).
struct SamplePoint {
    int r0, r1, k0, k1;
    float pcr, pck;
    float ar0, ar1, ak0, ak1;
    bool in_r, in_k;  // torch.clamp backward indicator (inclusive) on the unclamped coordinate
};

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// (ri, ki) = (i*s + pn_r, j*s + pn_k): the integer grid point p_0 + p_n; (hm, wm) = (float)(H-1), (float)(W-1)
__device__ __forceinline__ SamplePoint make_point_grid(int ri, int ki, float off_r, float off_k, float hm, float wm)
{
    SamplePoint q;
    const float pr = __fadd_rn((float)ri, off_r);
    const float pk = __fadd_rn((float)ki, off_k);
    const float fr = floorf(pr), fk = floorf(pk);
    q.r0 = (int)clampf(fr, 0.f, hm);
    q.r1 = (int)clampf(__fadd_rn(fr, 1.f), 0.f, hm);
    q.k0 = (int)clampf(fk, 0.f, wm);
    q.k1 = (int)clampf(__fadd_rn(fk, 1.f), 0.f, wm);
    q.pcr = clampf(pr, 0.f, hm);
    q.pck = clampf(pk, 0.f, wm);
    q.ar0 = __fadd_rn(1.f, __fsub_rn((float)q.r0, q.pcr));
    q.ar1 = __fsub_rn(1.f, __fsub_rn((float)q.r1, q.pcr));
    q.ak0 = __fadd_rn(1.f, __fsub_rn((float)q.k0, q.pck));
    q.ak1 = __fsub_rn(1.f, __fsub_rn((float)q.k1, q.pck));
    q.in_r = (pr >= 0.f) && (pr <= hm);
    q.in_k = (pk >= 0.f) && (pk <= wm);
    return q;
}

__device__ __forceinline__ SamplePoint make_point(int i, int j, int s, int pn_r, int pn_k, float off_r, float off_k,
                                                  int H, int W)
{
    return make_point_grid(i * s + pn_r, j * s + pn_k, off_r, off_k, (float)(H - 1), (float)(W - 1));
}

// reference summation order lt, rb, lb, rt (conv.py:402-405), products and sums rounded separately like the
// reference's element-wise kernels
__device__ __forceinline__ float bilinear(float g_lt, float g_rb, float g_lb, float g_rt, float x00, float x11,
                                          float x01, float x10)
{
    float v = __fmul_rn(g_lt, x00);
    v = __fadd_rn(v, __fmul_rn(g_rb, x11));
    v = __fadd_rn(v, __fmul_rn(g_lb, x01));
    v = __fadd_rn(v, __fmul_rn(g_rt, x10));
    return v;
}

// same four terms with fused multiply-adds: for kernels whose operand is rounded to bf16 anyway (one fp32 ulp of difference)
__device__ __forceinline__ float bilinear_fma(float g_lt, float g_rb, float g_lb, float g_rt, float x00, float x11,
                                              float x01, float x10)
{
    return fmaf(g_rt, x10, fmaf(g_lb, x01, fmaf(g_rb, x11, g_lt * x00)));
}

// ---- packed fp32 pairs (FFMA2 / FMUL2 on sm_100a): two independent IEEE fp32 operations per issued instruction, each
// lane rounds exactly like the scalar fmaf / multiply it replaces
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// bilinear() on two fp32 lanes at once: products and sums rounded separately, term order lt, rb, lb, rt -- bit-identical to
// two scalar bilinear() calls (the reference's element-wise kernels, conv.py:402-405).  The products stay scalar FMULs on
// purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (even under -fmad=false), which would drop a rounding;
// scalar mul.rn.f32 feeding add.rn.f32x2 is left alone (8 FMUL + 3 FADD2 instead of 8 FMUL + 6 FADD).
__device__ __forceinline__ uint64_t bilinear_pair(uint64_t x00, uint64_t x11, uint64_t x01, uint64_t x10, const float4& g)
{
    auto scaled = [](uint64_t x, float w) {
        float lo, hi;
        f2_unpack(x, lo, hi);
        return f2_pack(__fmul_rn(w, lo), __fmul_rn(w, hi));
    };
    uint64_t v = f2_add(scaled(x00, g.x), scaled(x11, g.y));
    v = f2_add(v, scaled(x01, g.z));
    v = f2_add(v, scaled(x10, g.w));
    return v;
}
// one 16-byte vector of T from its four corner vectors (exact reference rounding; bf16: rounded once at the end)
template <typename T> __device__ __forceinline__ uint4 bilinear_vec16(const uint4& q00, const uint4& q11, const uint4& q01,
                                                                      const uint4& q10, const float4& g);
template <> __device__ __forceinline__ uint4 bilinear_vec16<float>(const uint4& q00, const uint4& q11, const uint4& q01,
                                                                   const uint4& q10, const float4& g)
{
    auto pr = [](uint32_t lo, uint32_t hi) { return f2_pack(__uint_as_float(lo), __uint_as_float(hi)); };
    const uint64_t a = bilinear_pair(pr(q00.x, q00.y), pr(q11.x, q11.y), pr(q01.x, q01.y), pr(q10.x, q10.y), g);
    const uint64_t b = bilinear_pair(pr(q00.z, q00.w), pr(q11.z, q11.w), pr(q01.z, q01.w), pr(q10.z, q10.w), g);
    float a0, a1, b0, b1;
    f2_unpack(a, a0, a1);
    f2_unpack(b, b0, b1);
    return make_uint4(__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(b0), __float_as_uint(b1));
}
template <> __device__ __forceinline__ uint4 bilinear_vec16<__nv_bfloat16>(const uint4& q00, const uint4& q11, const uint4& q01,
                                                                           const uint4& q10, const float4& g)
{
    auto widen = [](uint32_t w) { return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };
    auto one = [&](uint32_t w00, uint32_t w11, uint32_t w01, uint32_t w10) {
        float lo, hi;
        f2_unpack(bilinear_pair(widen(w00), widen(w11), widen(w01), widen(w10), g), lo, hi);
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    };
    return make_uint4(one(q00.x, q11.x, q01.x, q10.x), one(q00.y, q11.y, q01.y, q10.y), one(q00.z, q11.z, q01.z, q10.z),
                      one(q00.w, q11.w, q01.w, q10.w));
}

// two bf16 lanes of four corner words -> the bilinear sum of bilinear_fma() per lane, rounded once to bf16x2
__device__ __forceinline__ uint32_t bilinear_bf16x2(uint32_t w00, uint32_t w11, uint32_t w01, uint32_t w10, const float4& g)
{
    auto widen = [](uint32_t w) { return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };
    uint64_t acc = f2_mul(widen(w00), f2_pack(g.x, g.x));
    acc = f2_fma(widen(w11), f2_pack(g.y, g.y), acc);
    acc = f2_fma(widen(w01), f2_pack(g.z, g.z), acc);
    acc = f2_fma(widen(w10), f2_pack(g.w, g.w), acc);
    float lo, hi;
    f2_unpack(acc, lo, hi);
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ---- epilogue of the tcgen05 kernels: 16 accumulator columns -> folded affine -> activation, on packed fp32 pairs ---------------
// sc_addr / sh_addr: shared-space addresses of the 16 fp32 scales / shifts of these columns (two separate arrays so that one
// LDS.128 brings four of each).  For SiLU the arrays hold HALF the scale / shift: silu(z) = hz + hz * tanh(hz) with hz = z / 2, and
// the halving commutes with the rounding of the FMA, so the result is bit-identical to silu_fast(fma(acc, scale, shift)).
// ~50 instructions per 16 columns instead of ~150 (an LDS per element, scalar FMAs, an activation select per element).
__device__ __forceinline__ float affine_half_for(int act) { return act == LDCONV_ACT_SILU ? 0.5f : 1.f; }

__device__ __forceinline__ void affine_act16(const uint32_t (&v)[16], uint32_t sc_addr, uint32_t sh_addr, int act, float (&z)[16])
{
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
        float4 sc, sh;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(sc_addr + e * 4));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh.x), "=f"(sh.y), "=f"(sh.z), "=f"(sh.w) : "r"(sh_addr + e * 4));
        uint64_t z0 = f2_fma(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(sc.x, sc.y), f2_pack(sh.x, sh.y));
        uint64_t z1 = f2_fma(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(sc.z, sc.w), f2_pack(sh.z, sh.w));
        float a0, a1, a2, a3;
        f2_unpack(z0, a0, a1);
        f2_unpack(z1, a2, a3);
        if (act == LDCONV_ACT_SILU) {
            float t0, t1, t2, t3;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(a2));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(a3));
            f2_unpack(f2_fma(z0, f2_pack(t0, t1), z0), a0, a1);
            f2_unpack(f2_fma(z1, f2_pack(t2, t3), z1), a2, a3);
        } else if (act == LDCONV_ACT_LEAKY01) {
            a0 = a0 > 0.f ? a0 : 0.1f * a0; a1 = a1 > 0.f ? a1 : 0.1f * a1;
            a2 = a2 > 0.f ? a2 : 0.1f * a2; a3 = a3 > 0.f ? a3 : 0.1f * a3;
        }
        z[e] = a0; z[e + 1] = a1; z[e + 2] = a2; z[e + 3] = a3;
    }
}

// The same epilogue with the 16 scales / shifts already in registers: affine_load16 issues the eight LDS.128 BEFORE the caller's
// tcgen05.wait::ld, so their latency (through an L1 data pipe that the epilogue's own 16-byte global stores keep busy: every FFMA2
// of affine_act16 sat on the short scoreboard behind them, 32 % of conv3x3_zc_kernel's stall samples, profiles/r2_ncu_zc32_64.txt)
// hides behind the TMEM load instead of following it.  Bit-identical to affine_act16.
__device__ __forceinline__ void affine_load16(uint32_t sc_addr, uint32_t sh_addr, float4 (&sc)[4], float4 (&sh)[4])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[q].x), "=f"(sc[q].y), "=f"(sc[q].z), "=f"(sc[q].w) : "r"(sc_addr + q * 16));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh[q].x), "=f"(sh[q].y), "=f"(sh[q].z), "=f"(sh[q].w) : "r"(sh_addr + q * 16));
    }
}

__device__ __forceinline__ void affine_act16_r(const uint32_t (&v)[16], const float4 (&scv)[4], const float4 (&shv)[4], int act, float (&z)[16])
{
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
        const float4 sc = scv[e >> 2], sh = shv[e >> 2];
        uint64_t z0 = f2_fma(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(sc.x, sc.y), f2_pack(sh.x, sh.y));
        uint64_t z1 = f2_fma(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(sc.z, sc.w), f2_pack(sh.z, sh.w));
        float a0, a1, a2, a3;
        f2_unpack(z0, a0, a1);
        f2_unpack(z1, a2, a3);
        if (act == LDCONV_ACT_SILU) {
            float t0, t1, t2, t3;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(a2));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(a3));
            f2_unpack(f2_fma(z0, f2_pack(t0, t1), z0), a0, a1);
            f2_unpack(f2_fma(z1, f2_pack(t2, t3), z1), a2, a3);
        } else if (act == LDCONV_ACT_LEAKY01) {
            a0 = a0 > 0.f ? a0 : 0.1f * a0; a1 = a1 > 0.f ? a1 : 0.1f * a1;
            a2 = a2 > 0.f ? a2 : 0.1f * a2; a3 = a3 > 0.f ? a3 : 0.1f * a3;
        }
        z[e] = a0; z[e + 1] = a1; z[e + 2] = a2; z[e + 3] = a3;
    }
}

// 16 fp32 values -> 16 bf16 as two 16-byte vectors
__device__ __forceinline__ void pack16_bf16(const float (&z)[16], uint4& lo, uint4& hi)
{
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(z[2 * i + 1]), "f"(z[2 * i]));
    lo = make_uint4(w[0], w[1], w[2], w[3]);
    hi = make_uint4(w[4], w[5], w[6], w[7]);
}

__device__ __forceinline__ float silu(float z) { return z / (1.f + __expf(-z)); }
// bf16 epilogues: silu(z) = z * sigmoid(z) = 0.5 z (1 + tanh(z/2)); one MUFU op, relative error ~2^-11 (below bf16's 2^-9)
__device__ __forceinline__ float silu_fast(float z)
{
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    const float hz = 0.5f * z;
    return fmaf(hz, t, hz);
}
// activation selector of the bf16 epilogues (LDCONV_ACT_*)
__device__ __forceinline__ float apply_act_fast(float z, int act)
{
    if (act == LDCONV_ACT_SILU) return silu_fast(z);
    if (act == LDCONV_ACT_LEAKY01) return z > 0.f ? z : 0.1f * z;
    return z;
}
__device__ __forceinline__ float silu_grad(float z)
{
    const float sg = 1.f / (1.f + __expf(-z));
    return sg * (1.f + z * (1.f - sg));
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ldc
