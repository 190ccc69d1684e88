// ldconv_fused_umma.cu -- the whole LDConv inference forward as ONE tcgen05 kernel (bf16, sm_100a).
//
// Replaces /root/reference/ultralytics/nn/modules/conv.py:366-410 (eval mode).  A CTA owns a TH x TW = 128-pixel tile of
// one image's output and runs five phases on shared memory; nothing but x is read from and nothing but the activated
// output is written to HBM (algorithmic bytes per call: e*B*C*H*W + e*B*h*w*O, vs. x twice + offsets twice + the
// (M, N*C) operand twice for the three-kernel path):
//   1. TMA   one 4-D cp.async.bulk.tensor stages the input footprint + halo (zero-filled outside the image = the conv's
//            padding) ; 2-D TMA loads stage the (O, K) weights K-block by K-block in the 128B-swizzled UMMA layout.
//   2. FFMA  offset conv (conv.py:368) from the staged tile: thread = (pixel pair, channel quarter), fp32 accumulation,
//            partial sums combined through shared memory.
//   3. LDS   sampling grid / clamps / bilinear weights (common.cuh::make_point, bit-identical to the other kernels),
//            four-corner gather from the staged tile (global/L2 when a corner leaves the halo), bilinear sum, bf16
//            rounding, and a 16-byte store straight into the K-major SWIZZLE_128B operand tile tcgen05 reads.
//   4. UMMA  one elected thread issues tcgen05.mma (M=128, N=O, K=16) over K = N*C, accumulator in TMEM.
//   5. LDTM  all 8 warps read their TMEM lane group, apply folded BatchNorm + SiLU, store bf16 NHWC rows.
// CTAs are independent (one tile each); several are resident per SM, so phases of different tiles overlap.
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

struct FusedGeom {
    int TH, TW, THin, TWin, halo, tiles_h, tiles_w;
    int num_kb, ON, ON2, Q, PP;             // K blocks of 64, O padded to 16, 2N padded to 4, channel split, pixels/thread
    uint32_t ofs_b, ofs_x, ofs_woff, ofs_part, ofs_aff, ofs_bar;   // byte offsets inside the 1024-aligned dynamic smem
    uint32_t tmem_cols;
    int dbg;
};

static constexpr int kFusedThreads = 256;

// debug phase timeline (LDCONV_DBG bit 64): thread 0 of one mid-grid CTA records clock64 after each phase
__device__ long long g_fused_trace[16];

template <int ON2>
__global__ void __launch_bounds__(kFusedThreads)
fused_umma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_off, const float* __restrict__ b_off,
                  const int* __restrict__ pn, const float* __restrict__ scale, const float* __restrict__ shift,
                  __nv_bfloat16* __restrict__ out, float* __restrict__ off_out, int C, int H, int W, int h, int w, int N,
                  int s, int O, int act, FusedGeom g)
{
    using T = __nv_bfloat16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* sA = smem;                                            // [num_kb][128 rows][128 B], SWIZZLE_128B
    uint8_t* sB = smem + g.ofs_b;                                  // [num_kb][ON rows][128 B], SWIZZLE_128B (TMA)
    T* sX = reinterpret_cast<T*>(smem + g.ofs_x);                  // [THin][TWin][C]
    float* sWoff = reinterpret_cast<float*>(smem + g.ofs_woff);    // [9][C][ON2]
    float* sPart = reinterpret_cast<float*>(smem + g.ofs_part);    // [Q][128][ON2]
    float2* sAff = reinterpret_cast<float2*>(smem + g.ofs_aff);    // [ON] (scale, shift) of the folded BatchNorm
    uint64_t* bar_x = reinterpret_cast<uint64_t*>(smem + g.ofs_bar);
    uint64_t* bar_w = bar_x + 1;
    uint64_t* bar_mma = bar_x + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_x + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tj = blockIdx.x % g.tiles_w;
    const int ti = (blockIdx.x / g.tiles_w) % g.tiles_h;
    const int b = blockIdx.x / (g.tiles_w * g.tiles_h);
    const int i0 = ti * g.TH, j0 = tj * g.TW;
    const int r_org = i0 * s - g.halo, k_org = j0 * s - g.halo;
    const int O2 = 2 * N, K = N * C;
    const int b_bytes = g.ON * 128;

    const bool tracing = g.dbg && tid == 0 && blockIdx.x == gridDim.x / 2;
    long long tr[10];
    if (tracing) tr[0] = clock64();
    // ---- phase 0: barriers, TMEM, TMA issue, offset-conv weights -------------------------------------------------------
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        mbar_init(bar_x, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, g.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (tracing) tr[1] = clock64();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_x, (uint32_t)((size_t)g.THin * g.TWin * C * sizeof(T)));
        tma_load_4d(sX, &tmX, bar_x, 0, k_org, r_org, b);
        mbar_arrive_expect_tx(bar_w, (uint32_t)(g.num_kb * b_bytes));
        for (int kb = 0; kb < g.num_kb; ++kb) tma_load_2d(sB + (size_t)kb * b_bytes, &tmW, bar_w, kb * 64, 0);
    }
    for (int t = tid; t < 9 * C * ON2; t += kFusedThreads) {
        const int o = t % ON2, tc = t / ON2;
        sWoff[t] = o < O2 ? w_off[(size_t)tc * O2 + o] : 0.f;
    }
    for (int o = tid; o < g.ON; o += kFusedThreads)
        sAff[o] = make_float2((scale && o < O) ? scale[o] : 1.f, (shift && o < O) ? shift[o] : 0.f);
    __syncthreads();
    if (tracing) tr[2] = clock64();
    mbar_wait(bar_x, 0);
    if (tracing) tr[3] = clock64();

    // ---- phase 2: offset conv.  thread = (pixel group pg, channel slice q); PP pixels per thread share weight loads -------
    {
        const int groups = 128 / g.PP;               // pixel groups
        const int pg = tid % groups, q = tid / groups;
        const int cq = C / g.Q;                      // channels per slice (multiple of 8)
        float acc[2][ON2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int o = 0; o < ON2; ++o) acc[pp][o] = 0.f;
        int xbase[2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const int p = pg + pp * groups;          // pp == 1 only used when PP == 2
            const int pi = p / g.TW, pj = p % g.TW;
            xbase[pp] = ((pi * s + g.halo - 1) * g.TWin + (pj * s + g.halo - 1)) * C + q * cq;
        }
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int tofs = ((tap / 3) * g.TWin + (tap % 3)) * C;
            const float* wp = sWoff + (size_t)(tap * C + q * cq) * ON2;
            for (int cc = 0; cc < cq; cc += 8) {
                float xv[2][8];
                Vec16<T>::load(sX + xbase[0] + tofs + cc, xv[0]);
                if (g.PP == 2) Vec16<T>::load(sX + xbase[1] + tofs + cc, xv[1]);
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const float4* w4 = reinterpret_cast<const float4*>(wp + (cc + v) * ON2);
#pragma unroll
                    for (int o4 = 0; o4 < ON2 / 4; ++o4) {
                        const float4 wv = w4[o4];
                        acc[0][o4 * 4 + 0] = fmaf(xv[0][v], wv.x, acc[0][o4 * 4 + 0]);
                        acc[0][o4 * 4 + 1] = fmaf(xv[0][v], wv.y, acc[0][o4 * 4 + 1]);
                        acc[0][o4 * 4 + 2] = fmaf(xv[0][v], wv.z, acc[0][o4 * 4 + 2]);
                        acc[0][o4 * 4 + 3] = fmaf(xv[0][v], wv.w, acc[0][o4 * 4 + 3]);
                        if (g.PP == 2) {
                            acc[1][o4 * 4 + 0] = fmaf(xv[1][v], wv.x, acc[1][o4 * 4 + 0]);
                            acc[1][o4 * 4 + 1] = fmaf(xv[1][v], wv.y, acc[1][o4 * 4 + 1]);
                            acc[1][o4 * 4 + 2] = fmaf(xv[1][v], wv.z, acc[1][o4 * 4 + 2]);
                            acc[1][o4 * 4 + 3] = fmaf(xv[1][v], wv.w, acc[1][o4 * 4 + 3]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            if (pp < g.PP) {
                const int p = pg + pp * groups;
                float4* dst = reinterpret_cast<float4*>(sPart + ((size_t)q * 128 + p) * ON2);
#pragma unroll
                for (int o4 = 0; o4 < ON2 / 4; ++o4)
                    dst[o4] = make_float4(acc[pp][o4 * 4], acc[pp][o4 * 4 + 1], acc[pp][o4 * 4 + 2], acc[pp][o4 * 4 + 3]);
            }
        }
    }
    __syncthreads();
    if (tracing) tr[4] = clock64();

    // ---- phase 3: grid + gather + bilinear -> swizzled operand tile -----------------------------------------------------------
    {
        const int CV = C / 8;
        const int items = 128 * N * CV;
        const int r_end = r_org + g.THin, k_end = k_org + g.TWin;
        const T* xb = x + (size_t)b * H * W * C;
        for (int it = tid; it < items; it += kFusedThreads) {
            const int cv = it % CV;
            const int n = (it / CV) % N;
            const int p = it / (CV * N);
            const int i = i0 + p / g.TW, j = j0 + p % g.TW;
            if (i >= h || j >= w) continue;
            float off_r = b_off ? b_off[n] : 0.f, off_k = b_off ? b_off[N + n] : 0.f;
            for (int q = 0; q < g.Q; ++q) {
                const float* pp = sPart + ((size_t)q * 128 + p) * ON2;
                off_r += pp[n];
                off_k += pp[N + n];
            }
            if (off_out != nullptr && cv == 0) {
                float* op = off_out + ((((size_t)b * h + i) * w + j) * O2);
                op[n] = off_r;
                op[N + n] = off_k;
            }
            const SamplePoint qd = make_point(i, j, s, pn[n], pn[N + n], off_r, off_k, H, W);
            const float g_lt = __fmul_rn(qd.ar0, qd.ak0), g_rb = __fmul_rn(qd.ar1, qd.ak1);
            const float g_lb = __fmul_rn(qd.ar0, qd.ak1), g_rt = __fmul_rn(qd.ar1, qd.ak0);
            const bool inside = qd.r0 >= r_org && qd.r1 < r_end && qd.k0 >= k_org && qd.k1 < k_end;
            float x00[8], x11[8], x01[8], x10[8], r[8];
            if (inside) {
                const T* t0 = sX + (size_t)cv * 8;
                const int ra = (qd.r0 - r_org) * g.TWin, rb = (qd.r1 - r_org) * g.TWin;
                const int ka = qd.k0 - k_org, kb = qd.k1 - k_org;
                Vec16<T>::load(t0 + (size_t)(ra + ka) * C, x00);
                Vec16<T>::load(t0 + (size_t)(rb + kb) * C, x11);
                Vec16<T>::load(t0 + (size_t)(ra + kb) * C, x01);
                Vec16<T>::load(t0 + (size_t)(rb + ka) * C, x10);
            } else {
                const T* g0 = xb + (size_t)cv * 8;
                Vec16<T>::load(g0 + ((size_t)qd.r0 * W + qd.k0) * C, x00);
                Vec16<T>::load(g0 + ((size_t)qd.r1 * W + qd.k1) * C, x11);
                Vec16<T>::load(g0 + ((size_t)qd.r0 * W + qd.k1) * C, x01);
                Vec16<T>::load(g0 + ((size_t)qd.r1 * W + qd.k0) * C, x10);
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) r[v] = bilinear(g_lt, g_rb, g_lb, g_rt, x00[v], x11[v], x01[v], x10[v]);
            const int k = n * C + cv * 8;
            T* dst = reinterpret_cast<T*>(sA + (size_t)(k >> 6) * 16384 + sw128_offset((uint32_t)p, (uint32_t)((k & 63) >> 3)));
            Vec16<T>::store(dst, r);
        }
    }
    fence_proxy_async_smem();      // generic-proxy writes of the operand tile -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tracing) tr[5] = clock64();

    // ---- phase 4: MMA ----------------------------------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_wait(bar_w, 0);
        tc_fence_after_sync();
        const uint32_t idesc = make_idesc_bf16(128, g.ON);
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
        const int steps = K / 16;
        for (int st = 0; st < steps; ++st) {
            const int kb = st >> 2, kk = st & 3;
            mma_bf16_ss(tmem_base, make_desc_k_sw128(a_addr + kb * 16384 + kk * 32),
                        make_desc_k_sw128(b_addr + kb * b_bytes + kk * 32), idesc, (uint32_t)(st != 0));
        }
        mma_commit(bar_mma);
        if (tracing) tr[6] = clock64();
    }

    // ---- phase 5: epilogue --------------------------------------------------------------------------------------------------
    mbar_wait(bar_mma, 0);
    tc_fence_after_sync();
    if (tracing) tr[7] = clock64();
    {
        const int lg = warp & 3, half = warp >> 2;
        const int p = lg * 32 + lane;
        const int i = i0 + p / g.TW, j = j0 + p % g.TW;
        const bool valid = i < h && j < w;
        const size_t m = ((size_t)b * h + i) * w + j;
        const int chunks = g.ON / 16;
        const int c_begin = half * ((chunks + 1) / 2), c_end = half == 0 ? (chunks + 1) / 2 : chunks;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
        for (int ch = c_begin; ch < c_end; ++ch) {
            const int c0 = ch * 16;
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (valid && c0 < O) {
                float lo[8], hi[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float2 a0 = sAff[c0 + e], a1 = sAff[c0 + 8 + e];
                    const float z0 = fmaf(__uint_as_float(v[e]), a0.x, a0.y);
                    const float z1 = fmaf(__uint_as_float(v[8 + e]), a1.x, a1.y);
                    lo[e] = apply_act_fast(z0, act);
                    hi[e] = apply_act_fast(z1, act);
                }
                T* dst = out + m * O + c0;
                Vec16<T>::store(dst, lo);
                Vec16<T>::store(dst + 8, hi);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tracing) {
        tr[8] = clock64();
        for (int i = 0; i < 9; ++i) g_fused_trace[i] = tr[i];
    }
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

static int fused_geometry(int C, int N, int s, int O, int max_pn_r, int max_pn_k, int h, int w, FusedGeom* out, size_t* smem)
{
    FusedGeom g;
    const int K = N * C;
    g.num_kb = (K + 63) / 64;
    g.ON = (O + 15) / 16 * 16;
    g.ON2 = (2 * N + 3) / 4 * 4;
    g.PP = (C % 32 == 0) ? 2 : 1;
    g.Q = 2 * g.PP;
    // tile orientation: prefer the one that wastes fewer pixels at the right / bottom border
    auto waste = [&](int th, int tw) {
        return (long long)((h + th - 1) / th * th) * ((w + tw - 1) / tw * tw) - (long long)h * w;
    };
    g.TH = 8; g.TW = 16;
    if (waste(16, 8) < waste(8, 16)) { g.TH = 16; g.TW = 8; }
    for (g.halo = 2; g.halo >= 1; --g.halo) {
        g.THin = (g.TH - 1) * s + 2 + max_pn_r + 2 * g.halo;
        g.TWin = (g.TW - 1) * s + 2 + max_pn_k + 2 * g.halo;
        uint32_t ofs = (uint32_t)g.num_kb * 16384;
        g.ofs_b = ofs; ofs += (uint32_t)g.num_kb * g.ON * 128;
        ofs = (ofs + 127) & ~127u;
        g.ofs_x = ofs; ofs += (uint32_t)g.THin * g.TWin * C * 2;
        ofs = (ofs + 15) & ~15u;
        g.ofs_woff = ofs; ofs += 9u * C * g.ON2 * 4;
        g.ofs_part = ofs; ofs += (uint32_t)g.Q * 128 * g.ON2 * 4;
        g.ofs_aff = ofs; ofs += (uint32_t)g.ON * 8;
        ofs = (ofs + 7) & ~7u;
        g.ofs_bar = ofs; ofs += 32;
        *smem = ofs + 1024;
        if (*smem <= 225 * 1024) break;
    }
    if (g.halo < 1) return 0;
    if (g.THin > 256 || g.TWin > 256) return 0;
    g.tiles_h = (h + g.TH - 1) / g.TH;
    g.tiles_w = (w + g.TW - 1) / g.TW;
    g.tmem_cols = 32;
    while (g.tmem_cols < (uint32_t)g.ON) g.tmem_cols <<= 1;
    { const char* e = getenv("LDCONV_DBG"); g.dbg = e ? (atoi(e) & 64) : 0; }
    *out = g;
    return 1;
}

static void pn_extent(int N, int* max_r, int* max_k)
{
    int32_t table[64];
    *max_r = *max_k = 0;
    if (N > 16 || ldconv_p_n(N, table) != LDCONV_OK) return;
    for (int n = 0; n < N; ++n) {
        if (table[n] > *max_r) *max_r = table[n];
        if (table[N + n] > *max_k) *max_k = table[N + n];
    }
}

int umma_fused_supported(int B, int C, int H, int W, int N, int s, int O, int dtype)
{
    if (dtype != LDCONV_BF16) return 0;
    if (C % 16 != 0 || C > 256 || N > 16 || N * C > 512 || O % 16 != 0 || O > 256 || 2 * N > 32) return 0;
    const int h = out_size(H, s), w = out_size(W, s);
    if ((long long)B * ((h + 7) / 8) * ((w + 7) / 8) > 0x7fffffffll) return 0;
    int mr, mk;
    pn_extent(N, &mr, &mk);
    FusedGeom g;
    size_t smem;
    return fused_geometry(C, N, s, O, mr, mk, h, w, &g, &smem);
}

template <int ON2>
static int launch_fused(const CUtensorMap& tmX, const CUtensorMap& tmW, const __nv_bfloat16* x, const float* w_off,
                        const float* b_off, const int* pn, const float* scale, const float* shift, __nv_bfloat16* out,
                        float* off_out, int B, int C, int H, int W, int h, int w, int N, int s, int O, int act,
                        const FusedGeom& g, size_t smem, cudaStream_t st)
{
    auto kern = fused_umma_kernel<ON2>;
    LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned ctas = (unsigned)((long long)B * g.tiles_h * g.tiles_w);
    kern<<<ctas, kFusedThreads, smem, st>>>(tmX, tmW, x, w_off, b_off, pn, scale, shift, out, off_out, C, H, W, h, w, N, s, O,
                                            act, g);
    LDC_LAUNCH_CHECK("fused_umma_kernel");
    return LDCONV_OK;
}

int umma_fused_fwd(const void* x, const float* w_off, const float* b_off, const int* pn, const void* wt,
                   const float* scale, const float* shift, void* out, float* off_out, int B, int C, int H, int W, int N,
                   int s, int O, int act, cudaStream_t st)
{
    const int h = out_size(H, s), w = out_size(W, s);
    int mr, mk;
    pn_extent(N, &mr, &mk);
    FusedGeom g;
    size_t smem;
    if (!fused_geometry(C, N, s, O, mr, mk, h, w, &g, &smem))
        return fail(LDCONV_E_ARG, "tcgen05 fused kernel: tile does not fit shared memory (C=%d N=%d O=%d)", C, N, O);
    if (!aligned16(x) || !aligned16(wt) || !aligned16(out))
        return fail(LDCONV_E_ALIGN, "tcgen05 fused kernel: x / wt / out must be 16-byte aligned");

    CUtensorMap tmX, tmW;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)g.TWin, (cuuint32_t)g.THin, 1};
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return e;
    }
    {
        const int K = N * C;
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)O};
        cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)g.ON};
        if (int e = encode_map(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return e;
    }
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    __nv_bfloat16* oo = (__nv_bfloat16*)out;
    set_impl(LDCONV_IMPL_TCGEN05);
#define LDC_FUSED_CASE(V)                                                                                              \
    case V:                                                                                                            \
        return launch_fused<V>(tmX, tmW, xx, w_off, b_off, pn, scale, shift, oo, off_out, B, C, H, W, h, w, N, s, O, act, g, \
                               smem, st)
    switch (g.ON2) {
        LDC_FUSED_CASE(4);
        LDC_FUSED_CASE(8);
        LDC_FUSED_CASE(12);
        LDC_FUSED_CASE(16);
        LDC_FUSED_CASE(20);
        LDC_FUSED_CASE(24);
        LDC_FUSED_CASE(28);
        LDC_FUSED_CASE(32);
        default: return fail(LDCONV_E_ARG, "tcgen05 fused kernel: num_param %d", N);
    }
#undef LDC_FUSED_CASE
}

}  // namespace ldc

LDC_API int ldconv_debug_fused_trace(long long* host_out)
{
    cudaMemcpyFromSymbol(host_out, ldc::g_fused_trace, sizeof(long long) * 9);
    return 9;
}

