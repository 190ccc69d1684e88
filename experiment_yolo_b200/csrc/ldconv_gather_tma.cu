// ldconv_gather_tma.cu -- the LDConv gather with a TMA-staged input tile (sm_100a).
//
// Replaces /root/reference/ultralytics/nn/modules/conv.py:369-407 + 413-503 (grid generation, floor / clamp, four
// gathers over a channel-expanded int64 index, bilinear sum, rearrange) -- same arithmetic as gather_fwd_kernel in
// ldconv_core.cu (shared through common.cuh::make_point / bilinear, so indices, coordinates and the fp32 operand stay
// bit-exact), different data movement:
//   * a CTA owns a TH x TW tile of output pixels of one image; ONE 4-D TMA load (cp.async.bulk.tensor, box =
//     C x TWin x THin x 1) stages the input footprint of that tile plus a halo of `halo` pixels into shared memory.
//     Out-of-image parts of the box are zero-filled by the TMA unit and never read (corner indices are clamped first).
//   * every (pixel, sample n, 16-byte channel vector) item computes its sampling point; if the four corners lie inside
//     the staged tile they are read from shared memory, otherwise (offset larger than the halo) from global memory,
//     i.e. from L2 -- offsets are unbounded in the reference (conv.py:368-372), so the halo cannot be a guarantee.
//   * the operand row m is written as consecutive 16-byte chunks (k = n*C + c): the (M, N*C) row-major layout is exactly
//     the K-major tile layout the tcgen05 GEMM's TMA loads consume.
// A per-launch counter of samples served from global memory is optional (halo miss rate for profiles/).
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

struct TileGeom {
    int TH, TW, THin, TWin, halo, tiles_h, tiles_w;
};

template <typename T>
__global__ void __launch_bounds__(256)
gather_fwd_tiled_kernel(const __grid_constant__ CUtensorMap tmX, const T* __restrict__ x, const float* __restrict__ off,
                        const int* __restrict__ pn, T* __restrict__ operand, int* __restrict__ dbg_idx,
                        float* __restrict__ dbg_coord, unsigned long long* __restrict__ miss_counter, int C, int H, int W,
                        int h, int w, int N, int s, TileGeom g)
{
    constexpr int V = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    T* tile = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);

    const int tj = blockIdx.x % g.tiles_w;
    const int ti = (blockIdx.x / g.tiles_w) % g.tiles_h;
    const int b = blockIdx.x / (g.tiles_w * g.tiles_h);
    const int i0 = ti * g.TH, j0 = tj * g.TW;
    const int r_org = i0 * s - g.halo, k_org = j0 * s - g.halo;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)((size_t)g.THin * g.TWin * C * sizeof(T)));
        tma_load_4d(tile, &tmX, &bar, 0, k_org, r_org, b);
    }

    const int CV = C / V;
    const int items = g.TH * g.TW * N * CV;
    const int r_end = r_org + g.THin, k_end = k_org + g.TWin;
    const T* xb = x + (size_t)b * H * W * C;
    unsigned misses = 0;
    bool waited = false;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int cv = it % CV;
        const int n = (it / CV) % N;
        const int p = it / (CV * N);
        const int i = i0 + p / g.TW, j = j0 + p % g.TW;
        if (i >= h || j >= w) continue;
        const long long m = ((long long)b * h + i) * w + j;
        const float* op = off + (size_t)m * 2 * N;
        const SamplePoint q = make_point(i, j, s, pn[n], pn[N + n], op[n], op[N + n], H, W);
        if (cv == 0) {
            const long long sn = m * N + n;
            if (dbg_idx) *reinterpret_cast<int4*>(dbg_idx + (size_t)sn * 4) = make_int4(q.r0, q.r1, q.k0, q.k1);
            if (dbg_coord) {
                dbg_coord[(size_t)sn * 2 + 0] = q.pcr;
                dbg_coord[(size_t)sn * 2 + 1] = q.pck;
            }
        }
        const float g_lt = __fmul_rn(q.ar0, q.ak0), g_rb = __fmul_rn(q.ar1, q.ak1);
        const float g_lb = __fmul_rn(q.ar0, q.ak1), g_rt = __fmul_rn(q.ar1, q.ak0);
        const bool inside = q.r0 >= r_org && q.r1 < r_end && q.k0 >= k_org && q.k1 < k_end;
        float x00[V], x11[V], x01[V], x10[V], r[V];
        if (inside) {
            if (!waited) {       // first use of the staged tile: wait for the TMA bytes to land
                mbar_wait(&bar, 0);
                waited = true;
            }
            const T* t0 = tile + (size_t)cv * V;
            const int ra = (q.r0 - r_org) * g.TWin, rb = (q.r1 - r_org) * g.TWin;
            const int ka = q.k0 - k_org, kb = q.k1 - k_org;
            Vec16<T>::load(t0 + (size_t)(ra + ka) * C, x00);
            Vec16<T>::load(t0 + (size_t)(rb + kb) * C, x11);
            Vec16<T>::load(t0 + (size_t)(ra + kb) * C, x01);
            Vec16<T>::load(t0 + (size_t)(rb + ka) * C, x10);
        } else {
            const T* g0 = xb + (size_t)cv * V;
            Vec16<T>::load(g0 + ((size_t)q.r0 * W + q.k0) * C, x00);
            Vec16<T>::load(g0 + ((size_t)q.r1 * W + q.k1) * C, x11);
            Vec16<T>::load(g0 + ((size_t)q.r0 * W + q.k1) * C, x01);
            Vec16<T>::load(g0 + ((size_t)q.r1 * W + q.k0) * C, x10);
            if (cv == 0) ++misses;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) r[v] = bilinear(g_lt, g_rb, g_lb, g_rt, x00[v], x11[v], x01[v], x10[v]);
        Vec16<T>::store(operand + (size_t)m * N * C + (size_t)n * C + (size_t)cv * V, r);
    }
    // the TMA write must have completed before the CTA's shared memory is released
    if (!waited) mbar_wait(&bar, 0);
    if (miss_counter) {
        misses = (unsigned)__reduce_add_sync(0xffffffffu, misses);
        if ((threadIdx.x & 31) == 0 && misses) atomicAdd(miss_counter, (unsigned long long)misses);
    }
}

static thread_local int g_gather_direct = 0;
static thread_local unsigned long long* g_miss_counter = nullptr;

int gather_set_direct(int v) { g_gather_direct = v ? 1 : 0; return LDCONV_OK; }
int gather_set_miss_counter(void* p) { g_miss_counter = (unsigned long long*)p; return LDCONV_OK; }

template <typename T>
static int gather_tiled_t(const T* x, const float* off, const int* pn, T* operand, int* dbg_idx, float* dbg_coord, int B,
                          int C, int H, int W, int N, int s, int max_pn_r, int max_pn_k, cudaStream_t st)
{
    const int h = out_size(H, s), w = out_size(W, s);
    TileGeom g;
    g.halo = 2;
    g.TH = 8;
    g.TW = 16;
    auto smem_of = [&](const TileGeom& t) {
        return (size_t)((t.TH - 1) * s + 2 + max_pn_r + 2 * t.halo) * ((t.TW - 1) * s + 2 + max_pn_k + 2 * t.halo) * C *
               sizeof(T);
    };
    while (smem_of(g) > 64 * 1024 && g.TH * g.TW > 16) {
        if (g.TW > g.TH) g.TW /= 2; else g.TH /= 2;
    }
    g.THin = (g.TH - 1) * s + 2 + max_pn_r + 2 * g.halo;
    g.TWin = (g.TW - 1) * s + 2 + max_pn_k + 2 * g.halo;
    if (g.THin > 256 || g.TWin > 256 || smem_of(g) > 200 * 1024) return 1;   // not eligible: caller uses the direct kernel
    g.tiles_h = (h + g.TH - 1) / g.TH;
    g.tiles_w = (w + g.TW - 1) / g.TW;
    const long long ctas = (long long)B * g.tiles_h * g.tiles_w;
    if (ctas > 0x7fffffffll) return 1;

    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)W * C * sizeof(T), (cuuint64_t)H * W * C * sizeof(T)};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)g.TWin, (cuuint32_t)g.THin, 1};
    const CUtensorMapDataType dt = sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (int e = encode_map(&tm, dt, 4, x, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;

    const size_t smem = smem_of(g) + 128;
    auto kern = gather_fwd_tiled_kernel<T>;
    LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)ctas, 256, smem, st>>>(tm, x, off, pn, operand, dbg_idx, dbg_coord, g_miss_counter, C, H, W, h, w, N, s,
                                            g);
    LDC_LAUNCH_CHECK("gather_fwd_tiled_kernel");
    return LDCONV_OK;
}

// returns LDCONV_OK when launched, 1 when the shape is not eligible (caller falls through to the direct-load kernel of
// ldconv_core.cu -- the same arithmetic without the staged tile), < 0 on error
int gather_fwd_tiled(const void* x, const float* off, const int* pn, void* operand, int* dbg_idx, float* dbg_coord, int B,
                     int C, int H, int W, int N, int s, int dtype, cudaStream_t st)
{
    if (g_gather_direct) return 1;
    const int V = dtype == LDCONV_BF16 ? 8 : 4;
    if (C % V != 0 || C > 256 || !aligned16(x) || !aligned16(operand)) return 1;
    int32_t table[64];
    if (N > 16) return 1;
    if (int e = ldconv_p_n(N, table)) return e;
    int max_r = 0, max_k = 0;
    for (int n = 0; n < N; ++n) {
        if (table[n] > max_r) max_r = table[n];
        if (table[N + n] > max_k) max_k = table[N + n];
    }
    if (dtype == LDCONV_BF16)
        return gather_tiled_t<__nv_bfloat16>((const __nv_bfloat16*)x, off, pn, (__nv_bfloat16*)operand, dbg_idx, dbg_coord,
                                             B, C, H, W, N, s, max_r, max_k, st);
    return gather_tiled_t<float>((const float*)x, off, pn, (float*)operand, dbg_idx, dbg_coord, B, C, H, W, N, s, max_r,
                                 max_k, st);
}

}  // namespace ldc
