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

__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

struct TileGeom {
    int TH, TW, THin, TWin, halo, tiles_h, tiles_w;
    int tw_shift;            // TW = 1 << tw_shift
    int cv_shift;            // CV = 1 << cv_shift, or -1 when CV is not a power of two
    unsigned inv_n;          // ceil(2^16 / N): sample / N for sample < 4096
    unsigned inv_row;        // ceil(2^32 / (TW*N*CV)): item / row_items
};

// Two phases per CTA (the kernel was issue-bound when every (sample, channel vector) item redid the coordinate arithmetic
// and its integer divisions -- profiles/r1_ncu_gatherL1v1.txt: 432 instructions per item, issue slots 79 % busy):
//   phase 1  one thread per SAMPLE (pixel, n) of the tile: make_point once, the four corner positions as 16-byte-vector
//            offsets (into the staged tile, or into the image when a corner lies outside tile + halo) and the four bilinear
//            weights go to shared memory.  Runs while the TMA load of the tile is in flight.
//   phase 2  one thread per ITEM (sample, 16-byte channel vector): two broadcast record loads, four 16-byte corner loads,
//            the bilinear sum, one 16-byte store.  Items of one tile row are a contiguous run of the operand, so the item
//            index maps linearly onto the output address.
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
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int CV = C / V;
    const int samples = g.TH * g.TW * N;
    const size_t tile_bytes = ((size_t)g.THin * g.TWin * C * sizeof(T) + 15) & ~(size_t)15;
    const uint32_t tile_s = smem_u32(sm);                                  // explicit shared-space addresses: LDS / STS
    const uint32_t rec_o = tile_s + (uint32_t)tile_bytes;                   // int4 per sample: corner offsets
    const uint32_t rec_g = rec_o + (uint32_t)samples * 16u;                 // float4 per sample: bilinear weights

    const int tj = blockIdx.x % g.tiles_w;
    const int ti = (blockIdx.x / g.tiles_w) % g.tiles_h;
    const int b = blockIdx.x / (g.tiles_w * g.tiles_h);
    const int i0 = ti * g.TH, j0 = tj * g.TW;
    const int r_org = i0 * s - g.halo, k_org = j0 * s - g.halo;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bar, (uint32_t)((size_t)g.THin * g.TWin * C * sizeof(T)));
        tma_load_4d(sm, &tmX, &bar, 0, k_org, r_org, b);
    }

    // ---- phase 1: sampling points of the tile ------------------------------------------------------------------------
    const int r_end = r_org + g.THin, k_end = k_org + g.TWin;
    unsigned misses = 0;
    for (int sidx = threadIdx.x; sidx < samples; sidx += blockDim.x) {
        const int p = (int)(((unsigned)sidx * g.inv_n) >> 16);
        const int n = sidx - p * N;
        const int i = i0 + (p >> g.tw_shift), j = j0 + (p & (g.TW - 1));
        if (i >= h || j >= w) {
            sts128(rec_o + sidx * 16, 0xffffffffu, 0, 0, 0);
            continue;
        }
        const long long m = ((long long)b * h + i) * w + j;
        const float* op = off + (size_t)m * 2 * N;
        const SamplePoint q = make_point(i, j, s, pn[n], pn[N + n], op[n], op[N + n], H, W);
        const long long sn = m * N + n;
        if (dbg_idx) *reinterpret_cast<int4*>(dbg_idx + (size_t)sn * 4) = make_int4(q.r0, q.r1, q.k0, q.k1);
        if (dbg_coord) {
            dbg_coord[(size_t)sn * 2 + 0] = q.pcr;
            dbg_coord[(size_t)sn * 2 + 1] = q.pck;
        }
        sts128(rec_g + sidx * 16, __float_as_uint(__fmul_rn(q.ar0, q.ak0)), __float_as_uint(__fmul_rn(q.ar1, q.ak1)),
               __float_as_uint(__fmul_rn(q.ar0, q.ak1)), __float_as_uint(__fmul_rn(q.ar1, q.ak0)));
        const bool inside = q.r0 >= r_org && q.r1 < r_end && q.k0 >= k_org && q.k1 < k_end;
        if (inside) {
            const int ra = (q.r0 - r_org) * g.TWin, rb = (q.r1 - r_org) * g.TWin;
            const int ka = q.k0 - k_org, kb = q.k1 - k_org;
            sts128(rec_o + sidx * 16, (uint32_t)((ra + ka) * CV), (uint32_t)((rb + kb) * CV), (uint32_t)((ra + kb) * CV),
                   (uint32_t)((rb + ka) * CV));
        } else {   // served from global memory (L2): offsets relative to the image, bit 31 of .x marks it
            const int ra = q.r0 * W, rb = q.r1 * W;
            sts128(rec_o + sidx * 16, (uint32_t)((ra + q.k0) * CV) | 0x80000000u, (uint32_t)((rb + q.k1) * CV),
                   (uint32_t)((ra + q.k1) * CV), (uint32_t)((rb + q.k0) * CV));
            ++misses;
        }
    }
    __syncthreads();
    mbar_wait(&bar, 0);      // the staged tile has landed (also: the TMA write completes before the CTA's smem is released)

    // ---- phase 2: bilinear resampling, one 16-byte channel vector per item ----------------------------------------------
    const int row_items = g.TW * N * CV;
    const int items = g.TH * row_items;
    const uint4* xb4 = reinterpret_cast<const uint4*>(x + (size_t)b * H * W * C);
    // operand row of tile pixel (0,0) in 16-byte vectors; a tile row is a contiguous run of row_items vectors, consecutive
    // tile rows are w*N*CV vectors apart
    uint4* out4 = reinterpret_cast<uint4*>(operand) + (((long long)b * h + i0) * w + j0) * (long long)(N * CV);
    const long long row_stride = (long long)w * N * CV;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int sidx = g.cv_shift >= 0 ? (it >> g.cv_shift) : it / CV;
        const int cv = it - sidx * CV;
        const uint4 o = lds128(rec_o + sidx * 16);
        if (o.x == 0xffffffffu) continue;
        const uint4 gq = lds128(rec_g + sidx * 16);
        const float4 gw = make_float4(__uint_as_float(gq.x), __uint_as_float(gq.y), __uint_as_float(gq.z), __uint_as_float(gq.w));
        uint4 q00, q11, q01, q10;
        if ((int)o.x >= 0) {
            const uint32_t t0 = tile_s + (uint32_t)cv * 16u;
            q00 = lds128(t0 + o.x * 16u); q11 = lds128(t0 + o.y * 16u); q01 = lds128(t0 + o.z * 16u); q10 = lds128(t0 + o.w * 16u);
        } else {
            const uint4* g0 = xb4 + cv;
            q00 = __ldg(g0 + (o.x & 0x7fffffffu)); q11 = __ldg(g0 + o.y); q01 = __ldg(g0 + o.z); q10 = __ldg(g0 + o.w);
        }
        const int pi = (int)__umulhi((unsigned)it, g.inv_row);
        out4[(long long)pi * row_stride + (it - pi * row_items)] = bilinear_vec16<T>(q00, q11, q01, q10, gw);
    }
    if (miss_counter) {
        misses = (unsigned)__reduce_add_sync(0xffffffffu, misses);
        if ((threadIdx.x & 31) == 0 && misses) atomicAdd(miss_counter, (unsigned long long)misses);
    }
}

// ---- shape-specialised variant (bf16) ------------------------------------------------------------------------------------------
// profiles/r1_ncu_gatherL1v3.txt: the kernel above is purely issue-bound at layer 1 (82.3 M warp instructions, issue slots 86 %
// busy, DRAM 58 %): its index arithmetic runs on run-time geometry (divisions by N, C/8, the tile row length; 64-bit pixel
// indices).  For the shapes of yolov8-LD-P2 -- (N, C/8, s) compile-time constants, an 8 x 16 tile, 32-bit indexing checked by
// the host -- every division becomes a shift or a constant multiply and the item rounds unroll with immediate offsets.
// Same records, same arithmetic (make_point_grid / bilinear_vec16): the operand bits are identical to the generic kernel.
template <int TN, int TCVS, int TS>
__global__ void __launch_bounds__(256)
gather_fwd_tiled2_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ x, const float* __restrict__ off,
                         const int* __restrict__ pn, __nv_bfloat16* __restrict__ operand, int H, int W, int h, int w, float hm,
                         float wm, TileGeom g)
{
    using T = __nv_bfloat16;
    constexpr int CV = 1 << TCVS, C = CV * 8, TW = 16, TH = 8;
    constexpr int SAMPLES = TH * TW * TN, ROW_ITEMS = TW * TN * CV, ITEMS = TH * ROW_ITEMS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t tile_bytes = ((uint32_t)(g.THin * g.TWin * C * 2) + 15u) & ~15u;
    const uint32_t tile_s = smem_u32(sm);
    const uint32_t rec_o = tile_s + tile_bytes, rec_g = rec_o + SAMPLES * 16u;

    const unsigned bid = blockIdx.x;
    const unsigned tpi = (unsigned)(g.tiles_w * g.tiles_h);
    const int b = (int)(bid / tpi);
    const unsigned rem = bid - (unsigned)b * tpi;
    const int ti = (int)(rem / (unsigned)g.tiles_w), tj = (int)(rem - (unsigned)ti * (unsigned)g.tiles_w);
    const int i0 = ti * TH, j0 = tj * TW;
    const int r_org = i0 * TS - g.halo, k_org = j0 * TS - g.halo;
    const int tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bar, (uint32_t)(g.THin * g.TWin * C * 2));
        tma_load_4d(sm, &tmX, &bar, 0, k_org, r_org, b);
    }

    // ---- phase 1: one record per sample; sample = pixel * N + n like the operand's row layout ---------------------------------
    const int rowB = g.TWin * (C * 2), imgRowB = W * (C * 2);
    const int m0 = (b * h + i0) * w + j0;                       // pixel index of the tile origin (host: B h w N C < 2^31)
#pragma unroll
    for (int r = 0; r < (SAMPLES + 255) / 256; ++r) {
        const int sidx = tid + r * 256;
        if (SAMPLES % 256 != 0 && sidx >= SAMPLES) break;
        const int p = sidx / TN, n = sidx - p * TN;
        const int di = p >> 4, dj = p & 15;
        if (i0 + di >= h || j0 + dj >= w) {
            sts128(rec_o + sidx * 16, 0xffffffffu, 0, 0, 0);
            continue;
        }
        const float* op = off + (size_t)(unsigned)((m0 + di * w + dj) * (2 * TN));
        const SamplePoint q = make_point_grid((i0 + di) * TS + pn[n], (j0 + dj) * TS + pn[TN + n], op[n], op[TN + n], hm, wm);
        sts128(rec_g + sidx * 16, __float_as_uint(__fmul_rn(q.ar0, q.ak0)), __float_as_uint(__fmul_rn(q.ar1, q.ak1)),
               __float_as_uint(__fmul_rn(q.ar0, q.ak1)), __float_as_uint(__fmul_rn(q.ar1, q.ak0)));
        const int t0 = q.r0 - r_org, t1 = q.r1 - r_org, u0 = q.k0 - k_org, u1 = q.k1 - k_org;
        const bool inside = (unsigned)t0 < (unsigned)g.THin && (unsigned)t1 < (unsigned)g.THin &&
                            (unsigned)u0 < (unsigned)g.TWin && (unsigned)u1 < (unsigned)g.TWin;
        if (inside) {      // byte offsets into the staged tile
            const int a0 = t0 * rowB, a1 = t1 * rowB, b0 = u0 * (C * 2), b1 = u1 * (C * 2);
            sts128(rec_o + sidx * 16, (uint32_t)(a0 + b0), (uint32_t)(a1 + b1), (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
        } else {           // served from global memory (L2): image-relative byte offsets, bit 31 of .x marks it
            const int a0 = q.r0 * imgRowB, a1 = q.r1 * imgRowB, b0 = q.k0 * (C * 2), b1 = q.k1 * (C * 2);
            sts128(rec_o + sidx * 16, (uint32_t)(a0 + b0) | 0x80000000u, (uint32_t)(a1 + b1), (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
        }
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    // ---- phase 2: one 16-byte channel vector per item; a tile row of the operand is one contiguous run of ROW_ITEMS vectors ----
    const uint8_t* xg = reinterpret_cast<const uint8_t*>(x) + (size_t)b * H * W * (C * 2);
    uint4* out4 = reinterpret_cast<uint4*>(operand) + (size_t)(unsigned)(m0 * (TN * CV));
    const int row_stride = w * (TN * CV);
#pragma unroll
    for (int r = 0; r < (ITEMS + 255) / 256; ++r) {
        const int it = tid + r * 256;
        if (ITEMS % 256 != 0 && it >= ITEMS) break;
        const int sidx = it >> TCVS, cv = it & (CV - 1);
        const uint4 o = lds128(rec_o + sidx * 16);
        if (o.x == 0xffffffffu) continue;
        const uint4 gq = lds128(rec_g + sidx * 16);
        const float4 gw = make_float4(__uint_as_float(gq.x), __uint_as_float(gq.y), __uint_as_float(gq.z), __uint_as_float(gq.w));
        uint4 q00, q11, q01, q10;
        if ((int)o.x >= 0) {
            const uint32_t t0 = tile_s + (uint32_t)cv * 16u;
            q00 = lds128(t0 + o.x); q11 = lds128(t0 + o.y); q01 = lds128(t0 + o.z); q10 = lds128(t0 + o.w);
        } else {
            const uint8_t* g0 = xg + cv * 16;
            q00 = __ldg(reinterpret_cast<const uint4*>(g0 + (o.x & 0x7fffffffu))); q11 = __ldg(reinterpret_cast<const uint4*>(g0 + o.y));
            q01 = __ldg(reinterpret_cast<const uint4*>(g0 + o.z)); q10 = __ldg(reinterpret_cast<const uint4*>(g0 + o.w));
        }
        const int pi = it / ROW_ITEMS;
        out4[pi * row_stride + (it - pi * ROW_ITEMS)] = bilinear_vec16<T>(q00, q11, q01, q10, gw);
    }
}

// ---- backward scatter on the same tiles (bf16 activations, bf16 grad_x accumulator) -----------------------------------------------
// autograd of conv.py:386-405 like gather_bwd_kernel (ldconv_core.cu), restructured like the forward kernel above: that kernel ran
// 512 instructions per (sample, 16-byte vector) item -- eight 32-bit divisions to decompose the item index, make_point per item --
// and fetched the four corner vectors of x through L1 / L2 (profiles/r2_ncu_scatterL1_acc16.txt: l1tex 76 %, 9.3 cycles of long
// scoreboard per issue).  Here the tile of x is TMA-staged, phase 1 computes one record per SAMPLE (interpolation factors, corner
// offsets in the tile and in the image), and phase 2 maps (sample, channel vector) items onto the contiguous run of grad_operand
// a tile row is: one coalesced 16-byte load of the gradient, four shared-memory corner loads, four dot products for the offset
// gradient (reduced over the channel lanes with shuffles, ONE plain store per sample and axis: every sample has exactly one owner,
// no atomics and no memset for grad_off), and one 16-byte bf16x8 reduction per distinct corner for grad_x.
__device__ __forceinline__ void red_bf16x8(uint8_t* dst, const float (&v)[8])
{
    uint32_t p0, p1, p2, p3;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(v[3]), "f"(v[2]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(v[5]), "f"(v[4]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(v[7]), "f"(v[6]));
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float (&f)[8])
{
    f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xffff0000u);
    f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xffff0000u);
    f[4] = __uint_as_float(q.z << 16); f[5] = __uint_as_float(q.z & 0xffff0000u);
    f[6] = __uint_as_float(q.w << 16); f[7] = __uint_as_float(q.w & 0xffff0000u);
}

__device__ __forceinline__ float dot_bf16x8(const float (&g)[8], const uint4& q)
{
    float f[8];
    unpack_bf16x8(q, f);
    float a = g[0] * f[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) a = fmaf(g[e], f[e], a);
    return a;
}

template <int TN, int TCVS, int TS>
__global__ void __launch_bounds__(256)
scatter_bwd_tiled2_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ x, const float* __restrict__ off,
                          const int* __restrict__ pn, const __nv_bfloat16* __restrict__ gop, __nv_bfloat16* __restrict__ grad_x,
                          float* __restrict__ grad_off, int H, int W, int h, int w, float hm, float wm, TileGeom g)
{
    constexpr int CV = 1 << TCVS, C = CV * 8, TW = 16, TH = 8;
    constexpr int SAMPLES = TH * TW * TN, ROW_ITEMS = TW * TN * CV, ITEMS = TH * ROW_ITEMS;
    static_assert(ITEMS % 256 == 0, "whole rounds of the CTA (the shuffles below need converged warps)");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t tile_bytes = ((uint32_t)(g.THin * g.TWin * C * 2) + 15u) & ~15u;
    const uint32_t tile_s = smem_u32(sm);
    // per sample: rec_o = corner byte offsets in the staged tile (.x = 0xffffffff: pixel outside the map; bit 31 of .x: a corner
    // lies outside tile + halo, read x from the image), rec_d = corner byte offsets in the image (bit 31 of .x / .y = the clamp
    // indicators in_r / in_k), rec_w = (ar0, ar1, ak0, ak1)
    const uint32_t rec_o = tile_s + tile_bytes, rec_d = rec_o + SAMPLES * 16u, rec_w = rec_d + SAMPLES * 16u;

    const unsigned bid = blockIdx.x;
    const unsigned tpi = (unsigned)(g.tiles_w * g.tiles_h);
    const int b = (int)(bid / tpi);
    const unsigned rem = bid - (unsigned)b * tpi;
    const int ti = (int)(rem / (unsigned)g.tiles_w), tj = (int)(rem - (unsigned)ti * (unsigned)g.tiles_w);
    const int i0 = ti * TH, j0 = tj * TW;
    const int r_org = i0 * TS - g.halo, k_org = j0 * TS - g.halo;
    const int tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bar, (uint32_t)(g.THin * g.TWin * C * 2));
        tma_load_4d(sm, &tmX, &bar, 0, k_org, r_org, b);
    }

    const int rowB = g.TWin * (C * 2), imgRowB = W * (C * 2);
    const int m0 = (b * h + i0) * w + j0;
#pragma unroll
    for (int r = 0; r < (SAMPLES + 255) / 256; ++r) {
        const int sidx = tid + r * 256;
        if (SAMPLES % 256 != 0 && sidx >= SAMPLES) break;
        const int p = sidx / TN, n = sidx - p * TN;
        const int di = p >> 4, dj = p & 15;
        if (i0 + di >= h || j0 + dj >= w) {
            sts128(rec_o + sidx * 16, 0xffffffffu, 0, 0, 0);
            continue;
        }
        const float* op = off + (size_t)(unsigned)((m0 + di * w + dj) * (2 * TN));
        const SamplePoint q = make_point_grid((i0 + di) * TS + pn[n], (j0 + dj) * TS + pn[TN + n], op[n], op[TN + n], hm, wm);
        sts128(rec_w + sidx * 16, __float_as_uint(q.ar0), __float_as_uint(q.ar1), __float_as_uint(q.ak0), __float_as_uint(q.ak1));
        {
            const int a0 = q.r0 * imgRowB, a1 = q.r1 * imgRowB, b0 = q.k0 * (C * 2), b1 = q.k1 * (C * 2);
            sts128(rec_d + sidx * 16, (uint32_t)(a0 + b0) | (q.in_r ? 0x80000000u : 0u), (uint32_t)(a1 + b1) | (q.in_k ? 0x80000000u : 0u),
                   (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
        }
        const int t0 = q.r0 - r_org, t1 = q.r1 - r_org, u0 = q.k0 - k_org, u1 = q.k1 - k_org;
        const bool inside = (unsigned)t0 < (unsigned)g.THin && (unsigned)t1 < (unsigned)g.THin &&
                            (unsigned)u0 < (unsigned)g.TWin && (unsigned)u1 < (unsigned)g.TWin;
        if (inside) {
            const int a0 = t0 * rowB, a1 = t1 * rowB, b0 = u0 * (C * 2), b1 = u1 * (C * 2);
            sts128(rec_o + sidx * 16, (uint32_t)(a0 + b0), (uint32_t)(a1 + b1), (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
        } else {
            sts128(rec_o + sidx * 16, 0x80000000u, 0, 0, 0);
        }
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    const uint8_t* xg = reinterpret_cast<const uint8_t*>(x) + (size_t)b * H * W * (C * 2);
    uint8_t* gxb = grad_x ? reinterpret_cast<uint8_t*>(grad_x) + (size_t)b * H * W * (C * 2) : nullptr;
    const uint4* gop4 = reinterpret_cast<const uint4*>(gop) + (size_t)(unsigned)(m0 * (TN * CV));
    const int row_stride = w * (TN * CV);
#pragma unroll 1
    for (int r = 0; r < ITEMS / 256; ++r) {
        const int it = tid + r * 256;
        const int sidx = it >> TCVS, cv = it & (CV - 1);
        const uint4 o = lds128(rec_o + sidx * 16);
        const bool valid = o.x != 0xffffffffu;
        float acc_r = 0.f, acc_k = 0.f;
        uint4 d = make_uint4(0, 0, 0, 0);
        float gv[8];
        float ar0 = 0.f, ar1 = 0.f, ak0 = 0.f, ak1 = 0.f;
        if (valid) {
            const int pi = it / ROW_ITEMS;
            const uint4 gq = __ldg(gop4 + pi * row_stride + (it - pi * ROW_ITEMS));
            d = lds128(rec_d + sidx * 16);
            const uint4 wq = lds128(rec_w + sidx * 16);
            ar0 = __uint_as_float(wq.x); ar1 = __uint_as_float(wq.y); ak0 = __uint_as_float(wq.z); ak1 = __uint_as_float(wq.w);
            uint4 q00, q11, q01, q10;
            if ((int)o.x >= 0) {
                const uint32_t t0 = tile_s + (uint32_t)cv * 16u;
                q00 = lds128(t0 + o.x); q11 = lds128(t0 + o.y); q01 = lds128(t0 + o.z); q10 = lds128(t0 + o.w);
            } else {
                const uint8_t* g0 = xg + cv * 16;
                q00 = __ldg(reinterpret_cast<const uint4*>(g0 + (d.x & 0x7fffffffu))); q11 = __ldg(reinterpret_cast<const uint4*>(g0 + (d.y & 0x7fffffffu)));
                q01 = __ldg(reinterpret_cast<const uint4*>(g0 + d.z)); q10 = __ldg(reinterpret_cast<const uint4*>(g0 + d.w));
            }
            unpack_bf16x8(gq, gv);
            const float s00 = dot_bf16x8(gv, q00), s11 = dot_bf16x8(gv, q11), s01 = dot_bf16x8(gv, q01), s10 = dot_bf16x8(gv, q10);
            acc_r = fmaf(ak0, s10 - s00, ak1 * (s11 - s01));
            acc_k = fmaf(ar0, s01 - s00, ar1 * (s11 - s10));
        }
#pragma unroll
        for (int sh = CV >> 1; sh > 0; sh >>= 1) {
            acc_r += __shfl_xor_sync(0xffffffffu, acc_r, sh);
            acc_k += __shfl_xor_sync(0xffffffffu, acc_k, sh);
        }
        if (!valid) continue;
        if (cv == 0) {
            const int p = sidx / TN, n = sidx - p * TN;
            float* gp = grad_off + (size_t)(unsigned)((m0 + (p >> 4) * w + (p & 15)) * (2 * TN));
            gp[n] = (d.x >> 31) ? acc_r : 0.f;
            gp[TN + n] = (d.y >> 31) ? acc_k : 0.f;
        }
        if (gxb == nullptr) continue;
        const uint32_t d00 = d.x & 0x7fffffffu, d11 = d.y & 0x7fffffffu, d01 = d.z, d10 = d.w;
        const float g_lt = ar0 * ak0, g_rb = ar1 * ak1, g_lb = ar0 * ak1, g_rt = ar1 * ak0;
        const bool same_r = d00 == d10, same_k = d00 == d01;
        uint8_t* base = gxb + cv * 16;
        auto corner = [&](uint32_t dofs, float wgt, bool clamped) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = gv[e] * wgt;
            if (clamped) {          // border pixels collect many samples: lanes of the warp that hit the same address add up first
                const unsigned act = __activemask();
                const unsigned peers = __match_any_sync(act, dofs + (uint32_t)cv * 16u);
                if (__any_sync(act, (peers & (peers - 1)) != 0)) {
                    const int lane = tid & 31;
                    const int leader = __ffs(peers) - 1;
                    float acc[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
                    for (unsigned mk = act; mk; mk &= mk - 1) {
                        const int src = __ffs(mk) - 1;
                        const bool take = (peers >> src) & 1u;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float t = __shfl_sync(act, v[e], src);
                            if (take) acc[e] += t;
                        }
                    }
                    if (lane != leader) return;
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = acc[e];
                }
            }
            red_bf16x8(base + dofs, v);
        };
        if (same_r && same_k) {
            corner(d00, (g_lt + g_rb) + (g_lb + g_rt), true);
        } else if (same_r) {
            corner(d00, g_lt + g_rt, true);
            corner(d01, g_lb + g_rb, true);
        } else if (same_k) {
            corner(d00, g_lt + g_lb, true);
            corner(d10, g_rt + g_rb, true);
        } else {
            corner(d00, g_lt, false);
            corner(d11, g_rb, false);
            corner(d01, g_lb, false);
            corner(d10, g_rt, false);
        }
    }
}

// returns LDCONV_OK when launched, 1 when the shape has no instance (caller: gather_bwd_kernel of ldconv_core.cu)
int scatter_bwd_tiled(const __nv_bfloat16* x, const float* off, const int* pn, const __nv_bfloat16* gop, __nv_bfloat16* grad_x,
                      float* grad_off, int B, int C, int H, int W, int N, int s, cudaStream_t st)
{
    if (C % 8 != 0 || C > 128 || N > 16 || !aligned16(x) || !aligned16(gop) || (grad_x && !aligned16(grad_x))) return 1;
    int32_t table[64];
    if (int e = ldconv_p_n(N, table)) return e;
    int max_r = 0, max_k = 0;
    for (int n = 0; n < N; ++n) {
        if (table[n] > max_r) max_r = table[n];
        if (table[N + n] > max_k) max_k = table[N + n];
    }
    const int h = out_size(H, s), w = out_size(W, s);
    TileGeom g = {};
    g.halo = 2;
    g.TH = 8;
    g.TW = 16;
    g.THin = (g.TH - 1) * s + 2 + max_r + 2 * g.halo;
    g.TWin = (g.TW - 1) * s + 2 + max_k + 2 * g.halo;
    const size_t tile = ((size_t)g.THin * g.TWin * C * 2 + 15) & ~(size_t)15;
    // larger tiles (C = 64 at stride 2: 99 KB, two CTAs per SM) measured slower than the direct kernel: 105 vs 77 us at layer 5
    if (tile > 64 * 1024 || g.THin > 256 || g.TWin > 256) return 1;
    if ((long long)B * h * w * N * C >= 0x7fffffffll || (long long)H * W * C * 2 >= 0x7fffffffll) return 1;
    g.tiles_h = (h + g.TH - 1) / g.TH;
    g.tiles_w = (w + g.TW - 1) / g.TW;
    const long long ctas = (long long)B * g.tiles_h * g.tiles_w;
    if (ctas > 0x7fffffffll) return 1;
    int cvs = -1;
    for (int sh = 0; sh < 8; ++sh)
        if ((8 << sh) == C) cvs = sh;
    using K2 = void (*)(CUtensorMap, const __nv_bfloat16*, const float*, const int*, const __nv_bfloat16*, __nv_bfloat16*, float*, int, int,
                        int, int, float, float, TileGeom);
    K2 k2 = nullptr;
    switch (N * 100 + cvs * 10 + s) {
        case 312: k2 = scatter_bwd_tiled2_kernel<3, 1, 2>; break;      // C = 16 (layer 1)
        case 322: k2 = scatter_bwd_tiled2_kernel<3, 2, 2>; break;      // C = 32 (layers 3, 18)
        case 121: k2 = scatter_bwd_tiled2_kernel<1, 2, 1>; break;      // C = 32 (layer 15)
        case 131: k2 = scatter_bwd_tiled2_kernel<1, 3, 1>; break;      // C = 64 (layers 10, 13)
        default: return 1;
    }
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)g.TWin, (cuuint32_t)g.THin, 1};
    if (int e = encode_map(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    const size_t smem = tile + 16 + (size_t)g.TH * g.TW * N * 48 + 128;
    LDC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2<<<(unsigned)ctas, 256, smem, st>>>(tm, x, off, pn, gop, grad_x, grad_off, H, W, h, w, (float)(H - 1), (float)(W - 1), g);
    LDC_LAUNCH_CHECK("scatter_bwd_tiled2_kernel");
    return LDCONV_OK;
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
    auto rec_bytes = [&](const TileGeom& t) { return (size_t)t.TH * t.TW * N * 32; };
    while (smem_of(g) > 64 * 1024 && g.TH * g.TW > 16) {
        if (g.TW > g.TH) g.TW /= 2; else g.TH /= 2;
    }
    g.THin = (g.TH - 1) * s + 2 + max_pn_r + 2 * g.halo;
    g.TWin = (g.TW - 1) * s + 2 + max_pn_k + 2 * g.halo;
    if (g.THin > 256 || g.TWin > 256 || smem_of(g) + rec_bytes(g) > 200 * 1024) return 1;   // not eligible: direct kernel
    constexpr int V = Vec16<T>::N;
    const int CV = C / V;
    g.tw_shift = 0;
    while ((1 << g.tw_shift) < g.TW) ++g.tw_shift;
    g.cv_shift = -1;
    for (int sh = 0; sh < 16; ++sh)
        if ((1 << sh) == CV) g.cv_shift = sh;
    g.inv_n = (65536u + (unsigned)N - 1) / (unsigned)N;
    const unsigned row_items = (unsigned)(g.TW * N * CV);
    g.inv_row = (unsigned)((0x100000000ull + row_items - 1) / row_items);
    // the image-relative corner offsets of phase 1 are 31-bit counts of 16-byte vectors
    if ((long long)H * W * CV >= 0x7fffffffll) return 1;
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

    const size_t smem = smem_of(g) + 16 + rec_bytes(g) + 128;
    if constexpr (sizeof(T) == 2) {      // shape-specialised instances (yolov8-LD-P2, 8 x 16 tile, no debug outputs, 32-bit indices)
        if (g.TH == 8 && g.TW == 16 && g.cv_shift >= 0 && !dbg_idx && !dbg_coord && !g_miss_counter &&
            (long long)B * h * w * N * C < 0x7fffffffll && (long long)H * W * C * 2 < 0x7fffffffll) {
            using K2 = void (*)(CUtensorMap, const __nv_bfloat16*, const float*, const int*, __nv_bfloat16*, int, int, int, int, float,
                                float, TileGeom);
            K2 k2 = nullptr;
            switch (N * 100 + g.cv_shift * 10 + s) {
                case 312: k2 = gather_fwd_tiled2_kernel<3, 1, 2>; break;      // C = 16 (layer 1)
                case 322: k2 = gather_fwd_tiled2_kernel<3, 2, 2>; break;      // C = 32 (layers 3, 18)
                case 121: k2 = gather_fwd_tiled2_kernel<1, 2, 1>; break;      // C = 32 (layer 15)
                case 131: k2 = gather_fwd_tiled2_kernel<1, 3, 1>; break;      // C = 64 (layers 10, 13)
                case 141: k2 = gather_fwd_tiled2_kernel<1, 4, 1>; break;      // C = 128 (layer 8)
                default: break;
            }
            if (k2) {
                LDC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k2<<<(unsigned)ctas, 256, smem, st>>>(tm, x, off, pn, operand, H, W, h, w, (float)(H - 1), (float)(W - 1), g);
                LDC_LAUNCH_CHECK("gather_fwd_tiled2_kernel");
                return LDCONV_OK;
            }
        }
    }
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
