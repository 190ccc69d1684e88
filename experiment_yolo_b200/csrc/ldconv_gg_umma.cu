// ldconv_gg_umma.cu -- LDConv gather + (N,1)-conv GEMM + BatchNorm + SiLU as ONE persistent tcgen05 kernel (bf16, sm_100a).
//
// Replaces /root/reference/ultralytics/nn/modules/conv.py:369-408 for inference: the sampling grid p_0 + p_n + offset
// (:413-454), floor / independent clamps / corner indices / bilinear weights (:375-393), the four gathers and the bilinear
// sum (:396-405), the 'b c h w n -> b c (h n) w' rearrange (:494-503) and Conv2d((N,1),(N,1)) + BatchNorm2d + SiLU (:355,
// :408).  The resampled (M, N*C) operand is never written to HBM: the gather warps write it straight into the K-major
// SWIZZLE_128B shared-memory tile that tcgen05.mma reads, so per call the kernel moves x + offsets + out instead of
// x + offsets + 2 x operand + out (for layer 1 of yolov8-LD-P2 at batch 64: 354 MB instead of 668 MB).
//
// Persistent CTA of 8 warps, one 128-pixel output tile (8 x 16 or 16 x 8) per pipeline step, up to three CTAs per SM.  Every
// warp is a worker (the resampling is issue-bound -- a B200 SM has ~5.5 thread instructions per byte of HBM traffic -- and a
// first version with dedicated TMA / MMA / epilogue warps spent 94 us of its 170 us at layer 1 in barrier hand-offs alone,
// profiles/r1_gg_stage_skips.txt); thread 0 issues the asynchronous work right after the block barriers:
//   phase 1   one thread per sample (pixel, n): common.cuh::make_point (bit-exact grid / indices / weights), corner
//             offsets + weights -> shared-memory records                                                  | barrier A
//   phase 2   one thread per (sample, 16-byte channel vector): four 16-byte corner loads from the TMA-staged tile (L2 when
//             a corner leaves the halo), bilinear sum, one 16-byte store into the swizzled operand tile    | barrier B
//   thread 0  TMA of the tile that reuses this input buffer; K/16 tcgen05.mma (M = 128 pixels, N = O, fp32 accumulators in
//             TMEM, double-buffered) + commit
//   all       epilogue of the PREVIOUS tile while this tile's MMA runs: tcgen05.ld -> folded BatchNorm -> SiLU -> 16-byte
//             NHWC stores (out may be a channel slice of a concat buffer)
// The offsets come from HBM (written by the tensor-core offset conv); the next tile's offsets are prefetched into registers
// before phase 2.
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

static constexpr int kGGThreads = 256;

struct GGGeom {
    int C, CV, cv_shift, N, s, H, W, h, w, O, ON, K, num_kb, ksteps;
    int TH, TW, tw_shift, THin, TWin, halo, tiles_h, tiles_w, num_tiles;
    int XB, AB, ldo, act, per_sm, dbg;
    int xb_mask, xb_shift, ab_mask;      // XB, AB in {1, 2}: buffer = it & mask, mbarrier parity = (it >> shift) & 1
    int TG;                              // v2 kernel: thread groups of 128 per CTA (3 when N % 3 == 0: one sample per thread)
    int merge;                           // v2 kernel: phase 1 of the next tile shares the barrier interval of phase 2 (two record buffers)
    int split;                           // v2 kernel, stride 2: staged tile rows are [column parity][column / 2][C] (see gg_geometry)
    float hm, wm;                        // (float)(H - 1), (float)(W - 1)
    unsigned long long img_bytes;        // H * W * C * 2
    unsigned inv_n, inv_img, inv_tw;
    uint32_t ofs_a, ofs_b, ofs_x, ofs_rec, ofs_aff, ofs_bar, x_bytes, x_tx_bytes, a_bytes, b_bytes, rec_bytes, tmem_cols;
};

__device__ __forceinline__ uint4 gg_lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 gg_lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void gg_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

template <int MINB>
__global__ void __launch_bounds__(kGGThreads, MINB)
ldconv_gg_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const __nv_bfloat16* __restrict__ x, const float* __restrict__ off, const int* __restrict__ pn,
                 const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                 const GGGeom g)
{
    using T = __nv_bfloat16;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared symbol (no integer round trip): the compiler keeps the
    // shared address space, so the C++ loads / stores below are LDS / STS that it may schedule freely between barriers
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_s = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.ofs_bar);
    uint64_t* x_full = bars;            // [2]  TMA bytes of an input tile have landed
    uint64_t* mma_done = bars + 2;      // [2]  tcgen05.commit of the tile that used TMEM buffer / operand buffer i
    uint64_t* w_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    float2* sAff = reinterpret_cast<float2*>(smem + g.ofs_aff);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = g.tiles_h * g.tiles_w;
    const int N = g.N, CV = g.CV;
    const int samples = 128 * N;
    const int items = samples * CV;

    auto tile_coords = [&](int tile, int& b, int& i0, int& j0) {      // exact magic-number divisions (host-checked range)
        b = g.inv_img ? (int)__umulhi((unsigned)tile, g.inv_img) : tile;
        const int rem = tile - b * tiles_per_img;
        const int ti = g.inv_tw ? (int)__umulhi((unsigned)rem, g.inv_tw) : rem;
        i0 = ti * g.TH;
        j0 = (rem - ti * g.tiles_w) * g.TW;
    };
    auto issue_x_tile = [&](int tile, int xb) {                         // thread 0 only
        if (g.dbg & 4) { mbar_arrive(&x_full[xb]); return; }
        int b, i0, j0;
        tile_coords(tile, b, i0, j0);
        mbar_arrive_expect_tx(&x_full[xb], g.x_tx_bytes);
        tma_load_4d(smem + g.ofs_x + (size_t)xb * g.x_bytes, &tmX, &x_full[xb], 0, j0 * g.s - g.halo, i0 * g.s - g.halo, b);
    };

    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&mma_done[i], 1); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, g.tmem_cols);
    pdl_wait();       // everything below may read what the previous kernel wrote (x, offsets, scale / shift)
    for (int o = tid; o < g.ON; o += kGGThreads)
        sAff[o] = make_float2((scale && o < g.O) ? scale[o] : 1.f, (shift && o < g.O) ? shift[o] : 0.f);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {       // weights once (resident), then the first XB input tiles of this CTA
        mbar_arrive_expect_tx(w_full, (uint32_t)g.num_kb * g.b_bytes);
        for (int kb = 0; kb < g.num_kb; ++kb) tma_load_2d(smem + g.ofs_b + (size_t)kb * g.b_bytes, &tmW, w_full, kb * 64, 0);
        for (int k = 0; k < g.XB; ++k)
            if ((int)blockIdx.x + k * (int)gridDim.x < g.num_tiles) issue_x_tile(blockIdx.x + k * gridDim.x, k);
    }

    // this thread's first two samples (rounds of 256): the (pixel, n) decomposition is tile-independent
    int my_p[2], my_n[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int sidx = tid + r * kGGThreads;
        my_p[r] = (int)(((unsigned)sidx * g.inv_n) >> 16);
        my_n[r] = sidx - my_p[r] * N;
    }
    const bool has0 = tid < samples, has1 = tid + kGGThreads < samples;
    const int pn_r0 = has0 ? pn[my_n[0]] : 0, pn_k0 = has0 ? pn[N + my_n[0]] : 0;
    const int pn_r1 = has1 ? pn[my_n[1]] : 0, pn_k1 = has1 ? pn[N + my_n[1]] : 0;
    float2 nxt[2];       // offsets (row, col) of those samples in the NEXT tile (software prefetch across phase 2)
    auto fetch_offsets = [&](int tile, float2 (&dst)[2]) {
        int b, i0, j0;
        tile_coords(tile, b, i0, j0);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            dst[r] = make_float2(0.f, 0.f);
            if (r == 0 ? has0 : has1) {
                const int i = i0 + (my_p[r] >> g.tw_shift), j = j0 + (my_p[r] & (g.TW - 1));
                if (i < g.h && j < g.w) {
                    const float* op = off + (((size_t)b * g.h + i) * g.w + j) * (size_t)(2 * N);
                    dst[r] = make_float2(__ldg(op + my_n[r]), __ldg(op + N + my_n[r]));
                }
            }
        }
    };
    if ((int)blockIdx.x < g.num_tiles) fetch_offsets(blockIdx.x, nxt);

    // epilogue of one finished tile: this warp's TMEM lane group (warp % 4) and its half of the column chunks
    const int lg = warp & 3, half = warp >> 2;
    const int ep = lg * 32 + lane;
    const int epi = ep >> g.tw_shift, epj = ep & (g.TW - 1);
    const int chunks = g.ON / 16;
    const int ch_begin = half == 0 ? 0 : (chunks + 1) / 2, ch_end = half == 0 ? (chunks + 1) / 2 : chunks;
    auto epilogue = [&](int tile, int eit) {
        const int tb = eit & 1;
        mbar_wait(&mma_done[tb], (eit >> 1) & 1);
        tc_fence_after_sync();
        int b, i0, j0;
        tile_coords(tile, b, i0, j0);
        const int i = i0 + epi, j = j0 + epj;
        const bool valid = i < g.h && j < g.w;
        T* orow = out + (((size_t)b * g.h + i) * g.w + j) * (size_t)g.ldo;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(tb * g.ON);
        for (int ch = ch_begin; ch < ch_end; ++ch) {
            const int c0 = ch * 16;
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (!valid || c0 >= g.O || (g.dbg & 8)) continue;
            const float4* aff4 = reinterpret_cast<const float4*>(sAff + c0);
            float z[16];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
                const float4 a = aff4[e >> 1];
                z[e] = fmaf(__uint_as_float(v[e]), a.x, a.y);
                z[e + 1] = fmaf(__uint_as_float(v[e + 1]), a.z, a.w);
            }
            if (g.act == LDCONV_ACT_SILU) {
#pragma unroll
                for (int e = 0; e < 16; ++e) z[e] = silu_fast(z[e]);
            } else if (g.act == LDCONV_ACT_LEAKY01) {
#pragma unroll
                for (int e = 0; e < 16; ++e) z[e] = z[e] > 0.f ? z[e] : 0.1f * z[e];
            }
            float lo[8], hi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { lo[e] = z[e]; hi[e] = z[8 + e]; }
            Vec16<T>::store(orow + c0, lo);
            Vec16<T>::store(orow + c0 + 8, hi);
        }
        tc_fence_before_sync();       // ordered before the block barrier that precedes the next MMA into this buffer
    };

    const uint32_t idesc = make_idesc_bf16(128, g.ON);
    int it = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const int xb = it % g.XB, ab = it % g.AB, tb = it & 1;
        int b, i0, j0;
        tile_coords(tile, b, i0, j0);
        const int r_org = i0 * g.s - g.halo, k_org = j0 * g.s - g.halo;
        const int r_end = r_org + g.THin, k_end = k_org + g.TWin;
        // one record buffer: it was last read in phase 2 of the previous tile, and block barrier (B) lies in between
        uint4* rec_o = reinterpret_cast<uint4*>(smem + g.ofs_rec);
        float4* rec_g = reinterpret_cast<float4*>(rec_o + samples);

        // ---- phase 1: one record per sample --------------------------------------------------------------------------------
        auto make_record = [&](int sidx, int p, int n, int pr_, int pk_, float o_r, float o_k, bool have_off) {
            const int i = i0 + (p >> g.tw_shift), j = j0 + (p & (g.TW - 1));
            if (i >= g.h || j >= g.w) {
                rec_o[sidx] = make_uint4(0xffffffffu, 0, 0, 0);
                return;
            }
            if (!have_off) {
                const float* op = off + (((size_t)b * g.h + i) * g.w + j) * (size_t)(2 * N);
                o_r = __ldg(op + n); o_k = __ldg(op + N + n);
            }
            const SamplePoint q = make_point(i, j, g.s, pr_, pk_, o_r, o_k, g.H, g.W);
            rec_g[sidx] = make_float4(__fmul_rn(q.ar0, q.ak0), __fmul_rn(q.ar1, q.ak1), __fmul_rn(q.ar0, q.ak1),
                                      __fmul_rn(q.ar1, q.ak0));
            const bool inside = q.r0 >= r_org && q.r1 < r_end && q.k0 >= k_org && q.k1 < k_end;
            if (inside) {
                const int ra = (q.r0 - r_org) * g.TWin, rb = (q.r1 - r_org) * g.TWin;
                const int ka = q.k0 - k_org, kb = q.k1 - k_org;
                rec_o[sidx] = make_uint4((uint32_t)((ra + ka) * CV), (uint32_t)((rb + kb) * CV), (uint32_t)((ra + kb) * CV),
                                         (uint32_t)((rb + ka) * CV));
            } else {      // served from global memory (L2): image-relative offsets, bit 31 of .x marks it
                const int ra = q.r0 * g.W, rb = q.r1 * g.W;
                rec_o[sidx] = make_uint4((uint32_t)((ra + q.k0) * CV) | 0x80000000u, (uint32_t)((rb + q.k1) * CV),
                                         (uint32_t)((ra + q.k1) * CV), (uint32_t)((rb + q.k0) * CV));
            }
        };
        if (!(g.dbg & 2) || it == 0) {
            if (has0) make_record(tid, my_p[0], my_n[0], pn_r0, pn_k0, nxt[0].x, nxt[0].y, true);
            if (has1) make_record(tid + kGGThreads, my_p[1], my_n[1], pn_r1, pn_k1, nxt[1].x, nxt[1].y, true);
            for (int sidx = tid + 2 * kGGThreads; sidx < samples; sidx += kGGThreads) {
                const int p = (int)(((unsigned)sidx * g.inv_n) >> 16);
                const int n = sidx - p * N;
                make_record(sidx, p, n, pn[n], pn[N + n], 0.f, 0.f, false);
            }
        }
        // prefetch the next tile's offsets; their latency hides behind phase 2
        if (tile + (int)gridDim.x < g.num_tiles) fetch_offsets(tile + gridDim.x, nxt);
        __syncthreads();                                   // (A) records of this tile are visible
        mbar_wait(&x_full[xb], (it / g.XB) & 1);           // the staged input tile has landed
        if (g.AB == 1 && it > 0) mbar_wait(&mma_done[(it - 1) & 1], ((it - 1) >> 1) & 1);   // operand buffer free again

        // ---- phase 2: bilinear resampling into the swizzled operand tile; two items per iteration, loads first ----------------
        {
            const uint4* tile4 = reinterpret_cast<const uint4*>(smem + g.ofs_x + (size_t)xb * g.x_bytes);
            const uint4* ro4 = rec_o;
            const float4* rg4 = rec_g;
            uint8_t* a_tile = smem + g.ofs_a + (size_t)ab * g.a_bytes;
            const uint4* xb4 = reinterpret_cast<const uint4*>(x + (size_t)b * g.H * g.W * g.C);
            auto corners = [&](const uint4& o, int cv, uint4& q00, uint4& q11, uint4& q01, uint4& q10) {
                if ((int)o.x >= 0) {
                    const uint4* t0 = tile4 + cv;
                    q00 = t0[o.x]; q11 = t0[o.y]; q01 = t0[o.z]; q10 = t0[o.w];
                } else {
                    const uint4* g0 = xb4 + cv;
                    q00 = __ldg(g0 + (o.x & 0x7fffffffu)); q11 = __ldg(g0 + o.y); q01 = __ldg(g0 + o.z); q10 = __ldg(g0 + o.w);
                }
            };
            auto resample = [&](const float4& gw, const uint4& q00, const uint4& q11, const uint4& q01, const uint4& q10) {
                float x00[8], x11[8], x01[8], x10[8], r[8];
                Vec16<T>::unpack(q00, x00);
                Vec16<T>::unpack(q11, x11);
                Vec16<T>::unpack(q01, x01);
                Vec16<T>::unpack(q10, x10);
#pragma unroll
                for (int v = 0; v < 8; ++v) r[v] = bilinear_fma(gw.x, gw.y, gw.z, gw.w, x00[v], x11[v], x01[v], x10[v]);
                return Vec16<T>::pack(r);
            };
            auto a_slot = [&](int sidx, int cv) {
                const uint32_t p = ((unsigned)sidx * g.inv_n) >> 16;
                const uint32_t n = (uint32_t)sidx - p * (uint32_t)N;
                const uint32_t k8 = n * (uint32_t)CV + (uint32_t)cv;            // 16-byte chunk index along K
                return reinterpret_cast<uint4*>(a_tile + (k8 >> 3) * 16384u + sw128_offset(p, k8 & 7u));
            };
            for (int ia = tid; ia < ((g.dbg & 1) ? 0 : items); ia += 2 * kGGThreads) {
                const int ib = ia + kGGThreads;
                const bool hb = ib < items;
                const int sa = g.cv_shift >= 0 ? (ia >> g.cv_shift) : ia / CV;
                const int sb = hb ? (g.cv_shift >= 0 ? (ib >> g.cv_shift) : ib / CV) : sa;
                const int cva = ia - sa * CV, cvb = hb ? ib - sb * CV : cva;
                const uint4 oa = ro4[sa], ob = ro4[sb];
                const float4 ga = rg4[sa], gb = rg4[sb];
                const bool va = oa.x != 0xffffffffu, vb = hb && ob.x != 0xffffffffu;
                uint4 a00, a11, a01, a10, b00, b11, b01, b10;
                a00 = a11 = a01 = a10 = b00 = b11 = b01 = b10 = make_uint4(0, 0, 0, 0);
                if (va) corners(oa, cva, a00, a11, a01, a10);
                if (vb) corners(ob, cvb, b00, b11, b01, b10);
                if (va) *a_slot(sa, cva) = resample(ga, a00, a11, a01, a10);
                if (vb) *a_slot(sb, cvb) = resample(gb, b00, b11, b01, b10);
            }
        }
        fence_proxy_async_smem();      // generic-proxy stores of the operand tile -> visible to tcgen05 (async proxy)
        __syncthreads();               // (B) operand tile complete, input tile consumed, epilogue(it-2) done by every warp

        if (tid == 0) {
            // the input buffer is free: stage the tile that will use it next
            if (tile + g.XB * (int)gridDim.x < g.num_tiles) issue_x_tile(tile + g.XB * gridDim.x, xb);
            if (it == 0) mbar_wait(w_full, 0);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)(tb * g.ON);
            const uint32_t a_addr = smem_s + g.ofs_a + (uint32_t)ab * g.a_bytes;
            const uint32_t b_addr = smem_s + g.ofs_b;
            for (int st = 0; st < g.ksteps; ++st) {
                const uint32_t kb = (uint32_t)st >> 2, kk = (uint32_t)st & 3;
                mma_bf16_ss(d_tmem, make_desc_k_sw128(a_addr + kb * 16384u + kk * 32u),
                            make_desc_k_sw128(b_addr + kb * g.b_bytes + kk * 32u), idesc, (uint32_t)(st != 0));
            }
            mma_commit(&mma_done[tb]);
        }
        __syncwarp();
        if (it > 0) epilogue(tile - (int)gridDim.x, it - 1);      // the previous tile's MMA ran during this tile's phases
    }
    if (it > 0) epilogue((int)blockIdx.x + (it - 1) * (int)gridDim.x, it - 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

// ---- v2 of the kernel: same pipeline, specialised index arithmetic ------------------------------------------------------------
// profiles/r1_ncu_ggL1v3.txt: the kernel above is issue-bound (114.6 M warp instructions at layer 1, 72 % of the issue slots),
// and ~40 % of those instructions are index arithmetic on run-time geometry.  This version removes them:
//   * samples are ordered n-major inside a tile (sample = n * 128 + pixel), so a thread keeps ONE pixel for all of its
//     phase-1 samples and for the epilogue, and every phase-2 round of 256 items keeps its channel vector: the pixel / n /
//     swizzled-slot decomposition is shifts and masks, and immediates once (N, C / 8, s) are template constants;
//   * tile coordinates are computed once per tile (for the offset prefetch) and carried over, buffer indices are masks,
//     records hold byte offsets, invalid edge pixels get an all-zero record instead of a branch per item;
//   * the bilinear sum and the folded BatchNorm + SiLU run on packed fp32 pairs (FFMA2): half the FMA-pipe instructions with
//     bit-identical results; the SiLU's 0.5 is folded into the affine.
// Requires C / 8 to be a power of two; other channel counts keep the kernel above.
template <int TN, int TCVS, int TS, int TG, int MINB, bool MERGE>
__global__ void __launch_bounds__(128 * TG, MINB)
ldconv_gg2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __nv_bfloat16* __restrict__ x, const float* __restrict__ off, const int* __restrict__ pn,
                  const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                  const GGGeom g)
{
    using T = __nv_bfloat16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_s = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.ofs_bar);
    uint64_t* x_full = bars;
    uint64_t* mma_done = bars + 2;
    uint64_t* w_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    float* sAff = reinterpret_cast<float*>(smem + g.ofs_aff);      // [0, ON) scale, [ON, 2 ON) shift (halved for SiLU)
    const uint32_t aff_s = smem_s + g.ofs_aff;

    constexpr int NTHR = 128 * TG;       // TG thread groups of 128: group = n phase of the samples, slice of the epilogue chunks
    constexpr int kIssuer = TG == 2 ? 128 : 0;      // the group with the least other work issues TMA / tcgen05.mma
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = TN > 0 ? TN : g.N;
    const int cvs = TCVS >= 0 ? TCVS : g.cv_shift;
    const int s = TS > 0 ? TS : g.s;
    const int tiles_per_img = g.tiles_h * g.tiles_w;
    const uint32_t rec_g_ofs = (uint32_t)(128 * N) * 16u;          // weights follow the 128 N offset records

    struct TC { int b, i0, j0; };
    auto tile_coords = [&](int tile) {
        TC t;
        t.b = g.inv_img ? (int)__umulhi((unsigned)tile, g.inv_img) : tile;
        const int rem = tile - t.b * tiles_per_img;
        const int ti = g.inv_tw ? (int)__umulhi((unsigned)rem, g.inv_tw) : rem;
        t.i0 = ti * g.TH;
        t.j0 = (rem - ti * g.tiles_w) * g.TW;
        return t;
    };
    auto issue_x_tile = [&](int tile, int xb) {                         // one thread
        const TC t = tile_coords(tile);
        mbar_arrive_expect_tx(&x_full[xb], g.x_tx_bytes);
        if (g.split)      // W viewed as (W/2, 2): coordinates (channel, column / 2, parity, row, image); the origin column is even
            tma_load_5d(smem + g.ofs_x + (size_t)xb * g.x_bytes, &tmX, &x_full[xb], 0, (t.j0 * s - g.halo) >> 1, 0,
                        t.i0 * s - g.halo, t.b);
        else
            tma_load_4d(smem + g.ofs_x + (size_t)xb * g.x_bytes, &tmX, &x_full[xb], 0, t.j0 * s - g.halo, t.i0 * s - g.halo, t.b);
    };

    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&mma_done[i], 1); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, g.tmem_cols);
    pdl_wait();       // everything below may read what the previous kernel wrote (x, offsets, scale / shift)
    const bool act_silu = g.act == LDCONV_ACT_SILU;
    {
        const float pre = act_silu ? 0.5f : 1.f;      // silu(z) = hz + hz tanh(hz), hz = z / 2: the halving is exact
        for (int o = tid; o < g.ON; o += NTHR) {
            sAff[o] = pre * ((scale && o < g.O) ? scale[o] : 1.f);
            sAff[g.ON + o] = pre * ((shift && o < g.O) ? shift[o] : 0.f);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(w_full, (uint32_t)g.num_kb * g.b_bytes);
        for (int kb = 0; kb < g.num_kb; ++kb) tma_load_2d(smem + g.ofs_b + (size_t)kb * g.b_bytes, &tmW, w_full, kb * 64, 0);
        for (int k = 0; k < g.XB; ++k)
            if ((int)blockIdx.x + k * (int)gridDim.x < g.num_tiles) issue_x_tile(blockIdx.x + k * gridDim.x, k);
    }

    // ---- per-thread constants: one pixel of the tile for phase 1 and the epilogue, one channel vector for phase 2 ------------
    const int p = tid & 127, n0 = tid >> 7;
    const int di = p >> g.tw_shift, dj = p & (g.TW - 1);
    const int pix_f = (di * g.w + dj) * 2 * N;                 // this pixel's offsets, in floats from the tile's origin pixel
    bool has[2];
    int br[2], bk[2];                                          // di * s + pn_r[n], dj * s + pn_k[n] of sample rounds 0, 1
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = n0 + TG * r;
        has[r] = (TN > 0 && TN <= TG * r) ? false : n < N;
        br[r] = di * s + (has[r] ? pn[n] : 0);
        bk[r] = dj * s + (has[r] ? pn[N + n] : 0);
    }
    float2 ofs[2], ofs_n[2];                                   // offsets of those samples: this tile / prefetched for the next
    auto fetch_offsets = [&](const TC& t, float2 (&dst)[2]) {
        const bool valid = t.i0 + di < g.h && t.j0 + dj < g.w;
        const float* op = off + (((size_t)t.b * g.h + t.i0) * g.w + t.j0) * (size_t)(2 * N) + pix_f + n0;
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (valid && has[r]) dst[r] = make_float2(__ldg(op + TG * r), __ldg(op + N + TG * r));
    };
    ofs[0] = ofs[1] = ofs_n[0] = ofs_n[1] = make_float2(0.f, 0.f);

    const int cv = tid & ((1 << cvs) - 1), sx0 = tid >> cvs;   // phase 2: item round k handles sample sx0 + k * (NTHR >> cvs)
    const int spr = NTHR >> cvs;
    const int rounds = (N << cvs) / TG;                        // 128 N CV items / NTHR threads (the host checks divisibility)
    const int grp = warp >> 2;
    const int chunks = g.ON / 16;
    const int ch_begin = chunks * grp / TG, ch_end = chunks * (grp + 1) / TG;

    // epilogue of one finished tile: TMEM lane = this thread's pixel, this warp's half of the 16-column chunks
    auto epilogue = [&](int m_out, int eit) {      // m_out: output pixel index ((b h + i) w + j) of this thread, -1 outside the map
        const int tb = eit & 1;
        mbar_wait(&mma_done[tb], (eit >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tb * g.ON);
        for (int ch = ch_begin; ch < ch_end; ++ch) {
            const int c0 = ch * 16;
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (m_out < 0 || c0 >= g.O) continue;
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                const float4 sc = gg_lds_f4(aff_s + (uint32_t)(c0 + e) * 4u);
                const float4 sh = gg_lds_f4(aff_s + (uint32_t)(g.ON + c0 + e) * 4u);
                uint64_t z0 = f2_fma(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(sc.x, sc.y),
                                     f2_pack(sh.x, sh.y));
                uint64_t z1 = f2_fma(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(sc.z, sc.w),
                                     f2_pack(sh.z, sh.w));
                float a0, a1, a2, a3;
                f2_unpack(z0, a0, a1);
                f2_unpack(z1, a2, a3);
                if (act_silu) {
                    float t0, t1, t2, t3;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(a2));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(a3));
                    z0 = f2_fma(z0, f2_pack(t0, t1), z0);
                    z1 = f2_fma(z1, f2_pack(t2, t3), z1);
                    f2_unpack(z0, a0, a1);
                    f2_unpack(z1, a2, a3);
                } else if (g.act == LDCONV_ACT_LEAKY01) {
                    a0 = a0 > 0.f ? a0 : 0.1f * a0; a1 = a1 > 0.f ? a1 : 0.1f * a1;
                    a2 = a2 > 0.f ? a2 : 0.1f * a2; a3 = a3 > 0.f ? a3 : 0.1f * a3;
                }
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[e >> 1]) : "f"(a1), "f"(a0));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[(e >> 1) + 1]) : "f"(a3), "f"(a2));
            }
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)m_out * (size_t)g.ldo + c0);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        tc_fence_before_sync();       // ordered before the block barrier that precedes the next MMA into this buffer
    };

    const uint32_t idesc = make_idesc_bf16(128, g.ON);
    const int rowB = g.TWin << (cvs + 4), pixB = 16 << cvs;    // bytes per staged tile row / per pixel
    const int imgRowB = g.W << (cvs + 4);
    const int halfRowB = rowB >> 1;
    // ---- phase 1 of one tile: one record per sample (n-major: sample = n * 128 + pixel) into the record buffer at rec_s; uses the
    // prefetched offsets in `ofs`; returns this thread's output pixel index ((b h + i) w + j), -1 outside the map
    auto phase1 = [&](const TC& t, uint32_t rec_s) -> int {
        const int r_org = t.i0 * s - g.halo, k_org = t.j0 * s - g.halo;
        const bool valid = t.i0 + di < g.h && t.j0 + dj < g.w;
        auto make_record = [&](int n, int ri, int ki, float o_r, float o_k) {
            const uint32_t ra = rec_s + (uint32_t)(n * 128 + p) * 16u;
            if (!valid) {
                gg_sts128(ra, 0u, 0u, 0u, 0u);
                gg_sts128(ra + rec_g_ofs, 0u, 0u, 0u, 0u);
                return;
            }
            const SamplePoint q = make_point_grid(ri, ki, o_r, o_k, g.hm, g.wm);
            gg_sts128(ra + rec_g_ofs, __float_as_uint(__fmul_rn(q.ar0, q.ak0)), __float_as_uint(__fmul_rn(q.ar1, q.ak1)),
                      __float_as_uint(__fmul_rn(q.ar0, q.ak1)), __float_as_uint(__fmul_rn(q.ar1, q.ak0)));
            const int t0 = q.r0 - r_org, t1 = q.r1 - r_org, u0 = q.k0 - k_org, u1 = q.k1 - k_org;
            const bool inside = (unsigned)t0 < (unsigned)g.THin && (unsigned)t1 < (unsigned)g.THin &&
                                (unsigned)u0 < (unsigned)g.TWin && (unsigned)u1 < (unsigned)g.TWin;
            if (inside) {
                const int a0 = t0 * rowB, a1 = t1 * rowB;
                int b0 = u0 * pixB, b1 = u1 * pixB;
                if (g.split) {      // [parity][column / 2][C] rows
                    b0 = (u0 & 1) * halfRowB + (u0 >> 1) * pixB;
                    b1 = (u1 & 1) * halfRowB + (u1 >> 1) * pixB;
                }
                gg_sts128(ra, (uint32_t)(a0 + b0), (uint32_t)(a1 + b1), (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
            } else {      // served from global memory (L2): image-relative byte offsets, bit 31 of .x marks it
                const int a0 = q.r0 * imgRowB, a1 = q.r1 * imgRowB, b0 = q.k0 * pixB, b1 = q.k1 * pixB;
                gg_sts128(ra, (uint32_t)(a0 + b0) | 0x80000000u, (uint32_t)(a1 + b1), (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
            }
        };
        const int gr = t.i0 * s, gk = t.j0 * s;
        if (has[0]) make_record(n0, gr + br[0], gk + bk[0], ofs[0].x, ofs[0].y);
        if (has[1]) make_record(n0 + TG, gr + br[1], gk + bk[1], ofs[1].x, ofs[1].y);
        if (TN == 0 || TN > 2 * TG) {
            for (int n = n0 + 2 * TG; n < N; n += TG) {
                float o_r = 0.f, o_k = 0.f;
                if (valid) {
                    const float* op = off + (((size_t)t.b * g.h + t.i0) * g.w + t.j0) * (size_t)(2 * N) + pix_f;
                    o_r = __ldg(op + n); o_k = __ldg(op + N + n);
                }
                make_record(n, gr + di * s + pn[n], gk + dj * s + pn[N + n], o_r, o_k);
            }
        }
        return valid ? ((t.b * g.h + t.i0 + di) * g.w + t.j0 + dj) : -1;
    };

    TC cur = tile_coords(blockIdx.x);
    if ((int)blockIdx.x < g.num_tiles) fetch_offsets(cur, ofs);
    int prev_m = -1, cur_m = -1;
    int it = 0;
    if (MERGE && (int)blockIdx.x < g.num_tiles) {      // one-barrier flow: the records of the first tile are made ahead of the loop
        cur_m = phase1(cur, smem_s + g.ofs_rec);
        if ((int)(blockIdx.x + gridDim.x) < g.num_tiles) fetch_offsets(tile_coords(blockIdx.x + gridDim.x), ofs);
        __syncthreads();
    }
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const int xb = it & g.xb_mask, ab = it & g.ab_mask, tb = it & 1;
        // MERGE: two record buffers, this tile's were written during the previous tile's interval.  Otherwise one buffer, last
        // read in phase 2 of the previous tile with barrier (B) in between.
        const uint32_t rec_s = smem_s + g.ofs_rec + (MERGE ? (uint32_t)(it & 1) * g.rec_bytes : 0u);
        const bool more = tile + (int)gridDim.x < g.num_tiles;
        if (!MERGE) {
            // next tile: its offsets are requested now and have both phases to arrive
            if (more) fetch_offsets(tile_coords(tile + gridDim.x), ofs_n);
            cur_m = phase1(cur, rec_s);
            __syncthreads();                                          // (A) records of this tile are visible
        }
        mbar_wait(&x_full[xb], (uint32_t)(it >> g.xb_shift) & 1u);   // the staged input tile has landed
        if (g.ab_mask == 0 && it > 0) mbar_wait(&mma_done[(it - 1) & 1], ((it - 1) >> 1) & 1);   // operand buffer free again

        // ---- phase 2: bilinear resampling into the swizzled operand tile; two items per step, loads first --------------------
        {
            const uint32_t x_s = smem_s + g.ofs_x + (uint32_t)xb * g.x_bytes + ((uint32_t)cv << 4);
            const uint32_t a_s = smem_s + g.ofs_a + (uint32_t)ab * g.a_bytes;
            const uint8_t* xg = reinterpret_cast<const uint8_t*>(x) + (size_t)cur.b * g.img_bytes + ((uint32_t)cv << 4);
            // the offset record of an item is fetched one step ahead of its corner loads (load_rec), so that a step's dependent
            // chain is corner loads -> arithmetic instead of record -> corner loads -> arithmetic
            auto load_rec = [&](int sx) { return gg_lds128(rec_s + (uint32_t)sx * 16u); };
            auto load_item = [&](int sx, const uint4& o, float4& gw, uint4 (&q)[4]) {
                gw = gg_lds_f4(rec_s + rec_g_ofs + (uint32_t)sx * 16u);
                if ((int)o.x >= 0) {
                    q[0] = gg_lds128(x_s + o.x); q[1] = gg_lds128(x_s + o.y); q[2] = gg_lds128(x_s + o.z); q[3] = gg_lds128(x_s + o.w);
                } else {
                    // cold path: the sample left the staged halo, its corners come from global memory (L2).  Written as a
                    // non-unrolled rotate loop so the compiler keeps it a branch instead of if-converting ~14 predicated
                    // instructions into the hot path of every item.
                    uint4 oo = make_uint4(o.x & 0x7fffffffu, o.y, o.z, o.w);
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        q[0] = q[1]; q[1] = q[2]; q[2] = q[3];
                        q[3] = __ldg(reinterpret_cast<const uint4*>(xg + oo.x));
                        oo = make_uint4(oo.y, oo.z, oo.w, oo.x);
                    }
                }
            };
            auto store_item = [&](int sx, const float4& gw, const uint4 (&q)[4]) {
                const uint32_t px = (uint32_t)sx & 127u, n = (uint32_t)sx >> 7;
                const uint32_t k8 = (n << cvs) + (uint32_t)cv;            // 16-byte chunk index along K
                gg_sts128(a_s + (k8 >> 3) * 16384u + px * 128u + (((k8 ^ px) & 7u) << 4),
                          bilinear_bf16x2(q[0].x, q[1].x, q[2].x, q[3].x, gw), bilinear_bf16x2(q[0].y, q[1].y, q[2].y, q[3].y, gw),
                          bilinear_bf16x2(q[0].z, q[1].z, q[2].z, q[3].z, gw), bilinear_bf16x2(q[0].w, q[1].w, q[2].w, q[3].w, gw));
            };
            // IF items in flight per thread (loads first, then the arithmetic): two under an 80-register cap, four when the
            // shared-memory footprint allows so few CTAs that 120+ registers are free anyway (latency-bound there)
            constexpr int IF = (65536 / (NTHR * MINB) >= 120 && TN != 0) ? 4 : 2;
            uint4 orec[IF];      // offset records of the step about to run
            auto fetch_recs = [&](int k, int nrounds) {
#pragma unroll
                for (int u = 0; u < IF; ++u)
                    if (k + u < nrounds) orec[u] = load_rec(sx0 + (k + u) * spr);
            };
            auto step = [&](int k, int nrounds) {
                float4 gw[IF];
                uint4 q[IF][4];
#pragma unroll
                for (int u = 0; u < IF; ++u)
                    if (k + u < nrounds) load_item(sx0 + (k + u) * spr, orec[u], gw[u], q[u]);
                fetch_recs(k + IF, nrounds);      // next step's records, in flight during this step's arithmetic
#pragma unroll
                for (int u = 0; u < IF; ++u)
                    if (k + u < nrounds) store_item(sx0 + (k + u) * spr, gw[u], q[u]);
            };
            if constexpr (TN > 0 && TCVS >= 0) {
                constexpr int R = (TN << TCVS) / TG;
                static_assert((TN << TCVS) % TG == 0, "items per tile must divide evenly over the threads");
                fetch_recs(0, R);
#pragma unroll
                for (int k = 0; k < R; k += IF) step(k, R);
            } else {
                fetch_recs(0, rounds);
                for (int k = 0; k < rounds; k += IF) step(k, rounds);
            }
        }
        int next_m = -1;
        TC nxt = cur;
        if (MERGE && more) {
            // phase 1 of the NEXT tile in the same barrier interval (its ALU work fills the load latencies of phase 2, and the
            // tile needs one block barrier instead of two); then request the offsets of the tile after it
            nxt = tile_coords(tile + gridDim.x);
            next_m = phase1(nxt, smem_s + g.ofs_rec + (uint32_t)((it + 1) & 1) * g.rec_bytes);
            if (tile + 2 * (int)gridDim.x < g.num_tiles) fetch_offsets(tile_coords(tile + 2 * gridDim.x), ofs);
        }
        fence_proxy_async_smem();      // generic-proxy stores of the operand tile -> visible to tcgen05 (async proxy)
        __syncthreads();               // (B) operand tile complete, input tile consumed, epilogue(it-2) done by every warp

        if (tid == kIssuer) {
            if (tile + g.XB * (int)gridDim.x < g.num_tiles) issue_x_tile(tile + g.XB * gridDim.x, xb);
            if (it == 0) mbar_wait(w_full, 0);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)(tb * g.ON);
            const uint64_t da = make_desc_k_sw128(smem_s + g.ofs_a + (uint32_t)ab * g.a_bytes);
            const uint64_t desc_b0 = make_desc_k_sw128(smem_s + g.ofs_b);      // + (bytes >> 4): address field, bits [0, 14), no carry
            for (int st = 0; st < g.ksteps; ++st) {
                const uint32_t kb = (uint32_t)st >> 2, kk = (uint32_t)st & 3;
                mma_bf16_ss(d_tmem, da + (uint64_t)(kb * 1024u + kk * 2u), desc_b0 + (uint64_t)(kb * (g.b_bytes >> 4) + kk * 2u), idesc,
                            (uint32_t)(st != 0));
            }
            mma_commit(&mma_done[tb]);
        }
        __syncwarp();
        if (it > 0) epilogue(prev_m, it - 1);      // the previous tile's MMA ran during this tile's phases
        prev_m = cur_m;
        if (MERGE) {
            cur_m = next_m;
            cur = nxt;
        } else {
            if (more) cur = tile_coords(tile + gridDim.x);      // recomputed: cheaper than 3 live registers
            ofs[0] = ofs_n[0]; ofs[1] = ofs_n[1];
        }
    }
    if (it > 0) epilogue(prev_m, it - 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

static constexpr int gg_v2_enabled() { return 2; }      // v2 kernel with the shape-specialised instances (settled in round 1)

static void gg_pn_extent(int N, int* max_r, int* max_k)
{
    int32_t table[64];
    *max_r = *max_k = 0;
    if (N > 16 || ldconv_p_n(N, table) != LDCONV_OK) return;
    for (int n = 0; n < N; ++n) {
        if (table[n] > *max_r) *max_r = table[n];
        if (table[N + n] > *max_k) *max_k = table[N + n];
    }
}

// Fills the geometry and the dynamic shared-memory size; returns 0 when the shape is not covered.
static int gg_geometry(int B, int C, int H, int W, int N, int s, int O, int ldo, int act, GGGeom* out, size_t* smem_bytes)
{
    if (C % 8 != 0 || C > 256 || N < 1 || N > 16 || O % 16 != 0 || O > 256 || ldo % 8 != 0 || ldo < O) return 0;
    const int K = N * C;
    if (K % 16 != 0 || K > 1024) return 0;
    GGGeom g;
    g.C = C; g.CV = C / 8; g.N = N; g.s = s; g.H = H; g.W = W; g.O = O; g.K = K; g.ldo = ldo; g.act = act;
    g.h = out_size(H, s); g.w = out_size(W, s);
    g.ON = (O + 15) / 16 * 16;
    g.num_kb = (K + 63) / 64;
    g.ksteps = K / 16;
    g.cv_shift = -1;
    for (int sh = 0; sh < 8; ++sh)
        if ((1 << sh) == g.CV) g.cv_shift = sh;
    g.inv_n = (65536u + (unsigned)N - 1) / (unsigned)N;
    if ((long long)H * W * C * 2 >= 0x7fffffffll) return 0;
    g.hm = (float)(H - 1); g.wm = (float)(W - 1);
    g.img_bytes = (unsigned long long)H * W * C * 2;
    int mr, mk;
    gg_pn_extent(N, &mr, &mk);
    auto waste = [&](int th, int tw) {
        return (long long)((g.h + th - 1) / th * th) * ((g.w + tw - 1) / tw * tw) - (long long)g.h * g.w;
    };
    g.TH = 8; g.TW = 16;
    if (waste(16, 8) < waste(8, 16)) { g.TH = 16; g.TW = 8; }
    g.tw_shift = g.TW == 16 ? 4 : 3;
    g.b_bytes = (uint32_t)g.ON * 128u;
    g.a_bytes = (uint32_t)g.num_kb * 16384u;
    g.rec_bytes = (uint32_t)(128 * N * 32);
    g.tmem_cols = 32;
    while (g.tmem_cols < (uint32_t)(2 * g.ON)) g.tmem_cols <<= 1;
    // candidate plans (halo, input-tile buffers, operand buffers): pick the one with the most CTAs per SM (<= 3: registers),
    // ties go to the deeper buffering / wider halo (listed first)
    static const int cfg[5][3] = {{2, 2, 2}, {2, 2, 1}, {2, 1, 1}, {1, 2, 1}, {1, 1, 1}};
    int best_ctas = 0;
    g.merge = 0;
    g.split = 0;
    GGGeom best = g;
    size_t best_smem = 0;
    // N = 3 (every strided LDConv of yolov8-LD-P2): 384 samples per tile -> 384 threads make phase 1 one balanced round
    // (profiles/r1_ncu_ggL1v4.txt: 17 % of the stall samples were warps 4-7 waiting at barrier A for the second round of warps 0-3)
    // measured (B = 64, bf16): C = 32 / 64 gain 3-7 % with 384 threads (shared memory limits them to 1-2 CTAs per SM, so the extra
    // warps are free); C = 16 (layer 1) loses 9 % (two CTAs of 12 warps instead of three of 8) and keeps 256 threads
    g.TG = (gg_v2_enabled() && g.cv_shift >= 2 && N % 3 == 0) ? 3 : 2;
    const int reg_cap = g.TG == 3 ? 2 : 3;      // CTAs per SM the register file allows at 80 registers per thread
    const bool split_ok = gg_v2_enabled() && g.cv_shift >= 0 && s == 2 && W % 2 == 0 && C <= 32;
    // shared-memory layout of plan ci with (mg + 1) record buffers -> CTAs per SM (0: does not fit)
    auto layout = [&](int ci, int mg, GGGeom& q, size_t& need) {
        q.halo = cfg[ci][0]; q.XB = cfg[ci][1]; q.AB = cfg[ci][2];
        q.THin = (q.TH - 1) * s + 2 + mr + 2 * q.halo;
        q.TWin = (q.TW - 1) * s + 2 + mk + 2 * q.halo;
        // Stride 2, pixels of 32 / 64 bytes: the corners of neighbouring output pixels are two input pixels apart, so the four (two)
        // samples of a quarter-warp's 16-byte LDS sat on two (one) of the four 32-byte bank groups: 41 % of the kernel's shared-memory
        // wavefronts were bank conflicts (profiles/r1_ncu_ggL1_s4.txt).  With the columns of a staged row split by parity
        // ([parity][column / 2][C], a 5-D TMA map whose W axis is viewed as (W/2, 2)) those corners are neighbours in shared memory.
        q.split = (split_ok && q.halo % 2 == 0) ? 1 : 0;
        if (q.split) q.TWin = (q.TWin + 1) & ~1;
        if (q.THin > 256 || q.TWin > 256) return 0;
        q.x_tx_bytes = (uint32_t)((size_t)q.THin * q.TWin * C * 2);
        q.x_bytes = (q.x_tx_bytes + 127u) & ~127u;
        uint32_t ofs = 0;
        q.ofs_a = ofs; ofs += (uint32_t)q.AB * q.a_bytes;
        q.ofs_b = ofs; ofs += (uint32_t)q.num_kb * q.b_bytes;
        q.ofs_x = ofs; ofs += (uint32_t)q.XB * q.x_bytes;
        q.ofs_rec = ofs; ofs += (uint32_t)(mg + 1) * q.rec_bytes;
        q.ofs_aff = ofs; ofs += (uint32_t)q.ON * 8u;
        ofs = (ofs + 7u) & ~7u;
        q.ofs_bar = ofs; ofs += 6u * 8u + 16u;
        need = (size_t)ofs + 1024;
        if (need > 225 * 1024) return 0;
        int ctas = (int)((227 * 1024) / (need + 1024));          // + the per-CTA reservation of the driver
        if (ctas > reg_cap) ctas = reg_cap;
        if (ctas > (int)(512u / q.tmem_cols)) ctas = (int)(512u / q.tmem_cols);
        q.merge = mg;
        return ctas;
    };
    int best_ci = -1;
    for (int ci = 0; ci < 5; ++ci) {
        size_t need = 0;
        const int ctas = layout(ci, 0, g, need);
        if (ctas > best_ctas) { best_ctas = ctas; best = g; best_smem = need; best_ci = ci; }
    }
    // One-barrier flow (phase 1 of the next tile shares the barrier interval of phase 2; two record buffers): measured to pay
    // for N = 1 (layer 15: 95 -> 86 us; phase 1 occupies half the warps there) and not for N = 3 (layer 1: 127 -> 134 us), and
    // only taken when the SAME buffering plan keeps its CTAs per SM with the second record buffer.
    if (best_ci >= 0 && gg_v2_enabled() && g.cv_shift >= 0 && N == 1) {
        size_t need = 0;
        GGGeom q = g;
        if (layout(best_ci, 1, q, need) == best_ctas) { best = q; best_smem = need; }
    }
    if (best_ctas == 0) return 0;
    g = best;
    g.per_sm = best_ctas;
    g.xb_mask = g.XB - 1; g.xb_shift = g.XB - 1; g.ab_mask = g.AB - 1;
    g.tiles_h = (g.h + g.TH - 1) / g.TH;
    g.tiles_w = (g.w + g.TW - 1) / g.TW;
    const long long nt = (long long)B * g.tiles_h * g.tiles_w;
    if (nt > 0x7fffffffll || (long long)B * g.h * g.w > 0x7fffffffll) return 0;
    g.num_tiles = (int)nt;
    g.dbg = 0;
    // tile / tiles_per_img and rem / tiles_w as __umulhi(x, ceil(2^32 / d)): exact while x * d < 2^32
    const unsigned tpi = (unsigned)(g.tiles_h * g.tiles_w);
    if (nt * tpi >= 0xffffffffll) return 0;
    g.inv_img = tpi == 1 ? 0u : (unsigned)((0x100000000ull + tpi - 1) / tpi);
    g.inv_tw = g.tiles_w == 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)g.tiles_w - 1) / (unsigned)g.tiles_w);
    *smem_bytes = best_smem;
    *out = g;
    return 1;
}


int gather_gemm_supported(int B, int C, int H, int W, int N, int s, int O, int ldo, int dtype)
{
    if (dtype != LDCONV_BF16) return 0;
    GGGeom g;
    size_t smem;
    return gg_geometry(B, C, H, W, N, s, O, ldo, LDCONV_ACT_SILU, &g, &smem);
}

int gather_gemm_fwd(const void* x, const float* off, const int* pn, const void* wt, const float* scale, const float* shift,
                    void* out, int ldo, int B, int C, int H, int W, int N, int s, int O, int act, cudaStream_t st)
{
    GGGeom g;
    size_t smem;
    if (!gg_geometry(B, C, H, W, N, s, O, ldo, act, &g, &smem))
        return fail(LDCONV_E_ARG, "gather+GEMM kernel: shape not covered (C=%d N=%d s=%d O=%d ldo=%d)", C, N, s, O, ldo);
    if (!aligned16(x) || !aligned16(wt) || !aligned16(out))
        return fail(LDCONV_E_ALIGN, "gather+GEMM kernel: x / wt / out must be 16-byte aligned");
    CUtensorMap tmX, tmW;
    if (g.split) {      // (C, W/2, 2, H, B): the box lands in shared memory as [row][column parity][column / 2][C]
        cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)(W / 2), 2, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[4] = {(cuuint64_t)2 * C * 2, (cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[5] = {(cuuint32_t)C, (cuuint32_t)(g.TWin / 2), 2, (cuuint32_t)g.THin, 1};
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    } else {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)g.TWin, (cuuint32_t)g.THin, 1};
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)g.K, (cuuint64_t)O};
        cuuint64_t gstr[1] = {(cuuint64_t)g.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)g.ON};
        if (int e = encode_map(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    using Kern = void (*)(CUtensorMap, CUtensorMap, const __nv_bfloat16*, const float*, const int*, const float*, const float*,
                          __nv_bfloat16*, GGGeom);
    Kern kern = ldconv_gg_kernel<3>;      // <= 84 registers, no spills; a 64-register build for four CTAs per SM measured equal
    int threads = kGGThreads;
    const int env_v = gg_v2_enabled();
    if (env_v && g.cv_shift >= 0) {
        threads = 128 * g.TG;
        // the register cap follows the CTAs per SM the shared-memory plan allows (80 / 128 / 255 registers at 256 threads)
        const int key0 = (g.TG * 10 + g.per_sm) * 10 + g.merge;
        switch (key0) {      // run-time geometry: the register cap follows the CTAs per SM (80 / 128 / 255 registers at 256 threads)
            case 320: kern = ldconv_gg2_kernel<0, -1, 0, 3, 2, false>; break;
            case 321: kern = ldconv_gg2_kernel<0, -1, 0, 3, 2, true>; break;
            case 310: kern = ldconv_gg2_kernel<0, -1, 0, 3, 1, false>; break;
            case 311: kern = ldconv_gg2_kernel<0, -1, 0, 3, 1, true>; break;
            case 230: kern = ldconv_gg2_kernel<0, -1, 0, 2, 3, false>; break;
            case 231: kern = ldconv_gg2_kernel<0, -1, 0, 2, 3, true>; break;
            case 220: kern = ldconv_gg2_kernel<0, -1, 0, 2, 2, false>; break;
            case 221: kern = ldconv_gg2_kernel<0, -1, 0, 2, 2, true>; break;
            case 210: kern = ldconv_gg2_kernel<0, -1, 0, 2, 1, false>; break;
            default: kern = ldconv_gg2_kernel<0, -1, 0, 2, 1, true>; break;
        }
        const int key = ((((N * 10 + g.cv_shift) * 10 + s) * 10 + g.TG) * 10 + g.per_sm) * 10 + g.merge;      // yolov8-LD-P2 shapes
        if (env_v == 2) switch (key) {
            case 312231: kern = ldconv_gg2_kernel<3, 1, 2, 2, 3, true>; break;       // C = 16 (layer 1)
            case 312230: kern = ldconv_gg2_kernel<3, 1, 2, 2, 3, false>; break;
            case 322320: kern = ldconv_gg2_kernel<3, 2, 2, 3, 2, false>; break;      // C = 32 (layers 3, 18)
            case 322310: kern = ldconv_gg2_kernel<3, 2, 2, 3, 1, false>; break;
            case 332311: kern = ldconv_gg2_kernel<3, 3, 2, 3, 1, true>; break;       // C = 64 (layers 5, 21)
            case 332310: kern = ldconv_gg2_kernel<3, 3, 2, 3, 1, false>; break;
            case 121231: kern = ldconv_gg2_kernel<1, 2, 1, 2, 3, true>; break;       // C = 32 (layer 15)
            case 121230: kern = ldconv_gg2_kernel<1, 2, 1, 2, 3, false>; break;
            case 131231: kern = ldconv_gg2_kernel<1, 3, 1, 2, 3, true>; break;       // C = 64 (layers 10, 13)
            case 131230: kern = ldconv_gg2_kernel<1, 3, 1, 2, 3, false>; break;
            case 141221: kern = ldconv_gg2_kernel<1, 4, 1, 2, 2, true>; break;       // C = 128 (layer 8)
            case 141220: kern = ldconv_gg2_kernel<1, 4, 1, 2, 2, false>; break;
            default: break;
        }
    }
    LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = num_sms() * g.per_sm;
    if (grid > g.num_tiles) grid = g.num_tiles;
    LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(threads), smem, st, tmX, tmW, (const __nv_bfloat16*)x, off, pn, scale, shift,
                        (__nv_bfloat16*)out, g));
    LDC_LAUNCH_CHECK("ldconv_gg_kernel");
    set_impl(LDCONV_IMPL_TCGEN05);
    return LDCONV_OK;
}

}  // namespace ldc

// Inference forward of everything after the offset conv (conv.py:369-408) in one kernel; `out` (B,h,w,O | ldo) may be a
// channel slice of a wider NHWC buffer (ldo = its pixel stride in elements).
LDC_API int ldconv_gather_gemm_supported(int B, int C, int H, int W, int N, int s, int O, int ldo, int dtype)
{
    return ldc::gather_gemm_supported(B, C, H, W, N, s, O, ldo, dtype);
}

LDC_API int ldconv_gather_gemm_fwd(const void* x, const float* off, const int32_t* p_n, const void* wt, const float* scale,
                                   const float* shift, void* out, int ldo, int B, int C, int H, int W, int N, int s, int O,
                                   int act, int dtype, void* stream)
{
    using namespace ldc;
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_gather_gemm_fwd: bf16 only");
    LDC_REQUIRE(x && off && p_n && wt && out, "ldconv_gather_gemm_fwd: null pointer");
    LDC_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1 && N >= 1 && s >= 1 && O >= 1, "ldconv_gather_gemm_fwd: bad dims");
    if (B == 0) return LDCONV_OK;
    return gather_gemm_fwd(x, off, p_n, wt, scale, shift, out, ldo, B, C, H, W, N, s, O, act, (cudaStream_t)stream);
}
