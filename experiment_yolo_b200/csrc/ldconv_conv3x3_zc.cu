// ldconv_conv3x3_zc.cu -- stride-1 3x3 / pad 1 convolution as a ZERO-COPY tcgen05 implicit GEMM (bf16, NHWC, sm_100a).
//
// Same job as ldconv_conv3x3_umma.cu (LDConv's offset conv, /root/reference/ultralytics/nn/modules/conv.py:356,368, for
// its stride-1 layers, and the Conv2d+BatchNorm2d+SiLU 3x3 blocks around LDConv, nn/modules/conv.py:41-59), without the
// shared->shared im2col copy, whose shared-memory traffic (and the MMA operand fetches competing with it) bounded that
// kernel at ~0.5 TB/s (profiles/r1_conv3x3_timeline.txt).
//
// Trick: the TMA-staged input tile is already a valid tcgen05 A operand for every filter tap.  The tile is stored
// [row][col][channel-block] with one pixel = one swizzle row (128 B for 64 channels, 64 B for 32, 32 B for 16), the output
// tile is 8 pixels wide x 16 rows, so an MMA row group (8 rows of the 128 x K operand) is 8 consecutive pixels of one
// staged row and the 16 groups are one staged row apart: a K-major descriptor with start = tile + ((dy*TWs + dx) pixels),
// stride-byte-offset = one staged row, the tile's swizzle mode, and (for starts that are not aligned to the swizzle
// pattern) the matrix-base-offset field.  Nine taps x Cin/16 MMAs accumulate the whole convolution in TMEM.
//
// Stride 2 (TAPS = 4): the same kernel on the space-to-depth view x'[i, j, (sy, sx, c)] = x[2i + sy, 2j + sx, c] (4C channels,
// H/2 x W/2): a 3x3 / stride-2 / pad-1 conv over x is a 2x2 / stride-1 conv over x' with taps (row' i-1 | i) x (col' j-1 | j)
// and zero weights for the (row' i-1, sy = 0) / (col' j-1, sx = 0) combinations.  x' is never materialised: a 5-D tensor map
// with dimensions (2C, sy, W/2, H/2, B) lets the TMA unit stage the two sy halves of every pixel' (2C channels each, one
// swizzle row per pixel) as two channel blocks, exactly like the two 64-channel blocks of a 128-channel stride-1 tile.
//
//   warp 0     TMA: input tiles (+1 pixel halo, zero fill = padding) into a ring; the weights once (resident)
//   warp 1     one thread issues the MMAs of a tile back to back; tcgen05.commit frees the tile slot and publishes the accumulator
//   warps 2-9  epilogue: tcgen05.ld -> folded BatchNorm / bias -> activation (-> + residual) -> 16-byte stores
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

static constexpr int kZcThreads = 320;
static constexpr int kZcTileH = 16, kZcTileW = 8, kZcTHs = 18, kZcTWs = 10;
static constexpr int kZcMaxX = 8;

enum { ZC_MODE_BN_ACT = 0, ZC_MODE_OFFSETS = 1 };

// debug timeline (LDCONV_DBG bit 32), same format as ldconv_conv3x3_umma.cu: warps 0..2 of CTA 0, tile iterations 8..11
static constexpr int kZcTraceN = 96;
__device__ long long g_zc_trace[3 * kZcTraceN * 2];
__device__ int g_zc_trace_n[3];
struct ZcTracer {
    long long* buf; int n; bool on;
    __device__ __forceinline__ void operator()(int it, int tag) {
        if (on && it >= 8 && it < 12 && n < kZcTraceN) { buf[2 * n] = tag; buf[2 * n + 1] = clock64(); ++n; }
    }
};

struct ZcGeom {
    int Cin, Cout, ON, H, W, B;
    int tiles_h, tiles_w, num_tiles;
    int pb;                 // bytes of one pixel row in the staged tile (= channel block * 2): 32, 64 or 128
    int halves;             // channel blocks per pixel (Cin / (pb/2)): 1, or 2 for Cin = 128
    int layout;             // UMMA layout type of the staged tile: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
    int num_kb, xbufs, ldo, ldr, base_mode, dbg;
    uint32_t ofs_b, ofs_x, ofs_aff, ofs_bar, x_bytes, x_tx_bytes, b_bytes, tmem_cols;
};

__device__ __forceinline__ uint64_t zc_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout, uint32_t base_off)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)layout << 61;
    return d;
}

// ISSUERS = 2 (one CTA per SM, i.e. the 64 / 128-channel shapes): a second MMA-issuing warp (warp 10) takes the odd tiles.  The
// issuing thread pays ~1000 cycles of barrier round trips per tile on top of its MMAs (benchmarks/trace_zc.py: 2600 cycles of
// MMA issue, 3650 per tile); with two issuers the waits of one overlap the MMAs of the other and the tensor pipe stays fed.
template <int MODE, int CIN, int TAPS = 9, int ISSUERS = 1>
__global__ void __launch_bounds__(kZcThreads + 32 * (ISSUERS - 1), ISSUERS == 1 ? 2 : 1)
conv3x3_zc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const float* __restrict__ scale, const float* __restrict__ shift,
                  const __nv_bfloat16* __restrict__ residual, void* __restrict__ out_v, int act, ZcGeom g)
{
    using T = __nv_bfloat16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* sX = smem + g.ofs_x;        // [xbufs][halves][18][10][pb]  (each half 1024-aligned)
    uint8_t* sB = smem + g.ofs_b;        // [num_kb][ON][128 B] SWIZZLE_128B
    float* sAff = reinterpret_cast<float*>(smem + g.ofs_aff);
    const uint32_t aff_s = smem_u32(sAff);
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + g.ofs_bar);
    uint64_t* x_empty = x_full + kZcMaxX;
    uint64_t* t_full = x_empty + kZcMaxX;
    uint64_t* t_empty = t_full + 2;
    uint64_t* w_full = t_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ long long s_trace[3 * kZcTraceN * 2];
    const bool tracing = (g.dbg & 32) && blockIdx.x == 0 && lane == 0 && warp < 3;
    ZcTracer tr{s_trace + (warp < 3 ? warp : 0) * kZcTraceN * 2, 0, tracing};
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < kZcMaxX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, g.tmem_cols);
    pdl_wait();       // everything below may read what the previous kernel wrote
    // [0, ON) scale, [ON, 2 ON) shift; halved for SiLU (affine_act16)
    const float aff_pre = MODE == ZC_MODE_OFFSETS ? 1.f : affine_half_for(act);
    for (int o = threadIdx.x; o < g.ON; o += blockDim.x) {
        sAff[o] = aff_pre * ((scale && o < g.Cout) ? scale[o] : 1.f);
        sAff[g.ON + o] = aff_pre * ((shift && o < g.Cout) ? shift[o] : 0.f);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = g.tiles_h * g.tiles_w;
    const uint32_t half_bytes = g.x_bytes / (uint32_t)g.halves;

    // Control warps: warp-uniform loops, one ELECTED lane issues.  Under `if (lane == 0)` every uniform-datapath instruction
    // (UTMALDG, UTCHMMA, commits) is wrapped in an ELECT / BRA.U.ANY loop over the active lanes -- ~10 dependent instructions per
    // MMA on the thread the whole CTA waits for (652 UTCHMMA / 918 such loops in this file's SASS before the change).
    if (warp == 0) {
        const bool leader = elect_one();
        {
            if (leader) {
                mbar_arrive_expect_tx(w_full, (uint32_t)g.num_kb * g.b_bytes);
                for (int kb = 0; kb < g.num_kb; ++kb) tma_load_2d(sB + (size_t)kb * g.b_bytes, &tmW, w_full, kb * 64, 0);
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
                const int buf = it % g.xbufs;
                mbar_wait(&x_empty[buf], ((it / g.xbufs) & 1) ^ 1);
                const int b = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int ti = rem / g.tiles_w, tj = rem % g.tiles_w;
                tr(it, 100000 + it * 100);
                if (!leader) continue;
                mbar_arrive_expect_tx(&x_full[buf], g.x_tx_bytes);
                for (int hf = 0; hf < g.halves; ++hf) {
                    if (TAPS == 4) {    // space-to-depth view: dims (2C [sx, c], sy, W/2, H/2, B); halves = 2 (block = one sy row of
                        // 2C channels) or, for C = 64, 4 (block = one (sy, sx) cell of 64 channels = one 128-byte swizzle row)
                        const int per_sy = g.halves >> 1;
                        tma_load_5d(sX + (size_t)buf * g.x_bytes + (size_t)hf * half_bytes, &tmX, &x_full[buf],
                                    (hf % per_sy) * (g.pb / 2), hf / per_sy, tj * kZcTileW - 1, ti * kZcTileH - 1, b);
                    }
                    else
                        tma_load_4d(sX + (size_t)buf * g.x_bytes + (size_t)hf * half_bytes, &tmX, &x_full[buf], hf * (g.pb / 2),
                                    tj * kZcTileW - 1, ti * kZcTileH - 1, b);
                }
            }
        }
    } else if (warp == 1 || (ISSUERS == 2 && warp == 10)) {
        const bool leader = elect_one();
        {
            const int q = warp == 1 ? 0 : 1;          // this issuer's tiles: it = q, q + ISSUERS, ...
            // The single issuing thread is the serial resource of this kernel (profiles/r1_conv3x3_timeline.txt: ~230 cycles
            // per MMA when the descriptors are rebuilt with run-time arithmetic), so everything that depends only on the
            // channel count is a compile-time constant and the 9 * CIN/16 MMAs of a tile are fully unrolled: per MMA one add
            // for the A descriptor (tile base + constant) and one for the B descriptor.
            // bytes per pixel row / channel blocks per pixel.  Space-to-depth (TAPS = 4, CIN = 4C): a block is one sy half of
            // the pixel' (2C channels = CIN bytes), always two blocks: a TMA box row narrower than the swizzle span is padded to
            // the span in shared memory (measured), so [sy][sx][c] cannot be one 128-byte row when 2C = 32 channels
            constexpr int PB = TAPS == 4 ? (CIN > 128 ? 128 : CIN) : (CIN >= 64 ? 128 : CIN * 2);
            constexpr int HALVES = TAPS == 4 ? (CIN > 128 ? 4 : 2) : (CIN >= 64 ? CIN / 64 : 1);
            constexpr int KS = PB / 32;                            // K steps of 16 per pixel row
            constexpr uint32_t LAYOUT = PB == 128 ? 2u : (PB == 64 ? 4u : 6u);
            const uint32_t idesc = make_idesc_bf16(128, g.ON);
            mbar_wait(w_full, 0);
            const uint64_t descA_hi = zc_desc(0, (uint32_t)(kZcTWs * PB), LAYOUT, 0) & ~0x3fffull;
            const uint64_t descB0 = make_desc_k_sw128(smem_u32(sB));
            const uint32_t b16 = g.b_bytes >> 4;
            const uint32_t x0 = smem_u32(sX) >> 4;
            const uint32_t xstep = g.x_bytes >> 4, hstep = half_bytes >> 4;
            for (int it = q; (long long)blockIdx.x + (long long)it * gridDim.x < g.num_tiles; it += ISSUERS) {
                const int buf = it & 1, xb = it % g.xbufs;
                tr(it, 600000 + it * 100);
                mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
                tr(it, 700000 + it * 100);
                mbar_wait(&x_full[xb], (it / g.xbufs) & 1);
                tr(it, 200000 + it * 100);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * g.ON);
                const uint64_t descA = descA_hi | (uint64_t)(x0 + (uint32_t)xb * xstep);
                if (leader) {
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
                    for (int hf = 0; hf < HALVES; ++hf) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            constexpr int dummy = 0; (void)dummy;
                            constexpr int TW_ = TAPS == 9 ? 3 : 2;      // taps per row of the filter
                            const uint32_t a_delta = (uint32_t)((((tap / TW_) * kZcTWs + (tap % TW_)) * PB + ks * 32) >> 4);
                            const int k = tap * CIN + hf * (PB / 2) + ks * 16;
                            const uint32_t b_delta = (uint32_t)(k >> 6) * b16 + (uint32_t)(((k & 63) * 2) >> 4);
                            mma_bf16_ss(d_tmem, descA + (uint64_t)(a_delta + (uint32_t)hf * hstep), descB0 + (uint64_t)b_delta, idesc,
                                        (uint32_t)((tap | hf | ks) != 0));
                        }
                    }
                }
                mma_commit(&x_empty[xb]);     // the tile slot is free once these MMAs have read it
                mma_commit(&t_full[buf]);
                }
                __syncwarp();
                tr(it, 300000 + it * 100);
            }
        }
    } else if (warp < 10) {
        const int ww = warp - 2;
        const int lg = warp & 3, half = ww >> 2;
        const int chunks16 = g.ON / 16;
        const int ch_begin = half == 0 ? 0 : (chunks16 + 1) / 2, ch_end = half == 0 ? (chunks16 + 1) / 2 : chunks16;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            tr(it, 400000 + it * 100);
            mbar_wait(&t_full[buf], (it >> 1) & 1);
            tr(it, 500000 + it * 100);
            tc_fence_after_sync();
            const int b = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int ti = rem / g.tiles_w, tj = rem % g.tiles_w;
            const int p = lg * 32 + lane;
            const int i = ti * kZcTileH + p / kZcTileW, j = tj * kZcTileW + p % kZcTileW;
            const bool valid = i < g.H && j < g.W;
            const size_t m = ((size_t)b * g.H + i) * g.W + j;
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * g.ON);
            for (int ch = ch_begin; ch < ch_end; ++ch) {
                const int c0 = ch * 16;
                float4 sc4[4], sh4[4];
                if (MODE != ZC_MODE_OFFSETS) affine_load16(aff_s + (uint32_t)c0 * 4u, aff_s + (uint32_t)(g.ON + c0) * 4u, sc4, sh4);
                uint32_t v[16];
                tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                if (!valid || c0 >= g.Cout) continue;
                if (MODE == ZC_MODE_OFFSETS) {
                    float* dst = reinterpret_cast<float*>(out_v) + m * g.ldo + c0;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (c0 + e < g.Cout) dst[e] = __uint_as_float(v[e]) + sAff[g.ON + c0 + e];
                } else {
                    float z[16];
                    affine_act16_r(v, sc4, sh4, act, z);
                    if (residual) {
                        float r0[8], r1[8];
                        Vec16<T>::load(residual + m * g.ldr + c0, r0);
                        Vec16<T>::load(residual + m * g.ldr + c0 + 8, r1);
#pragma unroll
                        for (int e = 0; e < 8; ++e) { z[e] += r0[e]; z[8 + e] += r1[e]; }
                    }
                    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(out_v) + m * g.ldo + c0);
                    pack16_bf16(z, dst[0], dst[1]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
            tr(it, 900000 + it * 100);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
    if (tracing) {
        for (int i = 0; i < 2 * tr.n; ++i) g_zc_trace[warp * kZcTraceN * 2 + i] = tr.buf[i];
        g_zc_trace_n[warp] = tr.n;
    }
}

static constexpr int zc_enabled() { return 1; }
static constexpr int zc_two_issuers() { return 1; }      // one CTA per SM: a second issuing warp takes the odd tiles

int conv3x3_zc_supported(int Cin, int Cout, int s, int mode)
{
    if (!zc_enabled() || s != 1) return 0;
    if (Cin != 16 && Cin != 32 && Cin != 64 && Cin != 128) return 0;
    if (Cout < 1 || Cout > 256) return 0;
    if (mode == ZC_MODE_BN_ACT && Cout % 16 != 0) return 0;
    const int ON = (Cout + 15) / 16 * 16;
    const size_t wbytes = (size_t)((9 * Cin + 63) / 64) * ON * 128;
    const size_t xbytes = (size_t)kZcTHs * kZcTWs * Cin * 2 + 2048;
    // one input buffer is enough to run (TMA and MMA of consecutive tiles then alternate): Cin = 128 -> Cout = 64 (Detect, P4)
    return wbytes + xbytes + 4096 <= 220 * 1024;
}

int conv3x3_zc(const void* x, int ldx, const void* wt, const float* scale, const float* shift, const void* residual,
               int ldr, void* out, int ldo, int B, int Cin, int H, int W, int Cout, int act, int mode, cudaStream_t st)
{
    if (!aligned16(x) || !aligned16(wt) || (ldx % 8) != 0)
        return fail(LDCONV_E_ALIGN, "conv3x3 zero-copy: x / wt must be 16-byte aligned and ldx a multiple of 8");
    if (mode == ZC_MODE_BN_ACT && (!aligned16(out) || (ldo % 8) != 0 || (residual && (!aligned16(residual) || ldr % 8))))
        return fail(LDCONV_E_ALIGN, "conv3x3 zero-copy: out / residual must be 16-byte aligned with strides multiple of 8");
    ZcGeom g;
    g.Cin = Cin; g.Cout = Cout; g.ON = (Cout + 15) / 16 * 16; g.H = H; g.W = W; g.B = B;
    g.tiles_h = (H + kZcTileH - 1) / kZcTileH;
    g.tiles_w = (W + kZcTileW - 1) / kZcTileW;
    const long long nt = (long long)B * g.tiles_h * g.tiles_w;
    if (nt > 0x7fffffffll) return fail(LDCONV_E_ARG, "conv3x3 zero-copy: too many tiles");
    g.num_tiles = (int)nt;
    const int cb = Cin >= 64 ? 64 : Cin;        // channel block = one swizzle row
    g.pb = cb * 2;
    g.halves = Cin / cb;
    g.layout = g.pb == 128 ? 2 : (g.pb == 64 ? 4 : 6);
    g.num_kb = (9 * Cin + 63) / 64;
    g.b_bytes = (uint32_t)g.ON * 128;
    const uint32_t half_tx = (uint32_t)kZcTHs * kZcTWs * g.pb;
    const uint32_t half_pitch = (half_tx + 1023) & ~1023u;
    g.x_tx_bytes = half_tx * g.halves;
    g.x_bytes = half_pitch * g.halves;
    g.ldo = ldo; g.ldr = ldr;
    // measured on B200: the swizzle XOR uses absolute shared-memory address bits, so the descriptor's matrix-base-offset
    // stays 0 even for starts that are not aligned to the swizzle pattern
    g.base_mode = 0;
    { const char* e = getenv("LDCONV_DBG"); g.dbg = e ? atoi(e) : 0; }
    const size_t wbytes = (size_t)g.num_kb * g.b_bytes;
    // two CTAs (two MMA issuers) per SM when the weights + a useful tile ring fit in half the shared memory
    const bool two = wbytes + 3 * (size_t)g.x_bytes + 4096 <= 104 * 1024 && 4 * g.ON <= 512;
    long long xb = ((long long)(two ? 104 : 220) * 1024 - (long long)wbytes - 4096) / (long long)g.x_bytes;
    long long want = ((two ? 64 : 96) * 1024 + g.x_bytes - 1) / g.x_bytes;
    if (want < 2) want = 2;
    if (xb > want) xb = want;
    if (xb > kZcMaxX) xb = kZcMaxX;
    if (xb < 1) return fail(LDCONV_E_ARG, "conv3x3 zero-copy: does not fit shared memory (Cin=%d Cout=%d)", Cin, Cout);
    g.xbufs = (int)xb;
    uint32_t ofs = 0;
    g.ofs_x = ofs; ofs += (uint32_t)g.xbufs * g.x_bytes;
    g.ofs_b = ofs; ofs += (uint32_t)wbytes;
    g.ofs_aff = ofs; ofs += (uint32_t)g.ON * 8;
    ofs = (ofs + 7) & ~7u;
    g.ofs_bar = ofs; ofs += (uint32_t)(2 * kZcMaxX + 5) * 8 + 16;
    const size_t smem = ofs + 1024;
    g.tmem_cols = 32;
    while (g.tmem_cols < (uint32_t)(2 * g.ON)) g.tmem_cols <<= 1;

    CUtensorMap tmX, tmW;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2};
        cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)kZcTWs, (cuuint32_t)kZcTHs, 1};
        const CUtensorMapSwizzle sw = g.pb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : (g.pb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gdim, gstr, box, sw)) return e;
    }
    {
        const int K = 9 * Cin;
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)g.ON};
        if (int e = encode_map(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return e;
    }
    int grid = num_sms() * (two ? 2 : 1);
    if (grid > g.num_tiles) grid = g.num_tiles;
#define LDC_ZC_LAUNCH(MODE_, CIN_)                                                                                     \
    do {                                                                                                               \
        if (two || !zc_two_issuers() || g.xbufs < 2) {      /* two issuers alternate tiles: each needs its own input buffer */ \
            auto kern = conv3x3_zc_kernel<MODE_, CIN_, 9, 1>;                                                          \
            LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
            LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(kZcThreads), smem, st, tmX, tmW, scale, shift,                    \
                                (const __nv_bfloat16*)residual, out, act, g));                                        \
        } else {                                                                                                       \
            auto kern = conv3x3_zc_kernel<MODE_, CIN_, 9, 2>;                                                          \
            LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
            LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(kZcThreads + 32), smem, st, tmX, tmW, scale, shift,               \
                                (const __nv_bfloat16*)residual, out, act, g));                                        \
        }                                                                                                              \
    } while (0)
#define LDC_ZC_CIN(MODE_)                                                                                              \
    switch (Cin) {                                                                                                     \
        case 16: LDC_ZC_LAUNCH(MODE_, 16); break;                                                                      \
        case 32: LDC_ZC_LAUNCH(MODE_, 32); break;                                                                      \
        case 64: LDC_ZC_LAUNCH(MODE_, 64); break;                                                                      \
        default: LDC_ZC_LAUNCH(MODE_, 128); break;                                                                     \
    }
    if (mode == ZC_MODE_OFFSETS) { LDC_ZC_CIN(ZC_MODE_OFFSETS) } else { LDC_ZC_CIN(ZC_MODE_BN_ACT) }
#undef LDC_ZC_CIN
#undef LDC_ZC_LAUNCH
    LDC_LAUNCH_CHECK("conv3x3_zc_kernel");
    set_impl(LDCONV_IMPL_TCGEN05);
    return LDCONV_OK;
}

// ---- stride-2 offset conv on the space-to-depth view (TAPS = 4) ---------------------------------------------------------
int conv3x3_zc_s2d_supported(int C, int N, int H, int W)
{
    if (!zc_enabled()) return 0;
    return (C == 16 || C == 32 || C == 64) && N >= 1 && 2 * N <= 16 && H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2;
}

// x (B,H,W,C) bf16 dense; wt (2N, 16C) bf16 with k = ((ty*2+tx)*2+sy)*2C + sx*C + c (see the header comment); bias (2N) fp32
// or NULL; off (B,H/2,W/2,2N) fp32
int conv3x3_zc_s2d(const void* x, const void* wt, const float* bias, float* off, int B, int C, int H, int W, int N,
                   cudaStream_t st)
{
    if (!aligned16(x) || !aligned16(wt)) return fail(LDCONV_E_ALIGN, "offset conv (space-to-depth): x / wt must be 16-byte aligned");
    const int Cin = 4 * C, Cout = 2 * N, h = H / 2, w = W / 2;
    ZcGeom g;
    g.Cin = Cin; g.Cout = Cout; g.ON = 16; g.H = h; g.W = w; g.B = B;
    g.tiles_h = (h + kZcTileH - 1) / kZcTileH;
    g.tiles_w = (w + kZcTileW - 1) / kZcTileW;
    const long long nt = (long long)B * g.tiles_h * g.tiles_w;
    if (nt > 0x7fffffffll) return fail(LDCONV_E_ARG, "offset conv (space-to-depth): too many tiles");
    g.num_tiles = (int)nt;
    g.pb = Cin > 128 ? 128 : Cin;     // one sy half of a pixel' = 2C channels = Cin bytes (64 or 128); C = 64: one (sy, sx) cell
    g.halves = Cin > 128 ? 4 : 2;
    g.layout = g.pb == 128 ? 2 : 4;
    g.num_kb = (4 * Cin + 63) / 64;
    g.b_bytes = (uint32_t)g.ON * 128;
    const uint32_t half_tx = (uint32_t)kZcTHs * kZcTWs * g.pb;
    const uint32_t half_pitch = (half_tx + 1023) & ~1023u;
    g.x_tx_bytes = half_tx * g.halves;
    g.x_bytes = half_pitch * g.halves;
    g.ldo = Cout; g.ldr = 0;
    g.base_mode = 0;
    { const char* e = getenv("LDCONV_DBG"); g.dbg = e ? atoi(e) : 0; }
    const size_t wbytes = (size_t)g.num_kb * g.b_bytes;
    // two CTAs (two MMA issuers) per SM when the weights + two input tiles fit in half the shared memory (C = 16)
    const bool two = wbytes + 2 * (size_t)g.x_bytes + 4096 <= 104 * 1024;
    long long xb = ((long long)(two ? 104 : 220) * 1024 - (long long)wbytes - 4096) / (long long)g.x_bytes;
    if (xb > kZcMaxX) xb = kZcMaxX;
    if (xb > 3) xb = 3;
    if (xb < 2) return fail(LDCONV_E_ARG, "offset conv (space-to-depth): does not fit shared memory");
    g.xbufs = (int)xb;
    uint32_t ofs = 0;
    g.ofs_x = ofs; ofs += (uint32_t)g.xbufs * g.x_bytes;
    g.ofs_b = ofs; ofs += (uint32_t)wbytes;
    g.ofs_aff = ofs; ofs += (uint32_t)g.ON * 8;
    ofs = (ofs + 7) & ~7u;
    g.ofs_bar = ofs; ofs += (uint32_t)(2 * kZcMaxX + 5) * 8 + 16;
    const size_t smem = ofs + 1024;
    g.tmem_cols = 32;

    CUtensorMap tmX, tmW;
    {
        // x viewed as (2C [sx, c], sy, W/2, H/2, B): innermost two dimensions make pixel' = [sy][sx][c]
        cuuint64_t gdim[5] = {(cuuint64_t)(2 * C), 2, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
        cuuint64_t gstr[4] = {(cuuint64_t)W * C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[5] = {(cuuint32_t)(g.pb / 2), 1, (cuuint32_t)kZcTWs, (cuuint32_t)kZcTHs, 1};
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, gdim, gstr, box,
                               g.pb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    }
    {
        const int K = 4 * Cin;
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)g.ON};
        if (int e = encode_map(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    int grid = num_sms() * (two ? 2 : 1);
    if (grid > g.num_tiles) grid = g.num_tiles;
    if (Cin == 256) {
        auto kern = conv3x3_zc_kernel<ZC_MODE_OFFSETS, 256, 4>;
        LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(kZcThreads), smem, st, tmX, tmW, (const float*)nullptr, bias,
                            (const __nv_bfloat16*)nullptr, (void*)off, (int)LDCONV_ACT_NONE, g));
    } else if (Cin == 64) {
        auto kern = conv3x3_zc_kernel<ZC_MODE_OFFSETS, 64, 4>;
        LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(kZcThreads), smem, st, tmX, tmW, (const float*)nullptr, bias,
                            (const __nv_bfloat16*)nullptr, (void*)off, (int)LDCONV_ACT_NONE, g));
    } else {
        auto kern = conv3x3_zc_kernel<ZC_MODE_OFFSETS, 128, 4>;
        LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(kZcThreads), smem, st, tmX, tmW, (const float*)nullptr, bias,
                            (const __nv_bfloat16*)nullptr, (void*)off, (int)LDCONV_ACT_NONE, g));
    }
    LDC_LAUNCH_CHECK("conv3x3_zc_kernel (space-to-depth)");
    set_impl(LDCONV_IMPL_TCGEN05);
    return LDCONV_OK;
}

}  // namespace ldc

// LDConv's offset conv at stride 2 (conv.py:356,368) as a zero-copy tcgen05 implicit GEMM on the space-to-depth view of x.
// w_s2d (2N, 16*C) bf16: the reference's (2N,C,3,3) weight scattered into k = ((ty*2+tx)*2+sy)*2C + sx*C + c with
// (ty, sy) = (0,1), (1,0), (1,1) for ky = 0, 1, 2 (same for kx -> (tx, sx)) and zeros elsewhere; the Python module builds it.
LDC_API int ldconv_offset_conv_s2d_supported(int C, int N, int H, int W, int dtype)
{
    return dtype == LDCONV_BF16 ? ldc::conv3x3_zc_s2d_supported(C, N, H, W) : 0;
}

LDC_API int ldconv_offset_conv_s2d_fwd(const void* x, const void* w_s2d, const float* bias, float* off, int B, int C, int H, int W,
                                       int N, int dtype, void* stream)
{
    using namespace ldc;
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_offset_conv_s2d_fwd: bf16 only");
    LDC_REQUIRE(x && w_s2d && off && B >= 0, "ldconv_offset_conv_s2d_fwd: bad arguments");
    LDC_REQUIRE(conv3x3_zc_s2d_supported(C, N, H, W), "ldconv_offset_conv_s2d_fwd: shape not covered (C=%d N=%d H=%d W=%d)", C, N, H, W);
    if (B == 0) return LDCONV_OK;
    return conv3x3_zc_s2d(x, w_s2d, bias, off, B, C, H, W, N, (cudaStream_t)stream);
}

LDC_API int ldconv_debug_trace_zc(long long* host_out, int max_pairs)
{
    int n[3] = {0, 0, 0};
    long long all[3 * ldc::kZcTraceN * 2];
    cudaMemcpyFromSymbol(n, ldc::g_zc_trace_n, sizeof(n));
    cudaMemcpyFromSymbol(all, ldc::g_zc_trace, sizeof(all));
    int k = 0;
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < n[r] && k < max_pairs; ++i, ++k) {
            host_out[2 * k] = all[(r * ldc::kZcTraceN + i) * 2];
            host_out[2 * k + 1] = all[(r * ldc::kZcTraceN + i) * 2 + 1];
        }
    return k;
}

