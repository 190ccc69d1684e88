// ldconv_onepass_umma.cu -- the WHOLE LDConv inference forward as ONE persistent tcgen05 kernel (bf16, sm_100a): x is read once.
//
// Replaces /root/reference/ultralytics/nn/modules/conv.py:366-410 (LDConv.forward, eval mode): the offset conv p_conv(x) (:356,
// :368), the sampling grid p_0 + p_n + offset (:413-454), floor / independent clamps / corner indices / bilinear weights
// (:375-393), the four gathers and the bilinear sum (:396-405), the rearrange (:494-503) and Conv2d((N,1),(N,1)) +
// BatchNorm2d + SiLU (:355, :408).  Per call the kernel moves x + out; neither the offsets nor the (M, N*C) operand reach HBM
// (ldconv_gather_gemm_fwd still reads the offsets a separate offset-conv launch wrote after its own pass over x).
//
// One TMA-staged input tile serves BOTH consumers:
//   * the 3x3 offset conv as a zero-copy tcgen05 implicit GEMM (the trick of ldconv_conv3x3_zc.cu): the tile is stored
//     [row][col][channel block] with one pixel = one swizzle row, the output tile is 16 rows x 8 pixels, so the A operand of
//     every filter tap is the staged tile itself behind a K-major descriptor (start = tap shift, stride-byte-offset = one
//     staged row); N = 16 accumulator columns (2 num_param offsets) in TMEM.  Stride 2 runs on the space-to-depth view
//     x'[i, j, (sy, sx, c)] through a 5-D tensor map: 4 taps instead of 9;
//   * the bilinear gather: the corner records hold SWIZZLED shared-memory addresses of the same tile (the swizzle is a pure
//     function of the byte offset inside the 1024-byte-aligned tile: 16-byte chunk index ^= bits [7, 7 + log2(span / 16)) of the
//     offset), a thread's channel vector enters with one XOR; samples that leave the tile (2-pixel halo) come from L2.
// Pipeline of a CTA (persistent, 16 x 8 output pixels per step, up to three CTAs per SM):
//   issuer warp (one thread)   TMA of the input tiles; offset-conv MMAs of tile t+1 + tcgen05.commit -> off_done;
//                              main MMAs of tile t (K/16 x (M = 128, N = O)) + commit -> mma_done[t & 1]
//   worker warps (4 TG warps)  phase 1: tcgen05.ld of the tile's offsets (+ bias) -> common.cuh::make_point_grid (bit-exact
//                              grid / indices / weights) -> one record per sample                                | barrier A
//                              phase 2: four 16-byte corner loads per (sample, channel vector), bilinear sum on packed fp32
//                              pairs, one 16-byte store into the K-major SWIZZLE_128B operand tile               | barrier B
//                              epilogue of the PREVIOUS tile while this tile's MMAs run: tcgen05.ld -> folded BatchNorm ->
//                              SiLU -> 16-byte NHWC stores (out may be a channel slice of a concat buffer)
// The arithmetic is the same code as ldconv_gather_gemm_fwd + ldconv_offset_conv_{tc,s2d}_fwd (same MMA order, same
// make_point_grid / bilinear_bf16x2 / affine), so the two paths agree bit for bit (tests/test_gpu_parity.py).
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

struct OPGeom {
    int N, H, W, h, w, O, ON, K, num_kb, ksteps;
    int tiles_h, tiles_w, num_tiles, ldo, act, per_sm;
    int XB, AB, xb_mask, xb_shift, ab_mask;
    int tstore, st_rowb;               // TMA-store epilogue: bytes per staged output row (min(O, 64) * 2), 0 = direct stores
    float hm, wm;
    unsigned long long img_bytes;
    unsigned inv_img, inv_tw;
    uint32_t ofs_a, ofs_b, ofs_bc, ofs_x, ofs_rec, ofs_aff, ofs_bias, ofs_bar, ofs_stage;
    uint32_t a_bytes, b_bytes, x_bytes, x_tx_bytes, tmem_cols;
};

// debug timeline (ldconv_debug_onepass_trace): clock64 stamps of CTA 0's first worker thread and issuer thread for tile iterations
// 6..9; two uniform branches per stamp, not inside the unrolled loops
static constexpr int kOpTraceN = 64;
struct OpTracer {
    long long* buf; int n; bool on;
    __device__ __forceinline__ void operator()(int it, int tag) {
        if (on && it >= 6 && it < 10 && n < kOpTraceN) { buf[2 * n] = (long long)it * 100 + tag; buf[2 * n + 1] = clock64(); ++n; }
    }
};
static thread_local long long* g_op_trace = nullptr;

static constexpr int kOpTH = 16, kOpTW = 8;       // output tile: MMA row group (8 rows) = 8 consecutive pixels of one row

__device__ __forceinline__ uint4 op_lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 op_lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void op_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void op_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void op_tmem_alloc_keep_permit(uint32_t* dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
}

// UMMA shared-memory descriptor of a K-major operand with stride-byte-offset sbo and swizzle layout code `layout`
__device__ __forceinline__ uint64_t op_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// compile-time shape of the staged tile for (channel-vector shift TCVS: C = 8 << TCVS, stride TS)
template <int TCVS, int TS> struct OPShape {
    static constexpr int C = 8 << TCVS;
    static constexpr int TAPS = TS == 1 ? 9 : 4;
    static constexpr int CIN = TS == 1 ? C : 4 * C;                       // K per tap
    static constexpr int PB = TS == 1 ? (C >= 64 ? 128 : C * 2) : (CIN > 128 ? 128 : CIN);      // bytes per staged pixel row
    static constexpr int HALVES = TS == 1 ? (C >= 64 ? C / 64 : 1) : (CIN > 128 ? 4 : 2);       // channel blocks per pixel(')
    static constexpr int KS = PB / 32;                                    // K steps of 16 per pixel row
    static constexpr uint32_t LAYOUT = PB == 128 ? 2u : (PB == 64 ? 4u : 6u);
    static constexpr uint32_t SWM = (uint32_t)(PB / 16 - 1) << 4;         // swizzle: chunk bits [4, 4 + log2(PB / 16)) ^= offset bits [7, ...)
    // stride 1: 16 x 8 outputs + 1 (second corner) + 2-pixel halo on every side; stride 2: 18 x 10 pixels' = input rows
    // 2 i0 - 2 .. 2 i0 + 33 (the space-to-depth tile of ldconv_conv3x3_zc.cu; 2 input pixels of halo above / left, 1 below / right
    // of the farthest second corner)
    static constexpr int THs = TS == 1 ? kOpTH + 1 + 4 : kOpTH + 2;
    static constexpr int TWs = TS == 1 ? kOpTW + 1 + 4 : kOpTW + 2;
    static constexpr int RIN = TS == 1 ? THs : 2 * THs, KIN = TS == 1 ? TWs : 2 * TWs;          // tile extent in INPUT pixels
    static constexpr int ORG = 2;                                         // input pixels between the tile origin and output pixel (i0 s, j0 s)
    static constexpr uint32_t HALF_TX = (uint32_t)THs * TWs * PB;
    static constexpr uint32_t HALF_BYTES = (HALF_TX + 1023u) & ~1023u;
    static constexpr uint32_t X_BYTES = HALF_BYTES * HALVES, X_TX = HALF_TX * HALVES;
    static constexpr int KC = TAPS * CIN;                                 // K of the offset conv
    static constexpr int KC_KB = (KC + 63) / 64;                          // 64-wide weight blocks of [16 rows][128 B]
    // stride 2: byte offset of input row parity / column parity inside a pixel'
    static constexpr uint32_t SYB = (uint32_t)(HALVES / 2) * HALF_BYTES;
    static constexpr uint32_t SXB = C <= 32 ? (uint32_t)C * 2u : HALF_BYTES;
    // Space-to-depth weights are zero wherever a 2x2 tap reaches outside the 3x3 window: filter row ky = 2 ty + sy - 1 needs
    // (ty, sy) != (0, 0), likewise for the columns -- 7 of the 16 (tap, sy, sx) cells.  Their MMAs add exact zeros and are not
    // issued (9/16 of the tensor work), all-zero 64-wide weight blocks are not staged (C = 64: 9 of 16 blocks).
    static constexpr bool mma_zero(int tap, int hf, int ks)
    {
        if (TS == 1) return false;
        const int ty = tap / 2, tx = tap % 2;
        const int sy = HALVES == 4 ? (hf >> 1) : hf;
        const int sx = HALVES == 4 ? (hf & 1) : (ks * 16) / C;
        return (ty == 0 && sy == 0) || (tx == 0 && sx == 0);
    }
    static constexpr bool blk_zero(int kb)
    {
        for (int k = kb * 64; k < kb * 64 + 64 && k < KC; k += 16) {
            const int tap = k / CIN, rem = k % CIN;
            if (!mma_zero(tap, rem / (PB / 2), (rem % (PB / 2)) / 16)) return false;
        }
        return true;
    }
    static constexpr int slot(int kb)      // index of weight block kb among the staged (non-zero) blocks
    {
        int n = 0;
        for (int i = 0; i < kb; ++i) n += blk_zero(i) ? 0 : 1;
        return n;
    }
    static constexpr int NZ_KB = slot(KC_KB);
};

template <int TN, int TCVS, int TS, int TG, int MINB>
__global__ void __launch_bounds__(128 * TG + 32, MINB)
ldconv_onepass_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmWc, const __grid_constant__ CUtensorMap tmO,
                      const __nv_bfloat16* __restrict__ x,
                      const float* __restrict__ bias, const int* __restrict__ pn, const float* __restrict__ scale,
                      const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, float* __restrict__ off_dbg,
                      long long* __restrict__ trace, const OPGeom g)
{
    using S = OPShape<TCVS, TS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_s = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.ofs_bar);
    uint64_t* x_full = bars;             // [2] TMA bytes of an input tile have landed
    uint64_t* mma_done = bars + 2;       // [2] main MMAs of the tile that used TMEM / operand buffer i are complete
    uint64_t* off_done = bars + 4;       // offset-conv MMAs of a tile are complete (and have released nothing: the tile stays)
    uint64_t* b_done = bars + 5;         // the workers have passed barrier B of a tile (operand complete, input tile consumed)
    uint64_t* w_full = bars + 6;
    uint64_t* e_done = bars + 7;         // every worker warp has parked its part of a tile's output in the staging tile (TMA-store epilogue)
    uint64_t* s_free = bars + 8;         // the TMA store of the staged tile has read it: the next epilogue may overwrite
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);       // [0] main accumulators, [1] offset accumulator
    float* sAff = reinterpret_cast<float*>(smem + g.ofs_aff);          // [0, ON) scale, [ON, 2 ON) shift (halved for SiLU)
    float* sBias = reinterpret_cast<float*>(smem + g.ofs_bias);        // [16] offset-conv bias
    const uint32_t aff_s = smem_s + g.ofs_aff;

    constexpr int NW = 128 * TG;         // worker threads: TG groups of 128 (group = n phase of the samples / slice of the epilogue)
    constexpr int N = TN;
    constexpr int cvs = TCVS;
    constexpr int s = TS;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int tiles_per_img = g.tiles_h * g.tiles_w;
    constexpr uint32_t rec_g_ofs = (uint32_t)(128 * N) * 16u;          // weights follow the 128 N address records

    struct TC { int b, i0, j0; };
    auto tile_coords = [&](int tile) {
        TC t;
        t.b = g.inv_img ? (int)__umulhi((unsigned)tile, g.inv_img) : tile;
        const int rem = tile - t.b * tiles_per_img;
        const int ti = g.inv_tw ? (int)__umulhi((unsigned)rem, g.inv_tw) : rem;
        t.i0 = ti * kOpTH;
        t.j0 = (rem - ti * g.tiles_w) * kOpTW;
        return t;
    };

    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmWc);
        if (g.tstore) tma_prefetch_desc(&tmO);
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&mma_done[i], 1); }
        mbar_init(off_done, 1);
        mbar_init(b_done, 1);
        mbar_init(w_full, 1);
        mbar_init(e_done, 4 * TG);
        mbar_init(s_free, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        op_tmem_alloc_keep_permit(&tmem_slot[0], g.tmem_cols);
        tmem_alloc(&tmem_slot[1], 32);
    }
    pdl_wait();       // everything below may read what the previous kernel wrote (x, scale / shift)
    const bool act_silu = g.act == LDCONV_ACT_SILU;
    {
        const float pre = act_silu ? 0.5f : 1.f;      // silu(z) = hz + hz tanh(hz), hz = z / 2: the halving is exact
        for (int o = tid; o < g.ON; o += NW + 32) {
            sAff[o] = pre * ((scale && o < g.O) ? scale[o] : 1.f);
            sAff[g.ON + o] = pre * ((shift && o < g.O) ? shift[o] : 0.f);
        }
        if (tid < 16) sBias[tid] = (bias && tid < 2 * N) ? bias[tid] : 0.f;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot[0];
    const uint32_t tmem_off = tmem_slot[1];

    OpTracer tr{trace ? trace + (warp == 4 * TG ? 2 * kOpTraceN + 2 : 0) : nullptr, 0,
                trace != nullptr && blockIdx.x == 0 && (tid == 0 || tid == NW)};
    if (warp == 4 * TG) {
        // ===================================== issuer warp: warp-uniform control flow, one elected lane issues =====================
        // (under `if (lane == 0)` the compiler wraps every uniform-datapath instruction -- UTCHMMA, UTMALDG -- in an ELECT /
        // BRA.U.ANY loop over the active lanes: ~10 dependent instructions and ~85 cycles per MMA with 27 warps sharing the SM,
        // 1500 cycles for the 18 offset-conv MMAs of a 32-channel tile -- benchmarks/trace_onepass.py)
        const bool leader = elect_one();
        {
            auto issue_x_tile_at = [&](const TC& t, int xb) {
                if (!leader) return;
                mbar_arrive_expect_tx(&x_full[xb], S::X_TX);
                uint8_t* dst = smem + g.ofs_x + (size_t)xb * S::X_BYTES;
#pragma unroll
                for (int hf = 0; hf < S::HALVES; ++hf) {
                    if constexpr (TS == 1)      // (channel block, column, row, image); out-of-range pixels arrive as zeros = the conv's padding
                        tma_load_4d(dst + (size_t)hf * S::HALF_BYTES, &tmX, &x_full[xb], hf * 64, t.j0 - S::ORG, t.i0 - S::ORG, t.b);
                    else {            // space-to-depth view (2C [sx, c], sy, W/2, H/2, B): one block = one sy row (C <= 32) / one (sy, sx) cell
                        constexpr int per_sy = S::HALVES / 2;
                        tma_load_5d(dst + (size_t)hf * S::HALF_BYTES, &tmX, &x_full[xb], (hf % per_sy) * (S::PB / 2), hf / per_sy,
                                    t.j0 - 1, t.i0 - 1, t.b);
                    }
                }
            };
            // L2 prefetch of a tile one step before its TMA: with a single input buffer the load can only be issued once the workers
            // have consumed the previous tile, and its latency sits on the tile's critical path (TMA -> offset conv -> phase 1);
            // measured: 693 -> 677 us over the nine yolov8-LD-P2 layers at batch 64
            auto issue_x_tile = [&](int tile, int xb) { issue_x_tile_at(tile_coords(tile), xb); };
            auto prefetch_x_tile_at = [&](const TC& t) {
                if (!leader) return;
#pragma unroll
                for (int hf = 0; hf < S::HALVES; ++hf) {
                    if constexpr (TS == 1)
                        tma_prefetch_4d(&tmX, hf * 64, t.j0 - S::ORG, t.i0 - S::ORG, t.b);
                    else {
                        constexpr int per_sy = S::HALVES / 2;
                        tma_prefetch_5d(&tmX, (hf % per_sy) * (S::PB / 2), hf / per_sy, t.j0 - 1, t.i0 - 1, t.b);
                    }
                }
            };
            auto prefetch_x_tile = [&](int tile) { prefetch_x_tile_at(tile_coords(tile)); };
            // TMA store of the staged output tile of iteration `eit` once every worker warp has parked its part: coordinates (channel,
            // column, row, image), pixels beyond the map are clipped by the tensor map; then the staging tile is handed back
            auto store_tile = [&](int eit, bool last) {
                mbar_wait_sleep(e_done, (uint32_t)eit & 1u);
                const TC t = tile_coords((int)blockIdx.x + eit * (int)gridDim.x);
                if (!leader) return;
                for (int cb = 0; cb * 64 < g.O; ++cb)
                    tma_store_4d(&tmO, smem + g.ofs_stage + (size_t)cb * (128 * 128), cb * 64, t.j0, t.i0, t.b);
                tma_store_commit();
                if (last) tma_store_wait_all(); else tma_store_wait_read();
                mbar_arrive(s_free);
            };
            const uint32_t idesc_c = make_idesc_bf16(128, 16);
            const uint32_t idesc_m = make_idesc_bf16(128, g.ON);
            const uint64_t descA_hi = op_desc(0, (uint32_t)(S::TWs * S::PB), S::LAYOUT) & ~0x3fffull;
            const uint64_t descBc0 = make_desc_k_sw128(smem_s + g.ofs_bc);
            const uint64_t descBm0 = make_desc_k_sw128(smem_s + g.ofs_b);
            // the 3x3 (stride 1) / 2x2-on-space-to-depth (stride 2) offset conv of the tile staged in buffer xb: TAPS * HALVES * KS
            // MMAs of M = 128 pixels x N = 16 x K = 16, fully unrolled (per MMA one add per descriptor), same order as
            // conv3x3_zc_kernel so the fp32 accumulation is bit-identical to ldconv_offset_conv_{tc,s2d}_fwd
            auto offset_conv_mma = [&](int xb) {
                const uint64_t descA = descA_hi | (uint64_t)(((smem_s + g.ofs_x + (uint32_t)xb * S::X_BYTES) >> 4) & 0x3fffu);
                if (!leader) return;
                uint32_t acc = 0;
#pragma unroll
                for (int tap = 0; tap < S::TAPS; ++tap) {
#pragma unroll
                    for (int hf = 0; hf < S::HALVES; ++hf) {
#pragma unroll
                        for (int ks = 0; ks < S::KS; ++ks) {
                            if (S::mma_zero(tap, hf, ks)) continue;                     // all-zero weights (space-to-depth padding cells)
                            constexpr int TWf = TS == 1 ? 3 : 2;                        // taps per filter row
                            constexpr int TO = TS == 1 ? S::ORG - 1 : 0;                // tile origin -> first tap, in staged pixels
                            const uint32_t a_delta = (uint32_t)(((((tap / TWf) + TO) * S::TWs + (tap % TWf) + TO) * S::PB + ks * 32) >> 4) +
                                                     (uint32_t)hf * (S::HALF_BYTES >> 4);
                            const int k = tap * S::CIN + hf * (S::PB / 2) + ks * 16;
                            const uint32_t b_delta = (uint32_t)S::slot(k >> 6) * (2048u >> 4) + (uint32_t)(((k & 63) * 2) >> 4);
                            mma_bf16_ss(tmem_off, descA + (uint64_t)a_delta, descBc0 + (uint64_t)b_delta, idesc_c, acc);
                            acc = 1;
                        }
                    }
                }
                mma_commit(off_done);
            };
            auto main_mma = [&](int ab, int tb) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(tb * g.ON);
                const uint64_t da = make_desc_k_sw128(smem_s + g.ofs_a + (uint32_t)ab * g.a_bytes);
                if (!leader) return;
                for (int st = 0; st < g.ksteps; ++st) {
                    const uint32_t kb = (uint32_t)st >> 2, kk = (uint32_t)st & 3;
                    mma_bf16_ss(d_tmem, da + (uint64_t)(kb * 1024u + kk * 2u), descBm0 + (uint64_t)(kb * (g.b_bytes >> 4) + kk * 2u),
                                idesc_m, (uint32_t)(st != 0));
                }
                mma_commit(&mma_done[tb]);
            };

            if (leader) {
                mbar_arrive_expect_tx(w_full, (uint32_t)g.num_kb * g.b_bytes + (uint32_t)S::NZ_KB * 2048u);
                for (int kb = 0; kb < g.num_kb; ++kb) tma_load_2d(smem + g.ofs_b + (size_t)kb * g.b_bytes, &tmW, w_full, kb * 64, 0);
#pragma unroll
                for (int kb = 0; kb < S::KC_KB; ++kb)
                    if (!S::blk_zero(kb)) tma_load_2d(smem + g.ofs_bc + (size_t)S::slot(kb) * 2048, &tmWc, w_full, kb * 64, 0);
            }
            for (int k = 0; k < g.XB; ++k)
                if ((int)blockIdx.x + k * (int)gridDim.x < g.num_tiles) issue_x_tile(blockIdx.x + k * gridDim.x, k);
            if ((int)blockIdx.x + g.XB * (int)gridDim.x < g.num_tiles) prefetch_x_tile(blockIdx.x + g.XB * gridDim.x);
            if ((int)blockIdx.x < g.num_tiles) {
                mbar_wait_sleep(w_full, 0);
                mbar_wait_sleep(&x_full[0], 0);
                tc_fence_after_sync();
                offset_conv_mma(0);
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
                const int xb = it & g.xb_mask, ab = it & g.ab_mask, tb = it & 1;
                const bool more = tile + (int)gridDim.x < g.num_tiles;
                // everything that does not depend on the workers is computed before the wait: this thread's instruction stream after
                // barrier B is the critical path of the next tile (TMA -> offset conv -> phase 1)
                const TC t_next = tile_coords(more ? tile + (int)gridDim.x : tile);
                const bool more2 = tile + 2 * (int)gridDim.x < g.num_tiles;
                const TC t_pf = tile_coords(more2 ? tile + 2 * (int)gridDim.x : tile);
                tr(it, 50);
                mbar_wait_sleep(b_done, (uint32_t)it & 1u);      // operand tile of `it` complete, its input tile consumed, epilogue(it - 2) done
                tc_fence_after_sync();
                tr(it, 51);
                if (g.XB == 2) {
                    // the next tile is already staged: its offsets first (phase 1 of the workers waits for them), then this tile's GEMM
                    if (more) {
                        mbar_wait_sleep(&x_full[(it + 1) & 1], (uint32_t)((it + 1) >> 1) & 1u);
                        tc_fence_after_sync();
                        offset_conv_mma((it + 1) & 1);
                    }
                    main_mma(ab, tb);
                    if (tile + 2 * (int)gridDim.x < g.num_tiles) issue_x_tile(tile + 2 * gridDim.x, xb);
                    if (tile + 3 * (int)gridDim.x < g.num_tiles) prefetch_x_tile(tile + 3 * gridDim.x);
                } else {
                    if (more) issue_x_tile_at(t_next, 0);      // the only input buffer is free again: refill first (latency)
                    tr(it, 52);
                    main_mma(ab, tb);
                    tr(it, 53);
                    if (more) {
                        mbar_wait_sleep(&x_full[0], (uint32_t)(it + 1) & 1u);
                        tc_fence_after_sync();
                        tr(it, 54);
                        offset_conv_mma(0);
                        tr(it, 55);
                    }
                    if (more2) prefetch_x_tile_at(t_pf);
                }
                // epilogue(it - 1) runs right after barrier B of this tile: its output is parked soon
                if (g.tstore && it > 0) store_tile(it - 1, false);
            }
            if (g.tstore && it > 0) store_tile(it - 1, true);
        }
    } else {
        // ===================================================== worker warps =====================================================
        // per-thread constants: one pixel of the tile for phase 1 and the epilogue, one channel vector for phase 2
        const int p = tid & 127, n0 = tid >> 7;
        const int di = p >> 3, dj = p & 7;
        bool has[2];
        int br[2], bk[2], ns[2];                                   // di * s + pn_r[n], dj * s + pn_k[n], n of sample rounds 0, 1
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int n = n0 + TG * r;
            has[r] = (TN <= TG * r) ? false : n < N;
            ns[r] = has[r] ? n : 0;
            br[r] = di * s + (has[r] ? pn[n] : 0);
            bk[r] = dj * s + (has[r] ? pn[N + n] : 0);
        }
        static_assert(TN <= 2 * TG, "one thread handles at most two samples of its pixel");

        constexpr int CVM = (1 << cvs) - 1;
        const int cv = tid & CVM, sx0 = tid >> cvs;                // phase 2: item round k handles sample sx0 + k * (NW >> cvs)
        constexpr int spr = NW >> cvs;
        // channel vector -> XOR mask on the swizzled corner address (+ the second 64-channel block of a 128-channel pixel)
        const uint32_t cx = (uint32_t)(cv & (S::PB / 16 - 1)) << 4;
        const uint32_t hoff = (TS == 1 && TCVS == 4) ? (uint32_t)(cv >> 3) * S::HALF_BYTES : 0u;
        const int grp = warp >> 2;
        const uint32_t stage_s = smem_s + g.ofs_stage;
        const int chunks = g.ON / 16;
        const int ch_begin = chunks * grp / TG, ch_end = chunks * (grp + 1) / TG;

        // epilogue of one finished tile: TMEM lane = this thread's pixel, this thread group's slice of the 16-column chunks
        auto epilogue = [&](int m_out, int eit) {      // m_out: output pixel index ((b h + i) w + j) of this thread, -1 outside the map
            const int tb = eit & 1;
            mbar_wait(&mma_done[tb], (eit >> 1) & 1);
            tc_fence_after_sync();
            if (g.tstore && eit > 0) mbar_wait(s_free, (uint32_t)(eit - 1) & 1u);      // the previous tile's store has read the staging tile
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
                    const float4 sc = op_lds_f4(aff_s + (uint32_t)(c0 + e) * 4u);
                    const float4 sh = op_lds_f4(aff_s + (uint32_t)(g.ON + c0 + e) * 4u);
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
                if (g.tstore) {
                    // park the 32 bytes in the staging tile [64-channel block][128 pixels][row] at the tensor map's swizzle: the 8
                    // lanes of a quarter-warp (8 consecutive pixels) land on 8 different 16-byte bank slots, and ONE TMA store per tile
                    // writes whole lines (a thread's own 16-byte global stores cost an L1 request each and half-filled L2 sectors)
                    const uint32_t rowb = (uint32_t)g.st_rowb;
                    const uint32_t row = stage_s + (uint32_t)(c0 >> 6) * (128u * 128u) + (uint32_t)p * rowb;
                    const uint32_t key = ((uint32_t)p * rowb >> 7) & (rowb / 16u - 1u);
                    const uint32_t cc = (uint32_t)(c0 & 63) >> 3;
                    op_sts128(row + (((cc) ^ key) << 4), w[0], w[1], w[2], w[3]);
                    op_sts128(row + (((cc + 1u) ^ key) << 4), w[4], w[5], w[6], w[7]);
                } else {
                    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)m_out * (size_t)g.ldo + c0);
                    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
            if (g.tstore) {
                fence_proxy_async_smem();      // staged rows -> visible to the TMA store the issuer warp launches once every warp has arrived
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(e_done);
            }
            tc_fence_before_sync();       // ordered before the barrier that precedes the next MMA into this buffer
        };
        // ---- phase 1 of one tile: offsets from TMEM, one record per sample (n-major: sample = n * 128 + pixel); returns this
        // thread's output pixel index ((b h + i) w + j), -1 outside the map
        auto phase1 = [&](const TC& t, uint32_t rec_s, uint32_t x_s, int it) -> int {
            const int r_org = t.i0 * s - S::ORG, k_org = t.j0 * s - S::ORG;
            const bool valid = t.i0 + di < g.h && t.j0 + dj < g.w;
            const int m_out = valid ? ((t.b * g.h + t.i0 + di) * g.w + t.j0 + dj) : -1;
            // the tile's offsets: TMEM lane = pixel, columns [0, N) row offsets, [N, 2 N) column offsets (+ bias, one rounding like
            // the stand-alone offset-conv kernels)
            mbar_wait(off_done, (uint32_t)it & 1u);
            tc_fence_after_sync();
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_off + ((uint32_t)((warp & 3) * 32) << 16), v);
            tmem_ld_wait();
            tc_fence_before_sync();      // the accumulator may be overwritten by the next tile's offset conv after barrier B
            tr(it, 1);
            float o[2 * N];
#pragma unroll
            for (int e = 0; e < 2 * N; ++e) o[e] = __uint_as_float(v[e]) + sBias[e];
            if (off_dbg != nullptr && valid && n0 == 0) {
#pragma unroll
                for (int e = 0; e < 2 * N; ++e) off_dbg[(size_t)m_out * (size_t)(2 * N) + e] = o[e];
            }
            auto make_record = [&](int n, int ri, int ki) {
                const uint32_t ra = rec_s + (uint32_t)(n * 128 + p) * 16u;
                if (!valid) {      // weights 0, corners at the tile origin: the operand row is never stored
                    op_sts128(ra, x_s, x_s, x_s, x_s);
                    op_sts128(ra + rec_g_ofs, 0u, 0u, 0u, 0u);
                    return;
                }
                float o_r = o[0], o_k = o[N];
#pragma unroll
                for (int e = 1; e < N; ++e) {
                    o_r = e == n ? o[e] : o_r;
                    o_k = e == n ? o[N + e] : o_k;
                }
                const SamplePoint q = make_point_grid(ri, ki, o_r, o_k, g.hm, g.wm);
                op_sts128(ra + rec_g_ofs, __float_as_uint(__fmul_rn(q.ar0, q.ak0)), __float_as_uint(__fmul_rn(q.ar1, q.ak1)),
                          __float_as_uint(__fmul_rn(q.ar0, q.ak1)), __float_as_uint(__fmul_rn(q.ar1, q.ak0)));
                const int t0 = q.r0 - r_org, t1 = q.r1 - r_org, u0 = q.k0 - k_org, u1 = q.k1 - k_org;
                const bool inside = (unsigned)t0 < (unsigned)S::RIN && (unsigned)t1 < (unsigned)S::RIN &&
                                    (unsigned)u0 < (unsigned)S::KIN && (unsigned)u1 < (unsigned)S::KIN;
                if (inside) {
                    // byte offset of an input pixel inside the staged tile (first channel vector), then the swizzle
                    auto rowpart = [](int tt) -> uint32_t {
                        return TS == 1 ? (uint32_t)tt * (uint32_t)(S::TWs * S::PB)
                                       : (uint32_t)(tt & 1) * S::SYB + (uint32_t)(tt >> 1) * (uint32_t)(S::TWs * S::PB);
                    };
                    auto colpart = [](int uu) -> uint32_t {
                        return TS == 1 ? (uint32_t)uu * (uint32_t)S::PB : (uint32_t)(uu & 1) * S::SXB + (uint32_t)(uu >> 1) * (uint32_t)S::PB;
                    };
                    auto sw = [&](uint32_t lin) { return x_s + (lin ^ ((lin >> 3) & S::SWM)); };
                    const uint32_t a0 = rowpart(t0), a1 = rowpart(t1), b0 = colpart(u0), b1 = colpart(u1);
                    op_sts128(ra, sw(a0 + b0), sw(a1 + b1), sw(a0 + b1), sw(a1 + b0));
                } else {      // served from global memory (L2): image-relative byte offsets, bit 31 of .x marks it
                    const int imgRowB = g.W << (cvs + 4), pixB = 16 << cvs;
                    const int a0 = q.r0 * imgRowB, a1 = q.r1 * imgRowB, b0 = q.k0 * pixB, b1 = q.k1 * pixB;
                    op_sts128(ra, (uint32_t)(a0 + b0) | 0x80000000u, (uint32_t)(a1 + b1), (uint32_t)(a0 + b1), (uint32_t)(a1 + b0));
                }
            };
            const int gr = t.i0 * s, gk = t.j0 * s;
            if (has[0]) make_record(ns[0], gr + br[0], gk + bk[0]);
            if (has[1]) make_record(ns[1], gr + br[1], gk + bk[1]);
            return m_out;
        };

        int prev_m = -1, cur_m = -1;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
            const int xb = it & g.xb_mask, ab = it & g.ab_mask;
            const TC cur = tile_coords(tile);
            tr(it, 0);
            const uint32_t rec_s = smem_s + g.ofs_rec;
            const uint32_t x_s = smem_s + g.ofs_x + (uint32_t)xb * S::X_BYTES;
            cur_m = phase1(cur, rec_s, x_s, it);
            tr(it, 2);
            op_bar_sync(1, NW);                                           // (A) records of this tile are visible
            tr(it, 3);
            mbar_wait(&x_full[xb], (uint32_t)(it >> g.xb_shift) & 1u);   // the staged input tile has landed (acquire for the generic loads)
            if (g.ab_mask == 0 && it > 0) mbar_wait(&mma_done[(it - 1) & 1], ((it - 1) >> 1) & 1);   // operand buffer free again

            tr(it, 4);
            // ---- phase 2: bilinear resampling into the swizzled operand tile; IF items in flight, loads first ------------------------
            {
                const uint32_t a_s = smem_s + g.ofs_a + (uint32_t)ab * g.a_bytes;
                const uint8_t* xg = reinterpret_cast<const uint8_t*>(x) + (size_t)cur.b * g.img_bytes + ((uint32_t)cv << 4);
                auto load_rec = [&](int sx) { return op_lds128(rec_s + (uint32_t)sx * 16u); };
                auto load_item = [&](int sx, const uint4& o, float4& gw, uint4 (&q)[4]) {
                    gw = op_lds_f4(rec_s + rec_g_ofs + (uint32_t)sx * 16u);
                    if ((int)o.x >= 0) {
                        q[0] = op_lds128((o.x ^ cx) + hoff); q[1] = op_lds128((o.y ^ cx) + hoff);
                        q[2] = op_lds128((o.z ^ cx) + hoff); q[3] = op_lds128((o.w ^ cx) + hoff);
                    } else {
                        // cold path: the sample left the staged tile, its corners come from global memory (L2); a non-unrolled
                        // rotate loop so the compiler keeps it a branch
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
                    op_sts128(a_s + (k8 >> 3) * 16384u + px * 128u + (((k8 ^ px) & 7u) << 4),
                              bilinear_bf16x2(q[0].x, q[1].x, q[2].x, q[3].x, gw), bilinear_bf16x2(q[0].y, q[1].y, q[2].y, q[3].y, gw),
                              bilinear_bf16x2(q[0].z, q[1].z, q[2].z, q[3].z, gw), bilinear_bf16x2(q[0].w, q[1].w, q[2].w, q[3].w, gw));
                };
                constexpr int IF = (65536 / ((NW + 32) * MINB) >= 120) ? 4 : 2;
                constexpr int R = (TN << TCVS) / TG;
                static_assert((TN << TCVS) % TG == 0, "items per tile must divide evenly over the threads");
                uint4 orec[IF];
                auto fetch_recs = [&](int k) {
#pragma unroll
                    for (int u = 0; u < IF; ++u)
                        if (k + u < R) orec[u] = load_rec(sx0 + (k + u) * spr);
                };
                fetch_recs(0);
#pragma unroll
                for (int k = 0; k < R; k += IF) {
                    float4 gw[IF];
                    uint4 q[IF][4];
#pragma unroll
                    for (int u = 0; u < IF; ++u)
                        if (k + u < R) load_item(sx0 + (k + u) * spr, orec[u], gw[u], q[u]);
                    fetch_recs(k + IF);      // next step's records, in flight during this step's arithmetic
#pragma unroll
                    for (int u = 0; u < IF; ++u)
                        if (k + u < R) store_item(sx0 + (k + u) * spr, gw[u], q[u]);
                }
            }
            tr(it, 5);
            fence_proxy_async_smem();      // generic-proxy stores of the operand tile -> visible to tcgen05 (async proxy)
            op_bar_sync(2, NW);            // (B) operand tile complete, input tile consumed, epilogue(it - 2) done by every warp
            if (tid == 0) mbar_arrive(b_done);
            tr(it, 6);
            if (it > 0) epilogue(prev_m, it - 1);      // the previous tile's MMAs ran during this tile's phases
            tr(it, 7);
            prev_m = cur_m;
        }
        if (it > 0) epilogue(prev_m, it - 1);
    }
    if (tr.on) tr.buf[2 * kOpTraceN] = tr.n;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tmem_dealloc(tmem_base, g.tmem_cols);
        tmem_dealloc(tmem_off, 32);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
struct OPPlan { int TG, cvs; size_t x_bytes, x_tx, kc_kb; };

static int op_shape(int C, int N, int s, OPPlan* pl)
{
    // the yolov8-LD-P2 shapes (cfg/models/yolov8-LD-P2.yaml:15-44): num_param 1 / stride 1 with 32, 64, 128 channels, num_param 3 /
    // stride 2 with 16, 32, 64 channels.  Everything else keeps ldconv_offset_conv_* + ldconv_gather_gemm_fwd.
#define LDC_OP_SHAPE(TCVS_, TS_, TG_)                                                                          \
    do {                                                                                                       \
        pl->TG = TG_; pl->cvs = TCVS_;                                                                         \
        pl->x_bytes = OPShape<TCVS_, TS_>::X_BYTES; pl->x_tx = OPShape<TCVS_, TS_>::X_TX;                      \
        pl->kc_kb = OPShape<TCVS_, TS_>::NZ_KB;                                                                \
        return 1;                                                                                              \
    } while (0)
    if (N == 1 && s == 1) {
        if (C == 32) LDC_OP_SHAPE(2, 1, 2);
        if (C == 64) LDC_OP_SHAPE(3, 1, 2);
        if (C == 128) LDC_OP_SHAPE(4, 1, 2);
    } else if (N == 3 && s == 2) {
        if (C == 16) LDC_OP_SHAPE(1, 2, 2);
        if (C == 32) LDC_OP_SHAPE(2, 2, 3);
        if (C == 64) LDC_OP_SHAPE(3, 2, 3);
    }
#undef LDC_OP_SHAPE
    return 0;
}

static int op_geometry(int B, int C, int H, int W, int N, int s, int O, int ldo, int act, OPGeom* out, OPPlan* pl, size_t* smem_bytes)
{
    if (!op_shape(C, N, s, pl)) return 0;
    if (O % 16 != 0 || O > 256 || ldo % 8 != 0 || ldo < O) return 0;
    if (s == 2 && ((H | W) & 1)) return 0;
    if (H < 2 || W < 2) return 0;
    if ((long long)H * W * C * 2 >= 0x7fffffffll) return 0;
    OPGeom g;
    g.N = N; g.H = H; g.W = W; g.O = O; g.ldo = ldo; g.act = act;
    g.h = out_size(H, s); g.w = out_size(W, s);
    g.ON = (O + 15) / 16 * 16;
    g.K = N * C;
    g.num_kb = (g.K + 63) / 64;
    g.ksteps = g.K / 16;
    g.hm = (float)(H - 1); g.wm = (float)(W - 1);
    g.img_bytes = (unsigned long long)H * W * C * 2;
    g.b_bytes = (uint32_t)g.ON * 128u;
    g.a_bytes = (uint32_t)g.num_kb * 16384u;
    g.x_bytes = (uint32_t)pl->x_bytes;
    g.x_tx_bytes = (uint32_t)pl->x_tx;
    g.tmem_cols = 32;
    while (g.tmem_cols < (uint32_t)(2 * g.ON)) g.tmem_cols <<= 1;
    const int threads = 128 * pl->TG + 32;
    const int reg_cap = threads <= 288 ? 3 : (threads <= 416 ? 2 : 1);      // CTAs per SM at <= 72 / 76 registers per thread
    const int tmem_cap = (int)(512u / (g.tmem_cols + 32u));
    // candidate plans (input-tile buffers, operand buffers): most CTAs per SM first, then the deeper buffering
    static const int cfg[4][2] = {{2, 2}, {2, 1}, {1, 2}, {1, 1}};
    int best_ctas = 0;
    OPGeom best = g;
    size_t best_smem = 0;
    const bool ts_shape = O == 16 || O == 32 || O == 64 || O == 128;      // measured: 675 vs 685 us over the nine layers with / without
    // with the TMA-store staging tile first; without it only when that costs a CTA per SM or does not fit (64 -> 128 channels)
    for (int pass = 0; pass < 2; ++pass) {
        const int ts = pass == 0 ? 1 : 0;
        if (ts && !ts_shape) continue;
        for (int ci = 0; ci < 4; ++ci) {
            OPGeom q = g;
            q.XB = cfg[ci][0]; q.AB = cfg[ci][1];
            uint32_t ofs = 0;
            q.ofs_a = ofs; ofs += (uint32_t)q.AB * q.a_bytes;
            q.ofs_b = ofs; ofs += (uint32_t)q.num_kb * q.b_bytes;
            q.ofs_bc = ofs; ofs += (uint32_t)pl->kc_kb * 2048u;
            q.ofs_x = ofs; ofs += (uint32_t)q.XB * q.x_bytes;
            q.ofs_rec = ofs; ofs += (uint32_t)(128 * N * 32);
            q.tstore = ts;
            q.st_rowb = (O < 64 ? O : 64) * 2;
            q.ofs_stage = 0;
            if (ts) {      // staging tile of the TMA-store epilogue: [64-channel block][128 pixels][row], 1024-aligned (swizzled rows)
                ofs = (ofs + 1023u) & ~1023u;
                q.ofs_stage = ofs;
                ofs += (uint32_t)((O + 63) / 64) * 128u * 128u;
            }
            q.ofs_aff = ofs; ofs += (uint32_t)q.ON * 8u;
            q.ofs_bias = ofs; ofs += 64u;
            ofs = (ofs + 7u) & ~7u;
            q.ofs_bar = ofs; ofs += 9u * 8u + 16u;
            const size_t need = (size_t)ofs + 1024;
            if (need > 225 * 1024) continue;
            int ctas = (int)((227 * 1024) / (need + 1024));
            if (ctas > reg_cap) ctas = reg_cap;
            if (ctas > tmem_cap) ctas = tmem_cap;
            if (ctas > best_ctas) { best_ctas = ctas; best = q; best_smem = need; }
        }
    }
    if (best_ctas == 0) return 0;
    g = best;
    g.per_sm = best_ctas;
    g.xb_mask = g.XB - 1; g.xb_shift = g.XB - 1; g.ab_mask = g.AB - 1;
    g.tiles_h = (g.h + kOpTH - 1) / kOpTH;
    g.tiles_w = (g.w + kOpTW - 1) / kOpTW;
    const long long nt = (long long)B * g.tiles_h * g.tiles_w;
    if (nt > 0x7fffffffll || (long long)B * g.h * g.w > 0x7fffffffll) return 0;
    g.num_tiles = (int)nt;
    const unsigned tpi = (unsigned)(g.tiles_h * g.tiles_w);
    if (nt * tpi >= 0xffffffffll) return 0;
    g.inv_img = tpi == 1 ? 0u : (unsigned)((0x100000000ull + tpi - 1) / tpi);
    g.inv_tw = g.tiles_w == 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)g.tiles_w - 1) / (unsigned)g.tiles_w);
    *smem_bytes = best_smem;
    *out = g;
    return 1;
}

int onepass_supported(int B, int C, int H, int W, int N, int s, int O, int ldo, int dtype)
{
    if (dtype != LDCONV_BF16) return 0;
    OPGeom g;
    OPPlan pl;
    size_t smem;
    return op_geometry(B, C, H, W, N, s, O, ldo, LDCONV_ACT_SILU, &g, &pl, &smem);
}

int onepass_fwd(const void* x, const void* w_conv, const float* bias, const int* pn, const void* wt, const float* scale,
                const float* shift, void* out, int ldo, float* off_dbg, int B, int C, int H, int W, int N, int s, int O, int act,
                cudaStream_t st)
{
    OPGeom g;
    OPPlan pl;
    size_t smem;
    if (!op_geometry(B, C, H, W, N, s, O, ldo, act, &g, &pl, &smem))
        return fail(LDCONV_E_ARG, "one-pass LDConv kernel: shape not covered (C=%d N=%d s=%d O=%d ldo=%d H=%d W=%d)", C, N, s, O, ldo, H, W);
    if (!aligned16(x) || !aligned16(wt) || !aligned16(w_conv) || !aligned16(out))
        return fail(LDCONV_E_ALIGN, "one-pass LDConv kernel: x / weights / out must be 16-byte aligned");
    CUtensorMap tmX, tmW, tmWc, tmO;
    if (s == 1) {
        const int cb = C >= 64 ? 64 : C;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)(kOpTW + 5), (cuuint32_t)(kOpTH + 5), 1};
        const CUtensorMapSwizzle sw = cb * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gdim, gstr, box, sw)) return e;
    } else {
        // x viewed as (2C [sx, c], sy, W/2, H/2, B): the innermost two dimensions make pixel' = [sy][sx][c]
        const int pb = 4 * C > 128 ? 128 : 4 * C;
        cuuint64_t gdim[5] = {(cuuint64_t)(2 * C), 2, (cuuint64_t)(W / 2), (cuuint64_t)(H / 2), (cuuint64_t)B};
        cuuint64_t gstr[4] = {(cuuint64_t)W * C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[5] = {(cuuint32_t)(pb / 2), 1, (cuuint32_t)(kOpTW + 2), (cuuint32_t)(kOpTH + 2), 1};
        if (int e = encode_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, gdim, gstr, box,
                               pb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)g.K, (cuuint64_t)O};
        cuuint64_t gstr[1] = {(cuuint64_t)g.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)g.ON};
        if (int e = encode_map(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    {
        const int KC = s == 1 ? 9 * C : 16 * C;
        cuuint64_t gdim[2] = {(cuuint64_t)KC, (cuuint64_t)(2 * N)};
        cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
        cuuint32_t box[2] = {64, 16};
        if (int e = encode_map(&tmWc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_conv, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    tmO = tmW;      // placeholder when the epilogue stores directly
    if (g.tstore) {      // out (B, h, w, O | ldo): box = (one row of <= 64 channels, 8 columns, 16 rows) = the staging tile's order
        const int cb = O < 64 ? O : 64;
        cuuint64_t gdim[4] = {(cuuint64_t)O, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)g.w * ldo * 2, (cuuint64_t)g.h * g.w * ldo * 2};
        cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)kOpTW, (cuuint32_t)kOpTH, 1};
        const CUtensorMapSwizzle sw = cb * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        if (int e = encode_map(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gdim, gstr, box, sw)) return e;
    }
    using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, const __nv_bfloat16*, const float*, const int*, const float*,
                          const float*, __nv_bfloat16*, float*, long long*, OPGeom);
    Kern kern = nullptr;
    const int key = ((N * 10 + pl.cvs) * 10 + s) * 10 + g.per_sm;
    switch (key) {
#define LDC_OP_CASE(N_, CVS_, S_, TG_)                                                         \
        case ((N_ * 10 + CVS_) * 10 + S_) * 10 + 1: kern = ldconv_onepass_kernel<N_, CVS_, S_, TG_, 1>; break; \
        case ((N_ * 10 + CVS_) * 10 + S_) * 10 + 2: kern = ldconv_onepass_kernel<N_, CVS_, S_, TG_, 2>; break; \
        case ((N_ * 10 + CVS_) * 10 + S_) * 10 + 3: kern = ldconv_onepass_kernel<N_, CVS_, S_, TG_, 3>; break;
        LDC_OP_CASE(1, 2, 1, 2)
        LDC_OP_CASE(1, 3, 1, 2)
        LDC_OP_CASE(1, 4, 1, 2)
        LDC_OP_CASE(3, 1, 2, 2)
#undef LDC_OP_CASE
        case 3221: kern = ldconv_onepass_kernel<3, 2, 2, 3, 1>; break;
        case 3222: kern = ldconv_onepass_kernel<3, 2, 2, 3, 2>; break;
        case 3321: kern = ldconv_onepass_kernel<3, 3, 2, 3, 1>; break;
        case 3322: kern = ldconv_onepass_kernel<3, 3, 2, 3, 2>; break;
        default: return fail(LDCONV_E_ARG, "one-pass LDConv kernel: no instance for key %d", key);
    }
    LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = num_sms() * g.per_sm;
    if (grid > g.num_tiles) grid = g.num_tiles;
    LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(128 * pl.TG + 32), smem, st, tmX, tmW, tmWc, tmO, (const __nv_bfloat16*)x, bias, pn, scale,
                        shift, (__nv_bfloat16*)out, off_dbg, g_op_trace, g));
    g_op_trace = nullptr;
    LDC_LAUNCH_CHECK("ldconv_onepass_kernel");
    set_impl(LDCONV_IMPL_TCGEN05);
    return LDCONV_OK;
}

}  // namespace ldc

// debug: the NEXT ldconv_onepass_fwd call of this thread writes its timeline into `device_buf` (2 x (2 * 64 + 2) long long: per
// role (worker thread 0, issuer thread) pairs (iteration * 100 + tag, clock64) and the pair count)
LDC_API int ldconv_debug_onepass_trace(void* device_buf)
{
    ldc::g_op_trace = (long long*)device_buf;
    return LDCONV_OK;
}

LDC_API int ldconv_onepass_supported(int B, int C, int H, int W, int N, int s, int O, int ldo, int dtype)
{
    return ldc::onepass_supported(B, C, H, W, N, s, O, ldo, dtype);
}

LDC_API int ldconv_onepass_fwd(const void* x, const void* w_offconv, const float* b_off, const int32_t* p_n, const void* wt,
                               const float* scale, const float* shift, void* out, int ldo, float* off_out, int B, int C, int H,
                               int W, int N, int s, int O, int act, int dtype, void* stream)
{
    using namespace ldc;
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_onepass_fwd: bf16 only");
    LDC_REQUIRE(x && w_offconv && p_n && wt && out, "ldconv_onepass_fwd: null pointer");
    LDC_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1 && N >= 1 && s >= 1 && O >= 1, "ldconv_onepass_fwd: bad dims");
    if (B == 0) return LDCONV_OK;
    return onepass_fwd(x, w_offconv, b_off, p_n, wt, scale, shift, out, ldo, off_out, B, C, H, W, N, s, O, act, (cudaStream_t)stream);
}
