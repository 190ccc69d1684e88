// ldconv_umma.cu -- tcgen05 / TMEM / TMA GEMM of the LDConv (N,1) conv for bf16 (sm_100a).
//
//   pre(M,O) = operand(M,K) . wt(O,K)^T   followed by the folded-BatchNorm affine and SiLU
//   replaces nn.Conv2d(inc, outc, (N,1), (N,1)) + BatchNorm2d + SiLU of
//   /root/reference/ultralytics/nn/modules/conv.py:355,408
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer: 128 x 64 bf16 operand blocks and O x 64 weight blocks into a 128B-swizzled smem ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=O, K=16) into a double-buffered TMEM accumulator
//   warps 2..5  epilogue: tcgen05.ld accumulator rows -> affine + SiLU -> bf16 -> 16-byte global stores
// For the YAML's shapes (K <= 192, O <= 128) the GEMM is HBM-bound (19-77 flop/B), so what matters
// is the number of operand bytes in flight per SM (the smem ring) rather than MMA issue rate.
#include "common.cuh"
#include "umma.cuh"
#include "tmap.cuh"

namespace ldc {

using namespace umma;

int col_stats_bf16(const __nv_bfloat16* pre, long long M, int O, double* sum, double* sqsum, cudaStream_t st);

static constexpr int kTileM = 128;
static constexpr int kBlockK = 64;                      // bf16 elements = 128 bytes = one swizzle row
static constexpr int kABytes = kTileM * kBlockK * 2;    // 16 KiB
static constexpr int kGemmThreads = 320;          // TMA warp, MMA warp, 8 epilogue warps

static thread_local int g_force_ffma = 0;   // ldconv_set_flag(LDCONV_FLAG_FORCE_FFMA, v)
static int g_env_force_ffma = -1;           // environment LDCONV_FORCE_FFMA=1 (debug A/B switch, wins over the flag)

// 2-D bf16 row-major (rows, cols) tensor, box (box_rows, 64 cols), 128-byte swizzle, zero fill out of bounds
static int make_map_2d(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, long long ld)
{
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    return encode_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// Epilogue of one 16-column chunk of one accumulator row: raw accumulator -> `pre`, affine + activation (+ residual) -> `out`.
// Optional second destination of the activated output: the channel window [c_lo, c_lo + c_n) (multiples of 16) is ALSO stored
// densely at out2 + m * ld2.  Lets a C2f's first 1x1 conv write its concat buffer and, in the same pass, a dense copy of the half
// the first Bottleneck reads: a 16-channel slice of a 48-channel NHWC buffer costs the whole buffer in DRAM traffic.
struct GemmOut2 {
    __nv_bfloat16* ptr;
    int ld, c_lo, c_n;
};

// Optional "maximum with two coarser maps" stage of the epilogue (SSFF ScalSeq, nn/extra_modules/block.py:3414-3443: the three
// pyramid levels go through the same point-wise Conv3d + BatchNorm3d + LeakyReLU, the coarser two are up-sampled with `nearest` and
// MaxPool3d((3,1,1)) takes the maximum over the levels): the finest level's GEMM takes the other two levels' outputs z1 (B,H1,W1,O)
// and z2 (B,H2,W2,O), dense bf16, and writes max(bf16(act(.)), z1[up], z2[up]) (+ residual = the following Add layer) -- the finest
// level's own (B,H,W,O) map never makes the round trip through HBM.
struct GemmMaxUp {
    const __nv_bfloat16 *z1, *z2;
    int H, W, H1, W1, H2, W2;
};

// Pixel packing for narrow 1x1 convs: P consecutive pixels form ONE operand row (P * Cin wide, contiguous because x is dense) and
// the weights are block-diagonal (P * Cout, P * Cin), so a row of the product holds the Cout outputs of P consecutive pixels.  The
// kernel sees an ordinary (rows / P) x (P Cin) x (P Cout) GEMM; only the epilogue's addressing knows: 16-column chunk c0 of row m is
// channel c0 mod Cout of pixel m P + c0 / Cout (Cout = 1 << o_shift).  The fixed cost per 128-row tile (~1.8 us per CTA) bounds the
// P2-resolution GEMMs of the model (12.8 K tiles of 16 KB), not their bytes; packing divides the tiles by P.
struct GemmPack {
    int P, o_shift;
};

// Detect head in the epilogue (nn/modules/head.py:55-77): the LAST 1x1 conv of a branch decodes its own accumulator rows instead of
// writing logits that a decode kernel reads back.  mode 1 (box branch, Cout = 4 x 16 DFL bins): per side softmax-expectation over
// the 16 bins (DFL, nn/modules/block.py:37-56), dist2bbox around the cell-centre anchor (utils/tal.py:309-319), times the level
// stride -> rows 0..3 of y (B, 4+nc, total) at column a0 + pixel.  mode 2 (class branch, Cout = nc <= 16): sigmoid -> rows 4..4+nc.
// The logits are rounded to bf16 first, exactly like the plain epilogue's store, and the arithmetic is detect_decode_kernel's: y is
// bit-identical to conv1x1 -> ldconv_detect_decode, minus a 275 MB round trip of logits per 64 images.
struct GemmDetect {
    __nv_bfloat16* y;
    int mode, H, W, nc, a0, total;
    float stride;
};

__device__ __forceinline__ void detect_logits16(const uint32_t (&v)[16], uint32_t sc_addr, uint32_t sh_addr, float (&z)[16])
{
    affine_act16(v, sc_addr, sh_addr, LDCONV_ACT_NONE, z);
#pragma unroll
    for (int e = 0; e < 16; ++e) z[e] = __bfloat162float(__float2bfloat16_rn(z[e]));      // what the unfused path stores and reloads
}

__device__ __forceinline__ float dfl_expectation16(const float (&v)[16])
{
    float mx = v[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) mx = fmaxf(mx, v[k]);
    float den = 0.f, num = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float e = __expf(v[k] - mx);
        den += e;
        num = fmaf(e, (float)k, num);
    }
    return num / den;
}

// aff_s = shared-space address of the scale of column c0, aff_s + sh_ofs = of its shift (explicit LDS: the generic loads this replaces were
// the top stall of the kernel, profiles/r1_ncu_gemmL1b.txt).
__device__ __forceinline__ void gemm_epilogue_chunk(const uint32_t (&v)[16], long long m, int c0, int O, uint32_t aff_s, int act,
                                                    __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ pre,
                                                    const __nv_bfloat16* __restrict__ residual, int ldo, int ldr, bool vec_store,
                                                    const GemmOut2& o2, uint32_t sh_ofs, const __nv_bfloat16* mx1 = nullptr,
                                                    const __nv_bfloat16* mx2 = nullptr, GemmPack pk = GemmPack{1, 0},
                                                    uint32_t stage_row_s = 0, uint32_t stage_swz = 0, const float4* pre_sc = nullptr,
                                                    const float4* pre_sh = nullptr)
{
    const bool full16 = vec_store && (c0 + 16 <= O);
    if (pre) {
        __nv_bfloat16* dst = pre + m * ldo + c0;
        if (full16) {
            float lo[8], hi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { lo[e] = __uint_as_float(v[e]); hi[e] = __uint_as_float(v[8 + e]); }
            Vec16<__nv_bfloat16>::store(dst, lo);
            Vec16<__nv_bfloat16>::store(dst + 8, hi);
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                if (c0 + e < O) dst[e] = __float2bfloat16_rn(__uint_as_float(v[e]));
        }
    }
    if (out) {
        float z[16];
        if (pre_sc) {      // scales / shifts already in registers (requested before the caller's tcgen05.wait::ld)
            const float4 (&scr)[4] = *reinterpret_cast<const float4 (*)[4]>(pre_sc);
            const float4 (&shr)[4] = *reinterpret_cast<const float4 (*)[4]>(pre_sh);
            affine_act16_r(v, scr, shr, act, z);
        } else {
            affine_act16(v, aff_s, aff_s + sh_ofs, act, z);
        }
        if (mx1) {      // host: O % 16 == 0, z1 / z2 dense and 16-byte aligned.  The level's own output is a bf16 tensor in the reference
            float a1[8], a2[8], b1[8], b2[8];
            Vec16<__nv_bfloat16>::load(mx1 + c0, a1);
            Vec16<__nv_bfloat16>::load(mx1 + c0 + 8, a2);
            Vec16<__nv_bfloat16>::load(mx2 + c0, b1);
            Vec16<__nv_bfloat16>::load(mx2 + c0 + 8, b2);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                z[e] = fmaxf(fmaxf(__bfloat162float(__float2bfloat16_rn(z[e])), a1[e]), b1[e]);
                z[8 + e] = fmaxf(fmaxf(__bfloat162float(__float2bfloat16_rn(z[8 + e])), a2[e]), b2[e]);
            }
        }
        if (residual) {
            const __nv_bfloat16* rp = residual + m * ldr + c0;
            if (full16 && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
                float r1[8], r2[8];
                Vec16<__nv_bfloat16>::load(rp, r1);
                Vec16<__nv_bfloat16>::load(rp + 8, r2);
#pragma unroll
                for (int e = 0; e < 8; ++e) { z[e] += r1[e]; z[8 + e] += r2[e]; }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (c0 + e < O) z[e] += __bfloat162float(rp[e]);
            }
        }
        __nv_bfloat16* dst = out + m * ldo + c0;
        if (pk.P > 1) dst = out + (m * pk.P + (c0 >> pk.o_shift)) * ldo + (c0 & ((1 << pk.o_shift) - 1));
        if (stage_row_s) {      // warp-staged store (see the kernel): this row's two 16-byte pieces go to the warp's scratch, swizzled
            uint4 lo, hi;
            pack16_bf16(z, lo, hi);
            const uint32_t ch = (uint32_t)c0 >> 3;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row_s + ((ch ^ stage_swz) << 4)), "r"(lo.x), "r"(lo.y),
                         "r"(lo.z), "r"(lo.w) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row_s + (((ch + 1) ^ stage_swz) << 4)), "r"(hi.x),
                         "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
        } else if (full16) {
            uint4 lo, hi;
            pack16_bf16(z, lo, hi);
            reinterpret_cast<uint4*>(dst)[0] = lo;
            reinterpret_cast<uint4*>(dst)[1] = hi;
            if (o2.ptr && (unsigned)(c0 - o2.c_lo) < (unsigned)o2.c_n) {      // host: window 16-aligned, out2 16-byte aligned
                uint4* d2 = reinterpret_cast<uint4*>(o2.ptr + m * o2.ld + (c0 - o2.c_lo));
                d2[0] = lo;
                d2[1] = hi;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                if (c0 + e < O) dst[e] = __float2bfloat16_rn(z[e]);
        }
    }
}

// Pipeline depth.  The accumulator lives in one of NB TMEM buffers (NB * ON <= 256 columns, so two CTAs share an SM's 512):
// with two buffers the dependency cycle MMA(t) -> commit -> epilogue(t) -> tempty -> MMA(t+2) costs ~2700 cycles per buffer
// and bounded the kernel at ~1350 cycles per tile whatever the tile did (scripts/gemm_dbg.sh: 31 us of pure hand-offs at
// K=48, O=32, M=1.6M, and operand traffic / epilogue added on top instead of overlapping).  With up to eight buffers the MMA
// warp runs ahead, and the eight epilogue warps work as two groups of four on alternate tiles, so two epilogues are in flight.
template <bool MAXUP>
__global__ void __launch_bounds__(kGemmThreads, 2)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                 __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ residual, int M, int O, int ON, int num_kb,
                 int num_tiles, int stages, int act, uint32_t tmem_cols, int vec_store, int ldo, int ldr, int NB, int dbg,
                 GemmOut2 o2, GemmMaxUp mu, GemmPack pk, int stage_rb, GemmDetect det)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS, not generic LD / ST
    const int b_bytes = ON * kBlockK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)stages * kABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)stages * b_bytes);
    uint64_t* empty = full + stages;
    uint64_t* tfull = empty + stages;         // [8]
    uint64_t* tempty = tfull + 8;             // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 8);
    float* s_affine = reinterpret_cast<float*>(tmem_slot + 4);      // [0, ON) scale, [ON, 2 ON) shift; halved for SiLU (affine_act16)
    // Warp-staged output (stage_rb = bytes of an output row, 64 or 128; 0 = off): a thread owns a row of the accumulator, so its
    // 16-byte global stores touched 32 cache lines per warp instruction -- one L1 wavefront per 16 bytes, 1024 per tile at O = 64,
    // the busiest unit of the kernel (profiles/r1_ncu_gemm_64_64_p2_s4.txt: l1tex 65 %).  Each epilogue warp instead parks its
    // 32 rows in 32 x stage_rb bytes of shared memory (16-byte pieces XOR-swizzled: conflict-free both ways) and writes them out
    // with row-contiguous 16-byte stores: 4 (8) lines per instruction.
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_affine + 2 * ON);
    s_stage += (16u - (smem_u32(s_stage) & 15u)) & 15u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nb_shift = NB == 8 ? 3 : (NB == 4 ? 2 : (NB == 2 ? 1 : 0));

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    pdl_wait();       // everything below may read what the previous kernel wrote (operand, scale / shift)
    {
        const float aff_pre = affine_half_for(act);
        for (int o = threadIdx.x; o < ON; o += blockDim.x) {
            s_affine[o] = aff_pre * ((scale && o < O) ? scale[o] : 1.f);
            s_affine[ON + o] = aff_pre * ((shift && o < O) ? shift[o] : 0.f);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // The two control warps run their loops WARP-UNIFORMLY and one elected lane issues: under `if (lane == 0)` the compiler wraps
    // every uniform-datapath instruction (UTMALDG, UTCHMMA, the commits) in an ELECT / BRA.U.ANY loop over the active lanes, ~10
    // dependent instructions per MMA on the one thread the whole CTA waits for (benchmarks/trace_onepass.py, round 2).
    if (warp == 0) {
        const bool leader = elect_one();
        {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (leader) {
                        if (dbg & 1) {          // experiment: no operand traffic (results are garbage)
                            mbar_arrive(&full[s]);
                        } else {
                            mbar_arrive_expect_tx(&full[s], (uint32_t)(kABytes + b_bytes));
                            tma_load_2d(sA + (size_t)s * kABytes, &tmA, &full[s], kb * kBlockK, tile * kTileM);
                            tma_load_2d(sB + (size_t)s * b_bytes, &tmB, &full[s], kb * kBlockK, 0);
                        }
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // (three issuing warps on alternate tiles were tried: the pure hand-off time fell from 31 to 20 us at layer 1, the full
        // kernel did not move -- it is bound by the TMA box-row rate of the operand loads, ~1 row per 10 cycles per SM whatever
        // the row width, scripts/gemm_dbg.sh -- and sharing the stage ring between issuers needs per-issuer sub-rings)
        const bool leader = elect_one();
        {
            const uint32_t idesc = make_idesc_bf16(kTileM, ON);
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int buf = it & (NB - 1);
                const uint32_t tph = (uint32_t)(it >> nb_shift) & 1u;
                mbar_wait(&tempty[buf], tph ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ON);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(sA + (size_t)s * kABytes);
                    const uint32_t b_addr = smem_u32(sB + (size_t)s * b_bytes);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            mma_bf16_ss(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc,
                                        (uint32_t)((kb | k) != 0));
                        }
                        mma_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                if (leader) mma_commit(&tfull[buf]);
                __syncwarp();
            }
        }
    } else {
        const int lg = warp & 3;              // TMEM lane group this warp may access (warps 2..9)
        const int grp = (warp - 2) >> 2;      // warps 2..5 take the even tiles of this CTA, warps 6..9 the odd ones
        const int chunks = ON / 16;
        const uint32_t aff_s = smem_u32(s_affine);
        for (int it = grp; (long long)blockIdx.x + (long long)it * gridDim.x < num_tiles; it += 2) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int buf = it & (NB - 1);
            const uint32_t tph = (uint32_t)(it >> nb_shift) & 1u;
            mbar_wait(&tfull[buf], tph);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * ON);
            const long long m = (long long)tile * kTileM + lg * 32 + lane;
            const uint32_t stage_s = stage_rb ? smem_u32(s_stage) + (uint32_t)((warp - 2) * 32 * stage_rb) : 0u;
            const uint32_t stage_row_s = stage_rb ? stage_s + (uint32_t)(lane * stage_rb) : 0u;
            const uint32_t stage_swz = stage_rb == 128 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
            const __nv_bfloat16 *mx1 = nullptr, *mx2 = nullptr;
            if (MAXUP && m < M) {        // torch `nearest`: src = floor(dst * in / out)
                const int mi = (int)m, j = mi % mu.W, bi = mi / mu.W, i = bi % mu.H, b = bi / mu.H;
                const int i1 = (int)(((long long)i * mu.H1) / mu.H), j1 = (int)(((long long)j * mu.W1) / mu.W);
                const int i2 = (int)(((long long)i * mu.H2) / mu.H), j2 = (int)(((long long)j * mu.W2) / mu.W);
                mx1 = mu.z1 + (((long long)b * mu.H1 + i1) * mu.W1 + j1) * O;
                mx2 = mu.z2 + (((long long)b * mu.H2 + i2) * mu.W2 + j2) * O;
            }
            if (det.mode) {
                // ---- Detect decode in the epilogue (see GemmDetect) ----------------------------------------------------------------
                const uint32_t sh_ofs = (uint32_t)ON * 4u;
                const long long per_img = (long long)det.H * det.W;
                const int b = (int)(m / per_img), a = (int)(m - (long long)b * per_img);
                __nv_bfloat16* yp = det.y + (size_t)b * (size_t)(4 + det.nc) * det.total + det.a0 + a;
                if (det.mode == 1) {
                    float d[4];
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        uint32_t v0[16], v1[16];
                        tmem_ld_32x32b_x16(taddr + (uint32_t)pr * 32u, v0);
                        tmem_ld_32x32b_x16(taddr + (uint32_t)pr * 32u + 16u, v1);
                        tmem_ld_wait();
                        float z[16];
                        detect_logits16(v0, aff_s + (uint32_t)(2 * pr) * 64u, aff_s + (uint32_t)(2 * pr) * 64u + sh_ofs, z);
                        d[2 * pr] = dfl_expectation16(z);
                        detect_logits16(v1, aff_s + (uint32_t)(2 * pr + 1) * 64u, aff_s + (uint32_t)(2 * pr + 1) * 64u + sh_ofs, z);
                        d[2 * pr + 1] = dfl_expectation16(z);
                    }
                    if (m < M) {
                        const int ai = a / det.W, aj = a - ai * det.W;
                        const float ax = (float)aj + 0.5f, ay = (float)ai + 0.5f;
                        const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
                        yp[0] = __float2bfloat16_rn(0.5f * (x1 + x2) * det.stride);
                        yp[(size_t)det.total] = __float2bfloat16_rn(0.5f * (y1 + y2) * det.stride);
                        yp[2 * (size_t)det.total] = __float2bfloat16_rn((x2 - x1) * det.stride);
                        yp[3 * (size_t)det.total] = __float2bfloat16_rn((y2 - y1) * det.stride);
                    }
                } else {
                    uint32_t v0[16];
                    tmem_ld_32x32b_x16(taddr, v0);
                    tmem_ld_wait();
                    float z[16];
                    detect_logits16(v0, aff_s, aff_s + sh_ofs, z);
                    if (m < M) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (c < det.nc) yp[(size_t)(4 + c) * det.total] = __float2bfloat16_rn(1.f / (1.f + __expf(-z[c])));
                    }
                }
            } else
            for (int ch = 0; ch < chunks; ch += 2) {
                const bool two = ch + 1 < chunks;
                float4 sc0[4], sh0[4];      // first chunk's scales / shifts: requested before the TMEM loads are waited for
                affine_load16(aff_s + (uint32_t)ch * 64u, aff_s + (uint32_t)ch * 64u + (uint32_t)ON * 4u, sc0, sh0);
                uint32_t v0[16], v1[16];
                tmem_ld_32x32b_x16(taddr + (uint32_t)ch * 16u, v0);
                if (two) tmem_ld_32x32b_x16(taddr + (uint32_t)(ch + 1) * 16u, v1);
                tmem_ld_wait();
                if (m < M && !(dbg & 2)) {
                    if (ch * 16 < O)
                        gemm_epilogue_chunk(v0, m, ch * 16, O, aff_s + (uint32_t)ch * 64u, act, out, pre, residual, ldo, ldr,
                                            vec_store != 0, o2, (uint32_t)ON * 4u, mx1, mx2, pk, stage_row_s, stage_swz, sc0, sh0);
                    if (two && (ch + 1) * 16 < O)
                        gemm_epilogue_chunk(v1, m, (ch + 1) * 16, O, aff_s + (uint32_t)(ch + 1) * 64u, act, out, pre, residual,
                                            ldo, ldr, vec_store != 0, o2, (uint32_t)ON * 4u, mx1, mx2, pk, stage_row_s, stage_swz);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
            if (stage_rb) {      // write the warp's 32 rows out: piece q = (row, 16-byte chunk), a warp instruction covers whole rows
                const int cpr = stage_rb >> 4;                      // 16-byte pieces per row: 4 or 8
                const long long m0 = (long long)tile * kTileM + lg * 32;
#pragma unroll 4
                for (int q = lane; q < 32 * cpr; q += 32) {
                    const int row = q / cpr, ch = q - row * cpr;
                    const uint32_t swz = stage_rb == 128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "r"(stage_s + (uint32_t)(row * stage_rb) + (((uint32_t)ch ^ swz) << 4)));
                    if (m0 + row < M) *reinterpret_cast<uint4*>(out + (m0 + row) * ldo + ch * 8) = v;
                }
                __syncwarp();      // the scratch is rewritten by this warp's next tile
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

static int gemm_dbg()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("LDCONV_GEMM_DBG"); v = e ? atoi(e) : 0; }
    return v;
}

int umma_gemm_supported(int M, int K, int O, int dtype, const void* a, const void* wt, const void* out, const void* pre)
{
    if (g_env_force_ffma < 0) {
        const char* e = getenv("LDCONV_FORCE_FFMA");
        g_env_force_ffma = (e && e[0] == '1') ? 1 : 0;
    }
    if (g_env_force_ffma || g_force_ffma) return 0;
    if (dtype != LDCONV_BF16) return 0;
    if (M < 1 || K % 8 != 0 || O > 256 || O < 1) return 0;
    if (!aligned16(a) || !aligned16(wt)) return 0;
    (void)out; (void)pre;
    return 1;
}

int umma_set_force_ffma(int v)
{
    g_force_ffma = v ? 1 : 0;
    return LDCONV_OK;
}

int umma_gemm_fwd_ld(const void* a, int lda, const void* wt, const float* scale, const float* shift, void* out, void* pre,
                     const void* residual, int ldr, int ldo, double* stat_sum, double* stat_sqsum, int M, int K, int O, int act,
                     cudaStream_t st);
static thread_local GemmOut2 g_out2 = {nullptr, 0, 0, 0};      // snapshotted and cleared at the entry of the next umma_gemm_fwd_ld of this thread
static thread_local GemmMaxUp g_maxup = {nullptr, nullptr, 0, 0, 0, 0, 0, 0};      // likewise
static thread_local GemmPack g_pack = {1, 0};
static thread_local GemmDetect g_detect = {nullptr, 0, 0, 0, 0, 0, 0, 0.f};                                       // likewise

int umma_gemm_fwd(const void* a, const void* wt, const float* scale, const float* shift, void* out, void* pre,
                  double* stat_sum, double* stat_sqsum, int M, int K, int O, int act, cudaStream_t st)
{
    return umma_gemm_fwd_ld(a, K, wt, scale, shift, out, pre, nullptr, 0, O, stat_sum, stat_sqsum, M, K, O, act, st);
}

int umma_gemm_fwd_ld(const void* a, int lda, const void* wt, const float* scale, const float* shift, void* out, void* pre,
                     const void* residual, int ldr, int ldo, double* stat_sum, double* stat_sqsum, int M, int K, int O, int act,
                     cudaStream_t st)
{
    // The epilogue extensions arrive through thread-local side channels set by the C entry points right before this call: they are
    // snapshotted AND cleared here, before any return path, so a failed call can never leak a stale second-output pointer / packing
    // factor into the next, unrelated GEMM of this thread.
    const GemmOut2 o2 = g_out2;
    const GemmMaxUp mu = g_maxup;
    const GemmPack pk = g_pack;
    g_out2 = GemmOut2{nullptr, 0, 0, 0};
    g_maxup = GemmMaxUp{nullptr, nullptr, 0, 0, 0, 0, 0, 0};
    g_pack = GemmPack{1, 0};
    const GemmDetect det = g_detect;
    g_detect = GemmDetect{nullptr, 0, 0, 0, 0, 0, 0, 0.f};
    if (det.mode && (out || pre || residual || o2.ptr || mu.z1 || pk.P != 1 || stat_sum))
        return fail(LDCONV_E_ARG, "tcgen05 GEMM: the Detect-decode epilogue takes no other output");
    if (stat_sum && !pre) return fail(LDCONV_E_ARG, "tcgen05 GEMM: batch statistics need the `pre` output");
    const int ON = (O + 15) / 16 * 16;
    const int num_kb = (K + kBlockK - 1) / kBlockK;
    const int num_tiles = (M + kTileM - 1) / kTileM;
    const int b_bytes = ON * kBlockK * 2;
    // two CTAs per SM when both accumulator pairs fit TMEM (2 x 2 x ON <= 512 columns): ~100 KB of smem ring each;
    // one CTA with the whole ~200 KB otherwise
    const bool two_per_sm = 4 * ON <= 512;
    // accumulator buffers: as many as fit 256 TMEM columns (two CTAs per SM), at most 8, a power of two
    int NB = 8;
    while (NB > 2 && NB * ON > 256) NB >>= 1;
    // warp-staged stores: plain outputs (no pre / residual / second output / max-up / packing) with rows of 64 or 128 bytes
    int stage_rb = 0;
    if (out && !pre && !residual && !o2.ptr && !mu.z1 && pk.P == 1 && (O == 32 || O == 64) && ldo % 8 == 0 &&
        aligned16(out))
        stage_rb = O * 2;
    const int stage_bytes = stage_rb ? 8 * 32 * stage_rb + 16 : 0;
    int stages = ((two_per_sm ? 100 : 200) * 1024 - stage_bytes) / (kABytes + b_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(LDCONV_E_ARG, "tcgen05 GEMM: tile does not fit shared memory (O=%d)", O);
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(NB * ON)) tmem_cols <<= 1;
    const size_t smem = 1024 + (size_t)stages * (kABytes + b_bytes) + (2 * stages + 16) * sizeof(uint64_t) + 16 +
                        (size_t)ON * sizeof(float2) + (size_t)stage_bytes;

    CUtensorMap tmA, tmB;
    if (int e = make_map_2d(&tmA, a, M, K, kTileM, lda)) return e;
    if (int e = make_map_2d(&tmB, wt, O, K, ON, K)) return e;

    if (pk.P > 1 && (mu.z1 || o2.ptr || residual || pre || stat_sum))
        return fail(LDCONV_E_ARG, "tcgen05 GEMM: pixel packing takes no residual / second output / statistics");
    auto kern = mu.z1 ? umma_gemm_kernel<true> : umma_gemm_kernel<false>;
    LDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = num_sms() * (two_per_sm ? 2 : 1);
    if (grid > num_tiles) grid = num_tiles;
    const int vec_store = (O % 8 == 0) && (ldo % 8 == 0) && (!out || aligned16(out)) && (!pre || aligned16(pre));
    if (o2.ptr && (!vec_store || !out || O % 16 != 0))
        return fail(LDCONV_E_ARG, "tcgen05 GEMM: the second output needs 16-byte stores and Cout %% 16 == 0");
    if (mu.z1 && (!vec_store || !out || pre || O % 16 != 0))
        return fail(LDCONV_E_ARG, "tcgen05 GEMM: the max-with-coarser-levels epilogue needs 16-byte stores and Cout %% 16 == 0");
    LDC_CUDA(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), smem, st, tmA, tmB, scale, shift, (__nv_bfloat16*)out,
                        (__nv_bfloat16*)pre, (const __nv_bfloat16*)residual, M, O, ON, num_kb, num_tiles, stages, act, tmem_cols,
                        vec_store, ldo, ldr, NB, gemm_dbg(), o2, mu, pk, stage_rb, det));
    LDC_LAUNCH_CHECK("umma_gemm_kernel");
    set_impl(LDCONV_IMPL_TCGEN05);
    if (stat_sum) {
        if (ldo != O) return fail(LDCONV_E_ARG, "tcgen05 GEMM: batch statistics need a dense `pre`");
        return col_stats_bf16((const __nv_bfloat16*)pre, M, O, stat_sum, stat_sqsum, st);
    }
    return LDCONV_OK;
}

}  // namespace ldc

using namespace ldc;

// The same 1x1 `Conv` block with a second, dense destination for the output channels [c2_lo, c2_lo + c2_n) (multiples of 16):
// out2 (rows, c2_n | ld2).  Used by the C2f executor (nn/modules/block.py:209-232: y = cv1(x).chunk(2); the second chunk feeds the
// first Bottleneck) so that the Bottleneck reads a dense tensor instead of a channel slice of the concat buffer.
LDC_API int ldconv_conv1x1_bn_act_fwd2(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                                       const void* residual, int ldr, void* out, int ldo, void* out2, int ld2, int c2_lo, int c2_n,
                                       long long rows, int Cin, int Cout, int act, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_conv1x1_bn_act_fwd2: bf16 only");
    LDC_REQUIRE(x && wt && out && out2 && rows >= 0 && rows < (1ll << 31), "ldconv_conv1x1_bn_act_fwd2: bad arguments");
    LDC_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && ldx >= Cin && ldo >= Cout && Cout <= 256 && Cout % 16 == 0,
                "ldconv_conv1x1_bn_act_fwd2: needs Cin %% 8 == 0, ldx %% 8 == 0, Cout %% 16 == 0, Cout <= 256");
    LDC_REQUIRE(c2_lo >= 0 && c2_n >= 16 && c2_lo % 16 == 0 && c2_n % 16 == 0 && c2_lo + c2_n <= Cout && ld2 >= c2_n && ld2 % 8 == 0,
                "ldconv_conv1x1_bn_act_fwd2: the second output's channel window must be 16-aligned and inside [0, Cout)");
    LDC_REQUIRE(aligned16(x) && aligned16(wt) && aligned16(out) && aligned16(out2), "ldconv_conv1x1_bn_act_fwd2: 16-byte alignment");
    if (rows == 0) return LDCONV_OK;
    g_out2 = GemmOut2{(__nv_bfloat16*)out2, ld2, c2_lo, c2_n};
    return umma_gemm_fwd_ld(x, ldx, wt, scale, shift, out, nullptr, residual, ldr, ldo, nullptr, nullptr, (int)rows, Cin, Cout,
                            act, (cudaStream_t)stream);
}

// The point-wise block of the finest SSFF level fused with the maximum over the three levels and the following Add
// (see GemmMaxUp): out(B,H,W,Cout|ldo) = max(bf16(act(x . wt^T * scale + shift)), up(z1), up(z2)) (+ residual).
LDC_API int ldconv_conv1x1_bn_act_maxup_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                                            const void* z1, int H1, int W1, const void* z2, int H2, int W2, const void* residual,
                                            int ldr, void* out, int ldo, int B, int H, int W, int Cin, int Cout, int act, int dtype,
                                            void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_conv1x1_bn_act_maxup_fwd: bf16 only");
    LDC_REQUIRE(x && wt && out && z1 && z2 && B >= 0 && H >= 1 && W >= 1 && H1 >= 1 && W1 >= 1 && H2 >= 1 && W2 >= 1 &&
                (long long)B * H * W < (1ll << 31), "ldconv_conv1x1_bn_act_maxup_fwd: bad arguments");
    LDC_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && ldx >= Cin && ldo >= Cout && ldo % 8 == 0 && Cout <= 256 && Cout % 16 == 0,
                "ldconv_conv1x1_bn_act_maxup_fwd: needs Cin %% 8 == 0, ldx %% 8 == 0, ldo %% 8 == 0, Cout %% 16 == 0, Cout <= 256");
    LDC_REQUIRE(aligned16(x) && aligned16(wt) && aligned16(out) && aligned16(z1) && aligned16(z2) && (!residual || aligned16(residual)),
                "ldconv_conv1x1_bn_act_maxup_fwd: 16-byte alignment");
    if (B == 0) return LDCONV_OK;
    g_maxup = GemmMaxUp{(const __nv_bfloat16*)z1, (const __nv_bfloat16*)z2, H, W, H1, W1, H2, W2};
    return umma_gemm_fwd_ld(x, ldx, wt, scale, shift, out, nullptr, residual, ldr, ldo, nullptr, nullptr, B * H * W, Cin, Cout, act,
                            (cudaStream_t)stream);
}

// Last 1x1 conv of a Detect branch with the decode in its epilogue (see GemmDetect): x (B*H*W, Cin | ldx) bf16, wt (Cout, Cin);
// mode 1: Cout = 64 (4 sides x 16 DFL bins) -> y rows 0..3 (xywh * stride); mode 2: Cout = nc <= 16 -> y rows 4..4+nc (sigmoid).
// y (B, 4+nc, total) bf16; this level's anchors occupy columns [a0, a0 + H*W).  Bit-identical to ldconv_conv1x1_bn_act_fwd (no
// activation) followed by ldconv_detect_decode.
LDC_API int ldconv_conv1x1_detect_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift, void* y, int mode,
                                      int B, int H, int W, int Cin, int Cout, int nc, float stride, int a0, int total, int dtype,
                                      void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_conv1x1_detect_fwd: bf16 only");
    LDC_REQUIRE(x && wt && y && B >= 0 && H >= 1 && W >= 1 && nc >= 1 && nc <= 16, "ldconv_conv1x1_detect_fwd: bad arguments");
    LDC_REQUIRE((mode == 1 && Cout == 64) || (mode == 2 && Cout == nc), "ldconv_conv1x1_detect_fwd: mode 1 needs Cout = 64, mode 2 Cout = nc");
    LDC_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && ldx >= Cin && aligned16(x) && aligned16(wt), "ldconv_conv1x1_detect_fwd: Cin / ldx / alignment");
    LDC_REQUIRE(a0 >= 0 && (long long)a0 + (long long)H * W <= total && (long long)B * H * W < (1ll << 31), "ldconv_conv1x1_detect_fwd: anchor range");
    if (B == 0) return LDCONV_OK;
    g_detect = GemmDetect{(__nv_bfloat16*)y, mode, H, W, nc, a0, total, stride};
    return umma_gemm_fwd_ld(x, ldx, wt, scale, shift, nullptr, nullptr, nullptr, 0, Cout, nullptr, nullptr, B * H * W, Cin, Cout,
                            LDCONV_ACT_NONE, (cudaStream_t)stream);
}

// The 1x1 `Conv` block on P consecutive pixels per operand row (see GemmPack): x (rows, Cin) dense, wt_packed (P Cout, P Cin) =
// block_diag(wt, ..., wt), scale_rep / shift_rep = the folded BatchNorm repeated P times, out (rows, Cout | ldo).
LDC_API int ldconv_conv1x1_bn_act_packed_fwd(const void* x, const void* wt_packed, const float* scale_rep, const float* shift_rep,
                                             void* out, int ldo, long long rows, int Cin, int Cout, int P, int act, int dtype,
                                             void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_conv1x1_bn_act_packed_fwd: bf16 only");
    LDC_REQUIRE(x && wt_packed && out && rows >= 0 && rows < (1ll << 31), "ldconv_conv1x1_bn_act_packed_fwd: bad arguments");
    int o_shift = 0;
    while ((1 << o_shift) < Cout) ++o_shift;
    LDC_REQUIRE((P == 2 || P == 4) && Cout >= 16 && (1 << o_shift) == Cout && P * Cout <= 256 && Cin % 8 == 0 && rows % P == 0 &&
                ldo >= Cout && ldo % 8 == 0,
                "ldconv_conv1x1_bn_act_packed_fwd: needs P in {2, 4}, Cout a power of two >= 16, P Cout <= 256, Cin %% 8 == 0, "
                "rows %% P == 0, ldo %% 8 == 0 (got P=%d Cin=%d Cout=%d rows=%lld ldo=%d)", P, Cin, Cout, rows, ldo);
    LDC_REQUIRE(aligned16(x) && aligned16(wt_packed) && aligned16(out), "ldconv_conv1x1_bn_act_packed_fwd: 16-byte alignment");
    if (rows == 0) return LDCONV_OK;
    g_pack = GemmPack{P, o_shift};
    return umma_gemm_fwd_ld(x, P * Cin, wt_packed, scale_rep, shift_rep, out, nullptr, nullptr, 0, ldo, nullptr, nullptr,
                            (int)(rows / P), P * Cin, P * Cout, act, (cudaStream_t)stream);
}

// 1x1 `Conv` block (Conv2d(1x1, no bias) + folded BatchNorm + activation, nn/modules/conv.py:41-59) = the same GEMM with
// pixel strides: x and out may be channel slices of wider NHWC buffers.
LDC_API int ldconv_conv1x1_bn_act_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                                      const void* residual, int ldr, void* out, int ldo, long long rows, int Cin, int Cout,
                                      int act, int dtype, void* stream)
{
    LDC_REQUIRE(dtype == LDCONV_BF16, "ldconv_conv1x1_bn_act_fwd: bf16 only (the fp32 path keeps the framework conv)");
    LDC_REQUIRE(x && wt && out && rows >= 0 && rows < (1ll << 31), "ldconv_conv1x1_bn_act_fwd: bad arguments");
    LDC_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && ldx >= Cin && ldo >= Cout && Cout <= 256,
                "ldconv_conv1x1_bn_act_fwd: needs Cin %% 8 == 0, ldx %% 8 == 0, Cout <= 256 (got Cin=%d ldx=%d Cout=%d)", Cin, ldx,
                Cout);
    LDC_REQUIRE(aligned16(x) && aligned16(wt), "ldconv_conv1x1_bn_act_fwd: x / wt must be 16-byte aligned");
    if (rows == 0) return LDCONV_OK;
    return umma_gemm_fwd_ld(x, ldx, wt, scale, shift, out, nullptr, residual, ldr, ldo, nullptr, nullptr, (int)rows, Cin, Cout,
                            act, (cudaStream_t)stream);
}

