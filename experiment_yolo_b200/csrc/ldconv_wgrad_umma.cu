// ldconv_wgrad_umma.cu -- weight gradient of the LDConv (N,1) conv as a tcgen05 tensor-core reduction (bf16, sm_100a).
//
//   dWt(O, K) += grad_pre(M, O)^T . operand(M, K)        (autograd of nn.Conv2d(inc, outc, (N,1), (N,1)),
//                                                          /root/reference/ultralytics/nn/modules/conv.py:355,408)
// The reduction runs over the long M axis (M = B*h*w, up to 6.5 M rows; O x K is at most 128 x 192 in the model).  Both
// operands are fed to the MMA TRANSPOSED straight from their row-major HBM layout: a TMA box of (64 columns x R rows) of a
// row-major matrix is exactly the MN-major SWIZZLE_128B shared-memory layout of tcgen05 (the M/N index contiguous, the MMA's
// K index = our row index m striding 128 B), so
//     D[o, k] (TMEM, fp32)  +=  sum over 16 rows m :  G[m, o] * A[m, k]
// is one tcgen05.mma with a_major = b_major = MN, M_ = 128 (O zero-padded by the TMA out-of-bounds fill), N_ = K tile <= 256.
//   warp 0   TMA producer: R-row chunks of grad_pre (two 64-column boxes) and of the operand (K-tile/64 boxes), 4-stage ring
//   warp 1   one thread issues R/16 MMAs per chunk, accumulating in TMEM over all chunks of this CTA
//   warps 2-5  after the last chunk: tcgen05.ld the O x K-tile accumulator and atomically add it to dWt (fp32)
// grid = (M splits, K tiles): every CTA reduces its share of the rows; fp32 atomics combine the CTAs (order is not
// deterministic, parity is tolerance-based like every gradient of the path).
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace ldc {

using namespace umma;

static constexpr int kWgThreads = 192;
static constexpr int kWgRows = 64;          // rows of M per stage
static constexpr int kWgStages = 4;

__device__ __forceinline__ uint64_t wg_desc_mn(uint32_t addr, uint32_t lbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;     // stride between 64-element blocks along M_/N_
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride between 8-row groups along the reduction index
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmA, float* __restrict__ dW,
                  int M, int K, int O, int nb, int chunks_total)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS, not generic LD / ST
    const uint32_t g_bytes = 2u * kWgRows * 128u;                 // two 64-column boxes of grad_pre
    const uint32_t a_bytes = (uint32_t)nb * kWgRows * 128u;       // nb 64-column boxes of the operand
    const uint32_t stage_bytes = g_bytes + a_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kWgStages * stage_bytes);
    uint64_t* empty = full + kWgStages;
    uint64_t* done = empty + kWgStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N_ = nb * 64;
    const int k_tile0 = blockIdx.y * 256;
    const uint32_t tmem_cols = N_ <= 32 ? 32u : (N_ <= 64 ? 64u : (N_ <= 128 ? 128u : 256u));

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmA);
        for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // chunks of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_chunks = chunks_total > (int)blockIdx.x ? (chunks_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    // control warps: warp-uniform loops, one elected lane issues (no ELECT / BRA.U.ANY loop around every UTMALDG / UTCHMMA)
    if (warp == 0) {
        const bool leader = elect_one();
        {
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < my_chunks; ++c) {
                const int row0 = (blockIdx.x + c * gridDim.x) * kWgRows;
                mbar_wait(&empty[s], ph ^ 1);
                if (leader) {
                    mbar_arrive_expect_tx(&full[s], stage_bytes);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    tma_load_2d(st, &tmG, &full[s], 0, row0);
                    tma_load_2d(st + kWgRows * 128, &tmG, &full[s], 64, row0);
                    for (int j = 0; j < nb; ++j) tma_load_2d(st + g_bytes + (size_t)j * kWgRows * 128, &tmA, &full[s], k_tile0 + j * 64, row0);
                }
                __syncwarp();
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        {
            // D format F32, A/B BF16, both operands MN-major (bits 15, 16), M_ = 128, N_ = nb*64
            const uint32_t idesc = make_idesc_bf16(128, N_) | (1u << 15) | (1u << 16);
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < my_chunks; ++c) {
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                const uint32_t g_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t a_addr = g_addr + g_bytes;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < kWgRows / 16; ++k)
                        mma_bf16_ss(tmem_base, wg_desc_mn(g_addr + k * 2048, kWgRows * 128), wg_desc_mn(a_addr + k * 2048, kWgRows * 128),
                                    idesc, (uint32_t)((c | k) != 0));
                    mma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
            if (leader) mma_commit(done);
            __syncwarp();
        }
    } else if (my_chunks > 0) {
        mbar_wait(done, 0);
        tc_fence_after_sync();
        const int lg = warp & 3;
        const int o = lg * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
        for (int c0 = 0; c0 < N_; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (o < O) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int k = k_tile0 + c0 + e;
                    if (k < K) atomicAdd(dW + (size_t)o * K + k, __uint_as_float(v[e]));
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int wgrad_umma_supported(int M, int K, int O, const void* g, const void* a)
{
    return M >= 1 && O >= 1 && O <= 128 && K % 8 == 0 && O % 8 == 0 && aligned16(g) && aligned16(a);
}

int wgrad_umma(const void* g, const void* a, float* dW, int M, int K, int O, cudaStream_t st)
{
    CUtensorMap tmG, tmA;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)O, (cuuint64_t)M};
        cuuint64_t gstr[1] = {(cuuint64_t)O * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)kWgRows};
        if (int e = encode_map(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
        cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)kWgRows};
        if (int e = encode_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
    const int k_tiles = (K + 255) / 256;
    const int k_last = K - (k_tiles - 1) * 256;
    // every K tile uses the same box count (the widest); boxes past K are zero-filled by the TMA unit
    const int nb = k_tiles > 1 ? 4 : (k_last + 63) / 64;
    const int chunks = (M + kWgRows - 1) / kWgRows;
    int gx = num_sms() / k_tiles;
    if (gx < 1) gx = 1;
    if (gx > chunks) gx = chunks;
    const size_t smem = 1024 + (size_t)kWgStages * (2u * kWgRows * 128u + (size_t)nb * kWgRows * 128u) + (2 * kWgStages + 1) * 8 + 16;
    LDC_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)gx, (unsigned)k_tiles);
    wgrad_umma_kernel<<<grid, kWgThreads, smem, st>>>(tmG, tmA, dW, M, K, O, nb, chunks);
    LDC_LAUNCH_CHECK("wgrad_umma_kernel");
    return LDCONV_OK;
}

}  // namespace ldc
