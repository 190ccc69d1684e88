// ubench.cu -- latency micro-benchmarks that size the warp-specialised pipelines (sm_100a):
//   tcgen05.mma issue -> tcgen05.commit -> mbarrier completion, for 1..64 MMAs of M=128, N in {16,64,256}, K=16
//   mbarrier arrive -> try_wait wake-up between two warps
//   TMA box load latency (2-D 128B-swizzled 16 KB box; 4-D box with 64-byte rows)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu    run: ./ubench
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../experiment_yolo_b200/csrc/umma.cuh"

using namespace ldc::umma;

__global__ void __launch_bounds__(128, 1) mma_latency(long long* out, int n_mma, int N)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
        uint32_t parity = 0;
        for (int rep = 0; rep < 4; ++rep) {
            long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i)
                mma_bf16_ss(tm, make_desc_k_sw128(a + (i & 3) * 32), make_desc_k_sw128(b + (i & 3) * 32), idesc, i != 0);
            long long t1 = clock64();
            mma_commit(&bar);
            mbar_wait(&bar, parity);
            long long t2 = clock64();
            parity ^= 1;
            out[rep * 2 + 0] = t1 - t0;
            out[rep * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// same measurement with the whole warp running the loop convergently and one ELECTED lane issuing (CUTLASS style)
__global__ void __launch_bounds__(128, 1) mma_latency_elect(long long* out, int n_mma, int N)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
        const uint64_t da = make_desc_k_sw128(a), db = make_desc_k_sw128(b);
        uint32_t parity = 0;
        for (int rep = 0; rep < 4; ++rep) {
            long long t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < n_mma; ++i)
                    mma_bf16_ss(tm, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, i != 0);
            }
            __syncwarp();
            long long t1 = clock64();
            if (elect_one()) mma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, parity);
            long long t2 = clock64();
            parity ^= 1;
            if (threadIdx.x == 0) { out[rep * 2 + 0] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

__global__ void __launch_bounds__(64, 1) handoff_latency(long long* out, int iters)
{
    __shared__ uint64_t ping, pong;
    if (threadIdx.x == 0) { mbar_init(&ping, 1); mbar_init(&pong, 1); fence_barrier_init(); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        if (warp == 0) {
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) { mbar_arrive(&ping); mbar_wait(&pong, i & 1); }
            out[0] = (clock64() - t0) / iters;   // one round trip = two hand-offs
        } else {
            for (int i = 0; i < iters; ++i) { mbar_wait(&ping, i & 1); mbar_arrive(&pong); }
        }
    }
}

__global__ void __launch_bounds__(32, 1) tma_latency(const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm4,
                                                     long long* out, int reps)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1); fence_barrier_init();
        uint32_t parity = 0;
        for (int r = 0; r < reps; ++r) {      // 2-D: 128 rows x 128 B, a fresh (cold) row block each time
            long long t0 = clock64();
            mbar_arrive_expect_tx(&bar, 16384);
            tma_load_2d(smem, &tm2, &bar, 0, (blockIdx.x * reps + r) * 128);
            mbar_wait(&bar, parity); parity ^= 1;
            out[r] = clock64() - t0;
        }
        for (int r = 0; r < reps; ++r) {      // 4-D: C=32 (64-byte rows) x 18 x 10 box
            long long t0 = clock64();
            mbar_arrive_expect_tx(&bar, 32 * 18 * 10 * 2);
            tma_load_4d(smem, &tm4, &bar, 0, 16 * r, 8 * (int)blockIdx.x, 0);
            mbar_wait(&bar, parity); parity ^= 1;
            out[reps + r] = clock64() - t0;
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    long long* d; cudaMalloc(&d, 4096); long long h[64];
    cudaFuncSetAttribute(mma_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int Ns[3] = {16, 64, 256}, cnt[5] = {1, 4, 16, 36, 64};
    for (int ni = 0; ni < 3; ++ni)
        for (int ci = 0; ci < 5; ++ci) {
            mma_latency<<<1, 128, 60 * 1024>>>(d, cnt[ci], Ns[ni]);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma_latency failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            printf("mma N=%3d count=%2d: issue %lld cyc, issue+commit+wait %lld cyc (rep3: %lld / %lld)\n", Ns[ni], cnt[ci], h[2], h[3], h[6], h[7]);
        }
    cudaFuncSetAttribute(mma_latency_elect, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int ni = 0; ni < 3; ++ni)
        for (int ci = 0; ci < 5; ++ci) {
            mma_latency_elect<<<1, 128, 60 * 1024>>>(d, cnt[ci], Ns[ni]);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma_latency_elect failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            printf("ELECT mma N=%3d count=%2d: issue %lld cyc, issue+commit+wait %lld cyc\n", Ns[ni], cnt[ci], h[6], h[7]);
        }
    handoff_latency<<<1, 64>>>(d, 1000);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("mbarrier ping-pong round trip (2 hand-offs): %lld cyc\n", h[0]);

    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const size_t rows = 1 << 20;
    void* buf; cudaMalloc(&buf, rows * 128); cudaMemset(buf, 0, rows * 128);
    CUtensorMap tm2, tm4;
    { cuuint64_t gd[2] = {64, rows}; cuuint64_t gs[1] = {128}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
      enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    { cuuint64_t gd[4] = {32, 1024, 1024, 2}; cuuint64_t gs[3] = {64, 64 * 1024, 64ull * 1024 * 1024}; cuuint32_t box[4] = {32, 18, 10, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
      enc(&tm4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    cudaFuncSetAttribute(tma_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    tma_latency<<<1, 32, 24 * 1024>>>(tm2, tm4, d, 8);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("tma_latency failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    printf("TMA 2-D 16 KB box (cold), cycles:"); for (int i = 0; i < 8; ++i) printf(" %lld", h[i]); printf("\n");
    printf("TMA 4-D 32ch x 18 x 10 box (11.5 KB, 64-byte rows), cycles:"); for (int i = 0; i < 8; ++i) printf(" %lld", h[8 + i]); printf("\n");
    return 0;
}
