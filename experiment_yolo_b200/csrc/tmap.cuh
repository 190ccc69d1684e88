// tmap.cuh -- host-side CUtensorMap construction through the driver entry point (no link against libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace ldc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rank-`rank` tiled map; dims / box innermost first; strides in bytes for dims 1..rank-1
inline int encode_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* gdim,
                      const cuuint64_t* gstride, const cuuint32_t* box, CUtensorMapSwizzle swz)
{
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(LDCONV_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LDCONV_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LDCONV_OK;
}

}  // namespace ldc
