// Host side of the TMA-fed kernels: a CUtensorMap over fp16 operand planes [pp][2 (hi, lo)][chunks][H][W][8]
// (include/aarmvs.h, aarmvs_plane_src) whose box is the staged halo tile of the implicit GEMMs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace planemap {

constexpr int BOX_COLS = 10, BOX_ROWS = 18;          // halo tile of a 16 x 8 pixel tile

// cuTensorMapEncodeTiled through the runtime's driver entry point query (the library links no libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Operand planes [pp][2][chunks][H][W][8 halfs] -> 5-D map with the halo-tile box.
inline int make_plane_map(CUtensorMap* map, const void* planes, int H, int W, int chunks, int pp) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return AARMVS_ERR_ARG;
    const cuuint64_t row = (cuuint64_t)W * 16, plane = (cuuint64_t)H * row;
    cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)chunks, 2, (cuuint64_t)pp};
    cuuint64_t strides[4] = {row, plane, plane * chunks, plane * chunks * 2};
    cuuint32_t box[5] = {(cuuint32_t)BOX_COLS * 8, (cuuint32_t)BOX_ROWS, (cuuint32_t)chunks, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(planes), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? AARMVS_OK : AARMVS_ERR_ARG;
}


}  // namespace planemap
