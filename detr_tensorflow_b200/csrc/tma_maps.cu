// tma_maps.cu -- im2col tensor maps (cuTensorMapEncodeIm2col) for the implicit-GEMM convolutions of gemm_tc.cu / wgrad_tc.cu.
// (The conventions of this driver entry point -- bounding-box corners, traversal stride, filter offsets, zero fill -- are pinned by
// tests/test_tma_im2col_gpu.py through a probe kernel that lives with the tests: tests/native/tma_probe.cu.)
#include "common.cuh"
#include <cuda.h>

namespace {

typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

void *detrb_get_im2col_encode()
{
    static void *fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult q;
        void *ptr = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = ptr;
    }
    return fn;
}

// NHWC bf16 tensor [B,H,W,C] -> im2col tensor map (channelsPerPixel = channels, pixelsPerColumn = pixels;
// swizzle128: 0 none, 1 128-byte, 2 32-byte)
int detrb_make_im2col_map(void *map_out, const void *x, int B, int H, int W, int C, int ldc, int lower_w, int lower_h,
                          int upper_w, int upper_h, int stride, int pixels, int swizzle128, int channels)
{
    EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(detrb_get_im2col_encode());
    if (!fn) DETRB_FAIL(DETRB_E_CUDA, "cuTensorMapEncodeIm2col not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ldc * 2, (cuuint64_t)W * ldc * 2, (cuuint64_t)H * W * ldc * 2};
    int lower[2] = {lower_w, lower_h}, upper[2] = {upper_w, upper_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(map_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x), dims, strides,
                    lower, upper, (cuuint32_t)channels, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DETRB_FAIL(DETRB_E_CUDA, "cuTensorMapEncodeIm2col failed: %d", (int)r);
    return DETRB_OK;
}
