// tma_probe.cu -- test helper: one TMA im2col load (cp.async.bulk.tensor.4d ... .im2col) of an NHWC bf16 tensor
// into shared memory, dumped raw to global memory.  Used by tests/test_tma_im2col_gpu.py to pin down the coordinate /
// bounding-box conventions of cuTensorMapEncodeIm2col that conv_tc.cu relies on.  Not on the product path.
#include "common.cuh"
#include <cuda.h>

namespace {

__global__ void tma_im2col_probe_kernel(const __grid_constant__ CUtensorMap map, int c0, int w, int h, int n,
                                        int off_w, int off_h, int bytes, uint8_t *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = smem_u32(&bar);
    const uint32_t dst = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < bytes + 1024; i += blockDim.x) smem[i] = 0xEE;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(bytes) : "memory");
        uint16_t ow = (uint16_t)off_w, oh = (uint16_t)off_h;
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
                     " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                     :: "r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh) : "memory");
    }
    // bounded wait so that a wrong transaction-byte count cannot hang the GPU
    bool done = false;
    for (int it = 0; it < 2000000 && !done; it++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar_a) : "memory");
        done = ok != 0;
    }
    __syncthreads();
    const unsigned char *src = smem + (dst - smem_u32(smem));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = src[i];
    if (threadIdx.x == 0) out[bytes] = done ? 1 : 0;
}

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

extern "C" int detrb_tma_im2col_probe(const detrb_bf16 *x, int B, int H, int W, int C, int lower_w, int lower_h, int upper_w,
                                      int upper_h, int stride, int pixels, int swizzle128, int c0, int w, int h, int n,
                                      int off_w, int off_h, uint8_t *out, detrb_stream_t stream)
{
    CUtensorMap map;
    int rc = detrb_make_im2col_map(&map, x, B, H, W, C, C, lower_w, lower_h, upper_w, upper_h, stride, pixels, swizzle128, 64);
    if (rc) return rc;
    const int bytes = pixels * 64 * 2;
    static bool configured = false;
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute(tma_im2col_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured = true;
    }
    DETRB_REQUIRE(bytes + 2048 <= 64 * 1024, "probe: too many pixels");
    tma_im2col_probe_kernel<<<1, 128, bytes + 2048, (cudaStream_t)stream>>>(map, c0, w, h, n, off_w, off_h, bytes, out);
    DETRB_CHECK_LAUNCH("tma_im2col_probe_kernel");
    return DETRB_OK;
}
