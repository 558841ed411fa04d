// tma_probe.cu -- TEST INFRASTRUCTURE (built into tests/_harness_build/libdetrb_probe.so by __graft_entry__.build(), linked against
// libdetrb.so; not part of the product library): one TMA im2col load (cp.async.bulk.tensor.4d ... .im2col) of an NHWC bf16 tensor
// into shared memory, dumped raw to global memory.  tests/test_tma_im2col_gpu.py uses it to pin down the coordinate / bounding-box
// conventions of cuTensorMapEncodeIm2col that the convolution kernels rely on.
#include "../../detr_tensorflow_b200/csrc/common.cuh"
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

}  // namespace

/* one TMA im2col load (channelsPerPixel = 64, pixelsPerColumn = pixels) of NHWC bf16 x[B,H,W,C] dumped raw into
 * out[pixels*128 + 1] (last byte: 1 if the load completed) */
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
