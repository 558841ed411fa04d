// handle.cu -- detrb_handle_t: the per-GPU / per-rank state of the C ABI (include/detrb.h, "Handles").
//
// The reference keeps its per-model state in the Keras model object (networks/detr.py:116-204).  Below the Python mirror the only
// state the library itself has is the device a step runs on and the kernel-policy switches; a handle owns a copy of both and
// detrb_bind() installs them for the calling thread (the switches are thread-local, see gemm_tc.cu / conv_halo.cu / wgrad_tc.cu /
// attention_tc.cu / abi.cu).  No device memory is owned here: buffers, workspaces and weights stay with the caller.
#include "common.cuh"
#include <new>

struct detrb_handle {
    uint32_t magic;
    int device;
    int opt[DETRB_OPT_COUNT];
};

namespace {
constexpr uint32_t HANDLE_MAGIC = 0x44545242u;   // "DTRB"
constexpr int OPT_DEFAULT[DETRB_OPT_COUNT] = {1, 1, 1, 1, 1, 1, 1, -1, 1, 1};
typedef int (*SetFn)(int);
// the thread-local setter behind each option, in DETRB_OPT_* order
const SetFn OPT_SET[DETRB_OPT_COUNT] = {detrb_set_pdl, detrb_set_tc, detrb_set_tc_conv, detrb_set_tc_tma_epilogue, detrb_set_tc_persistent,
                                        detrb_set_tc_stream, detrb_set_tc_halo, detrb_set_tc_pair, detrb_set_tc_wgrad, detrb_set_tc_attn};
bool live(const detrb_handle_t *h) { return h && h->magic == HANDLE_MAGIC; }
}  // namespace

extern "C" int detrb_create(int device, detrb_handle_t **out)
{
    DETRB_REQUIRE(out, "detrb_create: null output pointer");
    *out = nullptr;
    int count = 0;
    DETRB_CUDA(cudaGetDeviceCount(&count));
    DETRB_REQUIRE(device >= 0 && device < count, "detrb_create: device %d out of range (%d visible)", device, count);
    int major = 0, minor = 0;
    DETRB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    DETRB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10)
        DETRB_FAIL(DETRB_E_ARCH, "libdetrb is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)", device, major, minor);
    detrb_handle_t *h = new (std::nothrow) detrb_handle_t;
    DETRB_REQUIRE(h, "detrb_create: out of host memory");
    h->magic = HANDLE_MAGIC;
    h->device = device;
    // the creating thread's current policy: the defaults (OPT_DEFAULT, or what the DETRB_* environment variables say) unless the
    // thread changed them -- read through the setters, which return the previous value
    for (int i = 0; i < DETRB_OPT_COUNT; i++) {
        h->opt[i] = OPT_SET[i](OPT_DEFAULT[i]);
        (void)OPT_SET[i](h->opt[i]);
    }
    *out = h;
    return DETRB_OK;
}

extern "C" int detrb_destroy(detrb_handle_t *h)
{
    if (!h) return DETRB_OK;
    DETRB_REQUIRE(live(h), "detrb_destroy: not a live handle");
    h->magic = 0;
    delete h;
    return DETRB_OK;
}

extern "C" int detrb_handle_device(const detrb_handle_t *h)
{
    DETRB_REQUIRE(live(h), "detrb_handle_device: not a live handle");
    return h->device;
}

extern "C" int detrb_handle_set(detrb_handle_t *h, int option, int value)
{
    DETRB_REQUIRE(live(h), "detrb_handle_set: not a live handle");
    DETRB_REQUIRE(option >= 0 && option < DETRB_OPT_COUNT, "detrb_handle_set: unknown option %d", option);
    h->opt[option] = value;
    return DETRB_OK;
}

extern "C" int detrb_handle_get(const detrb_handle_t *h, int option, int *value)
{
    DETRB_REQUIRE(live(h) && value, "detrb_handle_get: not a live handle / null output");
    DETRB_REQUIRE(option >= 0 && option < DETRB_OPT_COUNT, "detrb_handle_get: unknown option %d", option);
    *value = h->opt[option];
    return DETRB_OK;
}

extern "C" int detrb_bind(const detrb_handle_t *h)
{
    DETRB_REQUIRE(live(h), "detrb_bind: not a live handle");
    DETRB_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < DETRB_OPT_COUNT; i++) (void)OPT_SET[i](h->opt[i]);
    return DETRB_OK;
}
