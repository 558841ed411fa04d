// abi.cu -- version / error plumbing of the C ABI (include/detrb.h).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void detrb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int detrb_version(void) { return 223; }   // 0.2.2.3: detrb_set_wgrad_tile;  0.2.2.2: handles (detrb_create / detrb_bind), thread-local policy switches;  0.2.2.1: detrb_attn_bwd_t.parts;  0.2.2: detrb_igemm_t.scratch;  0.2.1: 1-bit ReLU masks (mask_bits / out_bits), detrb_resize_affine_u8;  0.2.0: parity-precision planes, detrb_accumulate, set_loss status
// (bump on EVERY change of a struct or prototype in include/detrb.h: _lib.py refuses to load a library of another version)

extern "C" const char *detrb_last_error(void) { return g_err; }

extern "C" int detrb_check_device(void)
{
    int dev = 0;
    DETRB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DETRB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        DETRB_FAIL(DETRB_E_ARCH, "libdetrb is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)", dev, prop.major, prop.minor);
    return DETRB_OK;
}

static thread_local int g_pdl = 1;          // policy switches are per thread: see detrb_bind (handle.cu)
bool detrb_pdl_enabled() { return g_pdl != 0; }
extern "C" int detrb_set_pdl(int enable) { int old = g_pdl; g_pdl = enable; return old; }
