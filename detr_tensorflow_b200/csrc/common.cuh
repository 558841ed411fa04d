// common.cuh -- shared device/host helpers for libdetrb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/detrb.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- host-side error plumbing
void detrb_set_error(const char *fmt, ...);
#define DETRB_FAIL(code, ...) do { detrb_set_error(__VA_ARGS__); return (code); } while (0)
#define DETRB_REQUIRE(cond, ...) do { if (!(cond)) DETRB_FAIL(DETRB_E_BADARG, __VA_ARGS__); } while (0)
#define DETRB_CHECK_LAUNCH(name) do { cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) DETRB_FAIL(DETRB_E_CUDA, "%s: %s", name, cudaGetErrorString(e_)); } while (0)
#define DETRB_CUDA(call) do { cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) DETRB_FAIL(DETRB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute opt-ins (dynamic shared memory above 48 KB) belong to a DEVICE, not to the process: one flag per device
// ordinal, so that a process driving several GPUs configures each of them.  slot() is the current device's flag (a scratch flag,
// i.e. "configure every time", outside 0..63 or when the runtime cannot name the device).
struct detrb_per_device_flag {
    bool done[64] = {};
    bool scratch = false;
    bool &slot()
    {
        int d = -1;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) { scratch = false; return scratch; }
        return done[d];
    }
};

// gemm_tc.cu (tcgen05 / TMA / TMEM path)
bool detrb_gemm_tc_supported(const detrb_igemm_t &p);
bool detrb_gemm_tc_enabled();
bool detrb_gemm_tc_conv_enabled();
int detrb_gemm_tc_kind(const detrb_igemm_t &p);
int detrb_gemm_tc(const detrb_igemm_t &p, cudaStream_t stream);
// elementwise.cu: stride-2 scatter of densely written parity-class / shortcut results (detrb_igemm_t.scratch)
int detrb_scatter_s2(const bf16 *const src[4], bf16 *dst, int ldc, const uint8_t *mask_bits, int ldmb, int B, int H, int W, int C,
                     int accumulate, cudaStream_t stream);
// conv_halo.cu (3x3 / 64-channel convolution, every input pixel staged once)
bool detrb_conv_halo_supported(const detrb_igemm_t &p);
int detrb_conv_halo(const detrb_igemm_t &p, cudaStream_t stream);
// wgrad_tc.cu
bool detrb_wgrad_tc_supported(const detrb_wgrad_t &p);
bool detrb_wgrad_tc_enabled();
bool detrb_wgrad_tc_profitable(const detrb_wgrad_t &p);
int detrb_wgrad_tc(const detrb_wgrad_t &p, cudaStream_t stream);
// attention_tc.cu (tcgen05 attention core)
bool detrb_attn_tc_enabled();
bool detrb_attn_fwd_tc_supported(const detrb_attn_fwd_t &p);
int detrb_attn_fwd_tc(const detrb_attn_fwd_t &p, cudaStream_t stream);
bool detrb_attn_bwd_tc_supported(const detrb_attn_bwd_t &p);
int detrb_attn_bwd_tc(const detrb_attn_bwd_t &p, cudaStream_t stream);
// tma_maps.cu: im2col tensor maps
void *detrb_get_im2col_encode();
int detrb_make_im2col_map(void *map_out, const void *x, int B, int H, int W, int C, int ldc, int lower_w, int lower_h,
                          int upper_w, int upper_h, int stride, int pixels, int swizzle128, int channels = 64);

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it may become resident and run its
// prologue (barrier init, TMEM allocation, index setup) while the previous kernel of the stream drains, and only blocks in
// pdl_wait() -- which returns once the predecessor has completed and its writes are visible -- before touching global memory.
// ~660 dependent launches per train step: the launch gaps this removes are worth more than most kernel-level tuning.
bool detrb_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = detrb_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DETRB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    do { cudaError_t le_ = launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__); \
         if (le_ != cudaSuccess) DETRB_FAIL(DETRB_E_CUDA, "launch %s: %s", #kernel, cudaGetErrorString(le_)); } while (0)

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte (or 8-byte) async global->shared copy; src_bytes==0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" :: "n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r0), "=r"(r1) : "r"(addr));
}

// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162 *>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- parity precision: bf16 pairs (include/detrb.h header comment)
// element x = hi + lo, hi = bf16(x) at p, lo = bf16(x - hi) at p + split; split == 0: plain bf16 (lo does not exist)
__device__ __forceinline__ void sp_unpack8(const uint4 &u, float (&f)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 sp_pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}
// 8 consecutive elements (16-byte aligned)
__device__ __forceinline__ void sp_ld8(const bf16 *p, long long split, float (&f)[8]) {
    sp_unpack8(*reinterpret_cast<const uint4 *>(p), f);
    if (split) {
        float l[8];
        sp_unpack8(*reinterpret_cast<const uint4 *>(p + split), l);
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] += l[i];
    }
}
__device__ __forceinline__ void sp_st8(bf16 *p, long long split, const float (&f)[8]) {
    const uint4 hi = sp_pack8(f);
    *reinterpret_cast<uint4 *>(p) = hi;
    if (split) {
        float h[8], l[8];
        sp_unpack8(hi, h);
#pragma unroll
        for (int i = 0; i < 8; i++) l[i] = f[i] - h[i];
        *reinterpret_cast<uint4 *>(p + split) = sp_pack8(l);
    }
}
// the value a later sp_ld8 of the same location returns (the pair's rounding applied to f)
__device__ __forceinline__ void sp_round8(long long split, float (&f)[8]) {
    float h[8];
    sp_unpack8(sp_pack8(f), h);
    if (split) {
        float l[8], lr[8];
#pragma unroll
        for (int i = 0; i < 8; i++) l[i] = f[i] - h[i];
        sp_unpack8(sp_pack8(l), lr);
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = h[i] + lr[i];
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = h[i];
    }
}
// 2 consecutive elements (4-byte aligned)
__device__ __forceinline__ float2 sp_ld2(const bf16 *p, long long split) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(p));
    if (split) { const float2 l = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(p + split)); v.x += l.x; v.y += l.y; }
    return v;
}
__device__ __forceinline__ void sp_st2(bf16 *p, long long split, float a, float b) {
    const uint32_t hi = pack_bf16x2(a, b);
    *reinterpret_cast<uint32_t *>(p) = hi;
    if (split) { const float2 h = unpack_bf16x2(hi); *reinterpret_cast<uint32_t *>(p + split) = pack_bf16x2(a - h.x, b - h.y); }
}
__device__ __forceinline__ float sp_ld1(const bf16 *p, long long split) {
    float v = __bfloat162float(*p);
    if (split) v += __bfloat162float(p[split]);
    return v;
}
__device__ __forceinline__ void sp_st1(bf16 *p, long long split, float a) {
    const bf16 h = __float2bfloat16(a);
    *p = h;
    if (split) p[split] = __float2bfloat16(a - __bfloat162float(h));
}

// ---------------------------------------------------------------- counter-based dropout RNG
// 32 random bits for the counter (site, row, pair) under `seed`, from the "lowbias32" integer finaliser: one full round for the
// row (hoistable out of inner loops: dropout_rowhash) and one for the column pair.  One call serves two adjacent elements
// (16 bits each): element (row, col) uses pair = col >> 1 and the low / high half for even / odd col.
// keep <=> bits16 >= thresh16, thresh16 = round(p * 65536).  Forward and backward regenerate identical masks.
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t dropout_rowhash(uint64_t seed, uint32_t site, uint32_t row) {
    return lowbias32(row ^ (uint32_t)seed ^ (site * 0x9E3779B9U)) ^ (uint32_t)(seed >> 32);
}
// per-pair mixing: the row hash is already a full lowbias32 round, so the pair needs only the first half of one
// (xorshift, multiply, xorshift: 7 integer ops per two elements -- this sits in the inner loop of the attention kernels)
__device__ __forceinline__ uint32_t dropout_bits_rh(uint32_t rowhash, uint32_t pair) {
    uint32_t x = rowhash ^ (pair * 0x85EBCA77U);
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15;
    return x;
}
__device__ __forceinline__ uint32_t dropout_bits(uint64_t seed, uint32_t site, uint32_t row, uint32_t pair) {
    return dropout_bits_rh(dropout_rowhash(seed, site, row), pair);
}
__host__ __device__ __forceinline__ uint32_t dropout_thresh16(float p) {
    return (uint32_t)(p * 65536.0f + 0.5f);
}
// keep flags for the element pair (col even, col odd) of `pair`
__device__ __forceinline__ void dropout_keep2(uint32_t bits, uint32_t thresh, bool &k0, bool &k1) {
    k0 = (bits & 0xffffu) >= thresh;
    k1 = (bits >> 16) >= thresh;
}

// ---------------------------------------------------------------- attention-probability dropout (attention.cu)
// The attention kernels regenerate their masks three times per layer (forward, dK/dV, dQ) inside loops that are bound by
// instruction issue, so this RNG is built for instruction count: one 32-bit word carries two 15-bit uniform fields (bits
// [0,15) and [16,31)) for the KEY PAIR (k, k+8) of a 16-key group -- the two keys one thread owns in the transposed (dK/dV)
// MMA layout, and the two column groups (j, j+1) it owns in the forward / dQ layout -- so every orientation needs half a word
// per score.  word = ((rowhash' + pair * C') ^ (.. >> 15)) * M2: a Weyl step, one xorshift, one multiply (4 instructions;
// serial / cross-row correlations and field uniformity checked in tests/test_box_math_cpu.py::test_attention_dropout_rng).
// keep <=> field >= thresh15 = round(p * 32768).  Both compares are ONE subtraction on the guarded fields, and the keep bits
// (15 and 31) become all-ones / all-zeros AND-masks with one PRMT (sign replication), applied to packed bf16x2 probabilities
// or to fp32 values: no predicates, no selects.
constexpr uint32_t ATTN_C1 = 0x9E3779B1u, ATTN_M1 = 0x21F0AAADu, ATTN_M2 = 0x735A2D97u;
__device__ __forceinline__ uint32_t attn_drop_rowhash(uint64_t seed, uint32_t site, uint32_t row) {
    return dropout_rowhash(seed, site, row) * ATTN_M1;
}
__host__ __device__ __forceinline__ uint32_t attn_drop_pair(uint32_t k) { return ((k >> 4) << 3) | (k & 7u); }   // half = (k >> 3) & 1
__device__ __forceinline__ uint32_t attn_drop_word(uint32_t rowhash_m, uint32_t pair) {
    uint32_t x = rowhash_m + pair * (ATTN_C1 * ATTN_M1);
    x ^= x >> 15;
    return x * ATTN_M2;
}
__host__ __device__ __forceinline__ uint32_t attn_drop_thresh2(float p) {       // threshold replicated into both fields
    const uint32_t t = (uint32_t)(p * 32768.0f + 0.5f);
    return t | (t << 16);
}
// bit 15 / bit 31 of the result: low / high field kept (0x8000 + field - t never borrows across the field boundary)
__device__ __forceinline__ uint32_t attn_drop_keepbits(uint32_t word, uint32_t thresh2) {
    return ((word & 0x7FFF7FFFu) | 0x80008000u) - thresh2;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// 32-bit all-ones / all-zeros mask from the low (bit 15) / high (bit 31) keep bit: selector nibble 8|n = sign of byte n
__device__ __forceinline__ uint32_t attn_drop_mask_lo(uint32_t kb) { return prmt(kb, kb, 0x9999u); }
__device__ __forceinline__ uint32_t attn_drop_mask_hi(uint32_t kb) { return prmt(kb, kb, 0xBBBBu); }
// bf16x2 masks for the element pair (a, b) whose keep bits live in words ka (low half of the result) and kb (high half)
__device__ __forceinline__ uint32_t attn_drop_mask2_lo(uint32_t ka, uint32_t kb) { return prmt(ka, kb, 0xDD99u); }
__device__ __forceinline__ uint32_t attn_drop_mask2_hi(uint32_t ka, uint32_t kb) { return prmt(ka, kb, 0xFFBBu); }
