// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (gemm_tc.cu, wgrad_tc.cu): mbarrier, TMA, tcgen05.
#pragma once
#include "common.cuh"
#include <cuda.h>

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// the same wait for the single-thread producer / MMA warps of kernels whose other warps are issue-bound (attention): the hardware may
// suspend the thread for up to `hint_ns` inside try_wait, and a failed probe backs off with nanosleep -- the polling loop of
// mbar_wait (try_wait + branch) otherwise competes for issue slots with the softmax warp on the same scheduler (ncu: 38 % of all
// executed instructions of the attention forward kernel were polling)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
    while (true) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(4u * sleep_ns) : "memory");
        if (ok) break;
        __nanosleep(sleep_ns);
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// fetch a tensor map (kernel parameter) into the descriptor cache ahead of its first use
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}
// pull a 2-D tile into L2 only (no shared-memory destination, no barrier): hides DRAM latency of a later tma_load_2d
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" :: "l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
// One thread of the (converged) warp: elect.sync.  Branching on THIS predicate tells ptxas that exactly one thread runs the block,
// so single-thread instructions that take uniform-register operands (tcgen05.mma, tcgen05.commit, TMA) are emitted once; under
// `if (lane == 0)` it wraps every one of them in an ELECT / BRA.U.ANY loop over the active mask (~50 cycles per tcgen05.mma:
// more than the 32 cycles a 128x64x16 MMA occupies the tensor pipe).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}

// 2-D bf16 tensor map helper (gemm_tc.cu): [rows, cols] row stride ld, box [box_rows, box_cols], 128- / 64- / 32-byte swizzle
bool detrb_make_tiled_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                          uint32_t box_cols, int swizzle_bytes = 128);

// ------------------------------------------------------------------------------------------------ shared by the tcgen05 GEMM / convolution kernels
constexpr int TBM = 128;
constexpr int TBK = 64;                 // 64 bf16 = 128 B = one swizzle row

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row core groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                  // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                  // version
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major A and B,
// N>>3 at bits [17,23), M>>4 at bits [24,29)   (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// v[0..8) += 8 consecutive floats of the bias row staged in shared memory (all lanes read the same address: broadcast)
__device__ __forceinline__ void add_bias8(float (&v)[8], uint32_t saddr) {
    float4 b0, b1;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(saddr));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(saddr + 16u));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
}

// 1-bit ReLU masks (detrb_igemm_t.mask_bits / out_bits): the 8 bytes of one row x 64-column chunk travel as a uint2; byte c (0..7)
// holds columns 8c .. 8c+7
__device__ __forceinline__ uint2 ld_bits8(const uint8_t *p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void apply_bits8(float (&v)[8], const uint2 &bits, int c, float scale) {
    const uint32_t b = ((c < 4 ? bits.x : bits.y) >> (8 * (c & 3))) & 0xffu;
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = ((b >> i) & 1u) ? v[i] * scale : 0.f;
}
__device__ __forceinline__ void collect_bits8(const float (&v)[8], uint2 &bits, int c) {
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) b |= (v[i] > 0.f ? 1u : 0u) << i;
    if (c < 4) bits.x |= b << (8 * (c & 3)); else bits.y |= b << (8 * (c & 3));
}


__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y; t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y; t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}

// Straight-line epilogue of one 64-column chunk of one accumulator row (thread = row): acc_r = the row's 64 fp32 accumulators, rr =
// its 64 residual values (8 x 16 B as they lie in the swizzled tile), bias from shared memory (zeros when there is none).  No flag or
// data-dependent branches inside: every load of the chunk is in flight before the first use.  With branches between the 8-column
// groups each group is a serial load -> use chain at LOADED shared-memory latency, and the epilogue, not HBM, sets the pace of the
// memory-bound layers (measured on the layer1 1x1 conv + residual: 170 us branchy, 102 us straight-line, same memory pipeline).
// the dropout of a chunk: rowhash = dropout_rowhash(seed, site, row) (once per row), col0 = global column of the chunk
struct EpiDrop { uint32_t rowhash, thresh, col0; float scale; };
template <bool HAS_R, bool MBITS, bool OBITS, bool DROP = false>
__device__ __forceinline__ void epi_chunk_math(const uint32_t (&acc_r)[64], const uint4 (&rr)[8], uint32_t sbias64, bool relu, const uint2 &mbc,
                                               float mscale, uint2 &ob, uint32_t obuf_row, uint32_t sw, const EpiDrop &dr = EpiDrop())
{
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float v[8], res[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __uint_as_float(acc_r[c * 8 + i]);
        add_bias8(v, sbias64 + 32u * (uint32_t)c);
        if (HAS_R) {
            unpack8(rr[c], res);
            if (!DROP) {                                    // with dropout the residual is the un-dropped skip path: it joins last
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] += res[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = relu ? fmaxf(v[i], 0.f) : v[i];
        if (MBITS) apply_bits8(v, mbc, c, mscale);
        if (DROP) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                bool k0, k1;
                dropout_keep2(dropout_bits_rh(dr.rowhash, (dr.col0 + (uint32_t)(c * 8 + i)) >> 1), dr.thresh, k0, k1);
                v[i] = k0 ? v[i] * dr.scale : 0.f;
                v[i + 1] = k1 ? v[i + 1] * dr.scale : 0.f;
                if (HAS_R) { v[i] += res[i]; v[i + 1] += res[i + 1]; }
            }
        }
        if (OBITS) collect_bits8(v, ob, c);
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(obuf_row + (((uint32_t)c ^ sw) << 4)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
}
__device__ __forceinline__ void lds_row8(uint4 (&rr)[8], uint32_t row_addr, uint32_t sw) {
#pragma unroll
    for (int c = 0; c < 8; c++)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rr[c].x), "=r"(rr[c].y), "=r"(rr[c].z), "=r"(rr[c].w)
                     : "r"(row_addr + (((uint32_t)c ^ sw) << 4)));
}
__device__ __forceinline__ void tc_ld64(uint32_t taddr, uint32_t (&r)[64]) {
    tc_ld16(taddr, *reinterpret_cast<uint32_t (*)[16]>(&r[0]));
    tc_ld16(taddr + 16u, *reinterpret_cast<uint32_t (*)[16]>(&r[16]));
    tc_ld16(taddr + 32u, *reinterpret_cast<uint32_t (*)[16]>(&r[32]));
    tc_ld16(taddr + 48u, *reinterpret_cast<uint32_t (*)[16]>(&r[48]));
}

