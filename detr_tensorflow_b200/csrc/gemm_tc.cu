// gemm_tc.cu -- Blackwell-native GEMM:  C[M,N] = epilogue( A[M,K] * W[N,K]^T ), bf16 operands, fp32 accumulation
// in TENSOR MEMORY.  TMA (cp.async.bulk.tensor, 128B swizzle) stages the operand tiles into shared memory, a single
// elected thread issues tcgen05.mma (cta_group::1, M=128 x N=BN x K=16 per instruction), completion is tracked with
// mbarriers (tcgen05.commit), the epilogue warps read the accumulator with tcgen05.ld and apply the same fused epilogue as
// igemm.cu (bias / residual / ReLU / ReLU-mask / sigmoid / dropout / bf16+fp32 stores / strided scatter-accumulate).
//
// Used by detrb_igemm for every GEMM-shaped launch of the train step: plain GEMMs (1x1 stride-1 convolutions and their data
// gradients, input_proj, all Linear layers of the transformer and the heads), implicit-GEMM convolutions through TMA im2col
// tensor maps (3x3, strided, transposed / data-gradient; stride-2 data gradients as four parity-class sub-convolutions), and
// the 7x7 stem as a sliding-window GEMM over the zero-padded space-to-depth image (detrb_igemm_t.a_kb_rows: the A tensor map
// has overlapping 128-byte rows).  Only unaligned shapes (the N = 92 / 4 heads) fall through to igemm.cu (mma.sync).
//
// Two kernels.  gemm_tc_kernel: one output tile per CTA; warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer, warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  BN x STAGES are sized so that 3-5 CTAs are
// co-resident per SM (the BN = 64 variants are capped at 64 registers): one CTA's epilogue overlaps another's loads -- the
// k-loops of the HBM-bound layers are 1..4 blocks long.  gemm_tcp_kernel (further down): persistent, one CTA per SM, 128x256
// tiles, for the tensor-bound shapes.  dispatch_tc / dispatch_tcp hold the measured policy.
#include "tc_common.cuh"
#include <type_traits>

#ifdef DETRB_TRACE
// developer build (-DDETRB_TRACE): per-phase cycle sums of the one-tile kernel's CTA life, read back by detrb_trace_read
__device__ unsigned long long g_trace[16];
#define TRACE_ADD(i, v) atomicAdd(&g_trace[i], (unsigned long long)(v))
#define TRACE_NOW() clock64()
#else
#define TRACE_ADD(i, v) ((void)(v))
#define TRACE_NOW() 0ll
#endif

namespace {

constexpr int NTHREADS_TC = 192;

// convolution side-band for the IM2COL kernels: effective padding and, per im2col tap, the weight tap to pair it with
// (identity: forward; reversed: stride-1 data gradient; sparse: one parity class of a stride-2 data gradient)
struct ConvAux { int pad; int wtap[16]; };

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = TBM * TBK * 2;
    static constexpr int B_BYTES = BN * TBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCH = BN / 64;                     // 64-column chunks of the output tile
    static constexpr int BAR_OFF = 0;                       // barriers (256 B), then BN floats of bias (BN shift)
    static constexpr int BIAS_OFF = 256;
    static constexpr int STG_OFF = BN > 128 ? 2048 : 1024;  // pipeline stages (1024-byte aligned: SWIZZLE_128B)
    static constexpr int EARLY_OFF = STG_OFF + STAGES * STAGE_BYTES;     // optional: residual chunks fetched at kernel start
    static constexpr int BASE = EARLY_OFF + 1024;           // + slack for the 1024-byte alignment of the dynamic smem base
    static constexpr int TOTAL = BASE + NCH * 16384;        // with the early-residual region
    // the freed stages hold the late residual + mask tiles (one-stage kernels: one of them, the residual goes early)
    static_assert(STAGES * STAGE_BYTES >= (STAGES == 1 ? 1 : 2) * NCH * 16384, "epilogue tiles do not fit in the pipeline stages");
};

// IM2COL: the A operand is gathered by a TMA im2col tensor map over the NHWC activation (implicit-GEMM convolution:
// 3x3 / strided / transposed); `aux` carries the effective padding and the im2col-tap -> weight-tap map.
// SPLIT (parity precision, include/detrb.h): A and W are bf16 pairs (hi plane / lo plane, tensor maps map_a2 / map_b2 for the lo
// planes); the k-loop runs three passes -- A_hi*W_hi, A_lo*W_hi, A_hi*W_lo -- through the same stages into the same TMEM
// accumulator; the epilogue (direct stores) reads residuals as hi + lo and writes the result as a pair.
// BITS: the 1-bit ReLU-mask code (mask_bits / out_bits) is compiled only into its own instantiations (and the SPLIT ones): inside the
// common kernels its flag tests cost the 64-register variants spills and every small GEMM of the transformer ~8 us.
template <int BN, int STAGES, bool IM2COL, bool SPLIT = false, bool BITS_T = false>
__global__ void __launch_bounds__(NTHREADS_TC, (BN == 64 && !SPLIT && !BITS_T) ? 5 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
               const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_b2, const detrb_igemm_t p, const ConvAux aux, const int tma_epi)
{
    using L = SmemLayout<BN, STAGES>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;            // SWIZZLE_128B needs 1024-byte alignment
    const uint32_t bar_base = smem0 + L::BAR_OFF;
    const uint32_t smem_base = smem0 + L::STG_OFF;                           // pipeline stages
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
    const uint32_t epi_bar = bar_base + 8u * (2 * STAGES + 2);               // early residual chunks
    const uint32_t late_bar = bar_base + 8u * (2 * STAGES + 3);              // mask (and late residual) chunks
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const bool r_early = (tma_epi & 2) != 0;                                 // residual tiles have their own smem: fetched at kernel start

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
    const int nk1 = p.K / TBK;
    const int nk = SPLIT ? 3 * nk1 : nk1;
    const long long split = SPLIT ? (long long)p.split : 0ll;
    const long long t_entry = TRACE_NOW();
    (void)t_entry;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (SPLIT) { tma_prefetch_desc(&map_a2); tma_prefetch_desc(&map_b2); }
        if (tma_epi) {
            tma_prefetch_desc(&map_c);
            if (p.residual) tma_prefetch_desc(&map_r);
            if (p.mask) tma_prefetch_desc(&map_m);
        }
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        mbar_init(epi_bar, 1);
        mbar_init(late_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const long long t_sync = TRACE_NOW();
    pdl_wait();          // everything above (barriers, TMEM) overlapped the previous kernel's tail
    const long long t_pdl = TRACE_NOW();
    if (threadIdx.x == 0) { TRACE_ADD(0, 1); TRACE_ADD(1, t_sync - t_entry); TRACE_ADD(2, t_pdl - t_sync); }

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int w0 = 0, h0 = 0, img = 0;
            if (IM2COL) {                                   // first output pixel of this tile -> window origin in the input
                const int ohw = p.OH * p.OW;
                img = m0 / ohw;
                const int rem = m0 - img * ohw, oy = rem / p.OW, ox = rem - oy * p.OW;
                const int st = p.mode == 0 ? p.stride : 1;
                w0 = ox * st - aux.pad; h0 = oy * st - aux.pad;
            }
            for (int kbt = 0; kbt < nk; kbt++) {
                const int part = SPLIT ? kbt / nk1 : 0, kb = kbt - part * nk1;
                const CUtensorMap *pa = (SPLIT && part == 1) ? &map_a2 : &map_a;
                const CUtensorMap *pb = (SPLIT && part == 2) ? &map_b2 : &map_b;
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
                const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
                if (IM2COL) {
                    const int k0 = kb * TBK, tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    tma_load_im2col(a_dst, pa, full_bar(stage), c0, w0, h0, img, (uint16_t)kw, (uint16_t)kh);
                    tma_load_2d(a_dst + L::A_BYTES, pb, full_bar(stage), aux.wtap[tap] * p.Cin + c0, n0);
                } else {
                    // sliding-window A (a_kb_rows > 0): k-block kb is the 64-element run that starts kb * a_kb_rows rows further down
                    tma_load_2d(a_dst, pa, full_bar(stage), p.a_kb_rows ? 0 : kb * TBK, m0 + kb * p.a_kb_rows);
                    tma_load_2d(a_dst + L::A_BYTES, pb, full_bar(stage), kb * TBK, n0);
                }
                if (kbt == 0 && tma_epi && ((p.residual && !r_early) || p.mask)) {
                    // the epilogue's late residual / mask tiles start their trip from DRAM now (into L2), not after the main loop
                    for (int cb = 0; cb < BN / 64 && n0 + cb * 64 < p.N; cb++) {
                        if (p.residual && !r_early) tma_prefetch_2d(&map_r, n0 + cb * 64, m0);
                        if (p.mask) tma_prefetch_2d(&map_m, n0 + cb * 64, m0);
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TBM, BN);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nk; kb++) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_addr = smem_base + stage * L::STAGE_BYTES;
                const uint64_t da = make_smem_desc(a_addr), db = make_smem_desc(a_addr + L::A_BYTES);
#pragma unroll
                for (int k = 0; k < TBK / 16; k++)          // advance 32 B (16 bf16) inside the 128 B swizzle row
                    tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                tc_commit(empty_bar(stage));                // frees this smem stage when the MMAs have read it
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit(tmem_full_bar);                       // accumulator complete
            TRACE_ADD(3, TRACE_NOW() - t_pdl);              // pdl -> all operand tiles landed, MMAs issued
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const uint32_t sbias = smem0 + L::BIAS_OFF;
        const bool leader = (warp == 2 && lane == 0);
        int nch = (p.N - n0 + 63) / 64;                      // 64-column chunks of this tile that exist (uniform across the CTA)
        if (nch > L::NCH) nch = L::NCH;
        const uint32_t early = smem0 + L::EARLY_OFF;
        if (tma_epi && r_early && p.residual && leader) {    // residual tiles: in flight together with the operand tiles
            mbar_expect_tx(epi_bar, (uint32_t)nch * 16384u);
            for (int cb = 0; cb < nch; cb++) tma_load_2d(early + (uint32_t)cb * 16384u, &map_r, epi_bar, n0 + cb * 64, m0);
        }
        if (p.bias) {                                        // global-load latency hides behind the main loop
            for (int t = threadIdx.x - 64; t < BN; t += 128) {
                const float b = (n0 + t < p.N) ? p.bias[n0 + t] : 0.f;
                asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbias + 4u * t), "f"(b) : "memory");
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < p.M;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const long long t_tf = TRACE_NOW();
        if (leader) TRACE_ADD(4, t_tf - t_pdl);              // pdl -> accumulator complete
        size_t orow = m;
        if (row_ok && p.out_stride > 1) {
            const int ohw = p.OH * p.OW;
            int b = m / ohw, rem = m - b * ohw;
            int oy = rem / p.OW, ox = rem - oy * p.OW;
            orow = ((size_t)b * p.SH + (size_t)oy * p.out_stride) * p.SW + (size_t)ox * p.out_stride;
        }
        const uint32_t thresh = dropout_thresh16(p.drop_p);
        const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
        const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
        if (tma_epi) {
            // ---------- coalesced epilogue: residual / mask tiles arrive by TMA, the bf16 result leaves by TMA store.
            // The pipeline stages are free once tmem_full fired: they take the mask (and late residual) tiles, [128 x 64] each,
            // 128B-swizzled.  Every chunk has its own tile and the result overwrites the residual (or mask) tile IN PLACE -- a
            // thread reads and writes the same 16-byte slots -- so all chunks are computed back to back and stored together.
            const bool r_late = p.residual && !r_early;
            const uint32_t lateR = smem_base, lateM = smem_base + (r_late ? (uint32_t)L::NCH * 16384u : 0u);
            if (leader && (r_late || p.mask)) {
                mbar_expect_tx(late_bar, (uint32_t)nch * 16384u * ((r_late ? 1u : 0u) + (p.mask ? 1u : 0u)));
                for (int cb = 0; cb < nch; cb++) {
                    if (r_late) tma_load_2d(lateR + (uint32_t)cb * 16384u, &map_r, late_bar, n0 + cb * 64, m0);
                    if (p.mask) tma_load_2d(lateM + (uint32_t)cb * 16384u, &map_m, late_bar, n0 + cb * 64, m0);
                }
            }
            const uint32_t bufR = r_early ? early : lateR, bufM = lateM;
            const uint32_t bufO = p.residual ? bufR : bufM;   // without residual and mask: lateM == smem_base, the free stages
            const int row = q * 32 + lane;                    // row inside the tile == TMEM lane
            const uint32_t row_off = (uint32_t)row * 128u;
            const uint32_t sw = (uint32_t)(row & 7);          // 16-byte chunk c of a row lives at chunk c ^ (row % 8)
            if (r_early && p.residual) mbar_wait(epi_bar, 0);
            if (r_late || p.mask) mbar_wait(late_bar, 0);
            const long long t_in = TRACE_NOW();
            if (leader) TRACE_ADD(5, t_in - t_tf);            // accumulator complete -> residual / mask tiles landed
            // the bit-mask code is compiled into its own copy of the chunk loop: the common (bit-free) path keeps its instruction
            // count and registers (the extra flag tests inside the 8-column groups cost the BN = 128 kernel 40 % on layer3)
            constexpr bool BITS = BITS_T || SPLIT;
            {
            // this row's mask bits of the tile's (at most two) 64-column chunks (loaded here, not before the main loop: two more live
            // registers across the accumulator wait spill in the 64-register kernels, and the co-resident CTAs hide the latency)
            uint2 mb0 = make_uint2(0u, 0u), mb1 = make_uint2(0u, 0u);
            if (BITS && p.mask_bits && row_ok) {
                const uint8_t *mp = p.mask_bits + (size_t)m * p.ldmb + (n0 >> 3);
                mb0 = ld_bits8(mp);
                if (L::NCH > 1 && nch > 1) mb1 = ld_bits8(mp + 8);
            }
#pragma unroll 1
            for (int cb = 0; cb < nch; cb++) {
                const int nb = n0 + cb * 64;
                const uint32_t cboff = (uint32_t)cb * 16384u;
                const uint2 mbc = cb == 0 ? mb0 : mb1;
                uint2 ob = make_uint2(0u, 0u);
#pragma unroll
                for (int c32 = 0; c32 < 2; c32++) {
                    uint32_t r[32];                            // two TMEM loads in flight per wait
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 64 + c32 * 32), *reinterpret_cast<uint32_t (*)[16]>(&r[0]));
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 64 + c32 * 32 + 16), *reinterpret_cast<uint32_t (*)[16]>(&r[16]));
                    tc_wait_ld();
#pragma unroll
                    for (int hf = 0; hf < 4; hf++) {
                        const int n = nb + c32 * 32 + hf * 8;
                        const uint32_t chunk = (uint32_t)(c32 * 4 + hf);
                        const uint32_t soff = cboff + row_off + ((chunk ^ sw) << 4);
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[hf * 8 + i]);
                        if (p.bias) add_bias8(v, sbias + 4u * (uint32_t)(n - n0));
                        float res[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) res[i] = 0.f;
                        if (p.residual) {
                            uint4 u;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(bufR + soff));
                            float2 t;
                            t = unpack_bf16x2(u.x); res[0] = t.x; res[1] = t.y; t = unpack_bf16x2(u.y); res[2] = t.x; res[3] = t.y;
                            t = unpack_bf16x2(u.z); res[4] = t.x; res[5] = t.y; t = unpack_bf16x2(u.w); res[6] = t.x; res[7] = t.y;
                        }
                        if (!(p.drop_p > 0.f)) {
#pragma unroll
                            for (int i = 0; i < 8; i++) v[i] += res[i];
                        }
                        if (p.relu) {
#pragma unroll
                            for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (p.mask) {
                            uint4 u;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(bufM + soff));
                            float mk[8]; float2 t;
                            t = unpack_bf16x2(u.x); mk[0] = t.x; mk[1] = t.y; t = unpack_bf16x2(u.y); mk[2] = t.x; mk[3] = t.y;
                            t = unpack_bf16x2(u.z); mk[4] = t.x; mk[5] = t.y; t = unpack_bf16x2(u.w); mk[6] = t.x; mk[7] = t.y;
#pragma unroll
                            for (int i = 0; i < 8; i++) v[i] = mk[i] > 0.f ? v[i] * p.mask_scale : 0.f;
                        }
                        if (BITS && p.mask_bits) apply_bits8(v, mbc, (int)chunk, p.mask_scale);
                        if (p.sigmoid) {
#pragma unroll
                            for (int i = 0; i < 8; i++) v[i] = 1.f / (1.f + __expf(-v[i]));
                        }
                        if (p.drop_p > 0.f) {
#pragma unroll
                            for (int i = 0; i < 8; i += 2) {
                                bool k0, k1;
                                dropout_keep2(dropout_bits(seed, p.site, (uint32_t)m, (uint32_t)((n + i) >> 1)), thresh, k0, k1);
                                v[i] = (k0 ? v[i] * drop_scale : 0.f) + res[i];
                                v[i + 1] = (k1 ? v[i + 1] * drop_scale : 0.f) + res[i + 1];
                            }
                        }
                        if (BITS && p.out_bits) collect_bits8(v, ob, (int)chunk);
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(bufO + soff), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    }
                }
                if (BITS && p.out_bits && row_ok) *reinterpret_cast<uint2 *>(p.out_bits + (size_t)m * p.ldob + (nb >> 3)) = ob;
            }
            }
            const long long t_math = TRACE_NOW();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy writes -> visible to TMA
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (leader) {
                const long long t_bar = TRACE_NOW();
                for (int cb = 0; cb < nch; cb++)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 :: "l"(&map_c), "r"(bufO + (uint32_t)cb * 16384u), "r"(n0 + cb * 64), "r"(m0) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");         // smem may be released after the reads
                TRACE_ADD(6, t_math - t_in);                  // tcgen05.ld + math + st.shared, all chunks
                TRACE_ADD(7, t_bar - t_math);                 // fence + barrier
                TRACE_ADD(8, TRACE_NOW() - t_bar);            // TMA store issue + smem read drained
            }
        } else {
        bf16 *C = reinterpret_cast<bf16 *>(p.C);
        const bf16 *R = reinterpret_cast<const bf16 *>(p.residual);
        const bf16 *Mk = reinterpret_cast<const bf16 *>(p.mask);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tc_wait_ld();
            if (!row_ok || n0 + c0 >= p.N) continue;
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                const int n = n0 + c0 + hf * 8;
                if (n >= p.N) continue;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[hf * 8 + i]);
                if (p.bias) add_bias8(v, sbias + 4u * (uint32_t)(n - n0));
                float res[8];
#pragma unroll
                for (int i = 0; i < 8; i++) res[i] = 0.f;
                if (R) sp_ld8(R + orow * p.ldr + n, split, res);
                if (!(p.drop_p > 0.f)) {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] += res[i];
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i], 0.f);
                }
                if (Mk) {
                    uint4 u = *reinterpret_cast<const uint4 *>(Mk + orow * p.ldm + n);
                    float mk[8]; float2 t;
                    t = unpack_bf16x2(u.x); mk[0] = t.x; mk[1] = t.y; t = unpack_bf16x2(u.y); mk[2] = t.x; mk[3] = t.y;
                    t = unpack_bf16x2(u.z); mk[4] = t.x; mk[5] = t.y; t = unpack_bf16x2(u.w); mk[6] = t.x; mk[7] = t.y;
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = mk[i] > 0.f ? v[i] * p.mask_scale : 0.f;
                }
                if ((BITS_T || SPLIT) && p.mask_bits) apply_bits8(v, make_uint2((uint32_t)p.mask_bits[orow * p.ldmb + (n >> 3)], 0u), 0, p.mask_scale);
                if (p.sigmoid) {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = 1.f / (1.f + __expf(-v[i]));
                }
                if (p.drop_p > 0.f) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        bool k0, k1;
                        dropout_keep2(dropout_bits(seed, p.site, (uint32_t)m, (uint32_t)((n + i) >> 1)), thresh, k0, k1);
                        v[i] = (k0 ? v[i] * drop_scale : 0.f) + res[i];
                        v[i + 1] = (k1 ? v[i + 1] * drop_scale : 0.f) + res[i + 1];
                    }
                }
                if (C) {
                    bf16 *dst = C + orow * p.ldc + n;
                    if (p.accumulate) {
                        float old[8];
                        sp_ld8(dst, split, old);
#pragma unroll
                        for (int i = 0; i < 8; i++) v[i] += old[i];
                    }
                    sp_st8(dst, split, v);
                }
                if ((BITS_T || SPLIT) && p.out_bits) {
                    uint2 ob = make_uint2(0u, 0u);
                    collect_bits8(v, ob, 0);
                    p.out_bits[orow * p.ldob + (n >> 3)] = (uint8_t)ob.x;
                }
                if (p.Cf) {
                    float4 *dst = reinterpret_cast<float4 *>(p.Cf + orow * p.ldcf + n);
                    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        }   // direct-store epilogue
    }
    // ---- teardown: everyone done with TMEM -> the allocating warp frees it
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)BN) : "memory");
        if (lane == 0) TRACE_ADD(9, TRACE_NOW() - t_entry); // whole CTA life
    }
}

// ================================================================================================ persistent variant
// One CTA per SM walks the output tiles (static round-robin); every latency chain of the one-tile kernel above is
// overlapped across tiles:
//   warp 0      TMA producer for the A / W stages (the ring runs ahead across tile boundaries)
//   warp 1      MMA issuer; NACC accumulators rotate through TMEM (tmem_full / tmem_empty barriers)
//   warp 2      TMEM allocator
//   warp 3      TMA producer for the residual / mask chunks of upcoming epilogues (ring of `rs` [128 x 64] slots)
//   warps 4-11  two epilogue warpgroups: the 64-column chunks alternate between them, each owns one swizzled staging tile
//               and its TMA stores (tcgen05.ld -> fused math -> st.shared -> cp.async.bulk.tensor store)
// BN = 256 halves the shared-memory operand traffic per flop (one 128x256x16 UMMA reads 96 B/clk instead of the 128 B/clk
// of a 128x128 tile, which saturates the SM's shared-memory port) -- the compute-bound convolutions; BN = 64 / 128 with
// 4 accumulators serve the HBM-bound 1x1 layers, whose tiles are one to four k-blocks long.
constexpr int NTHREADS_P = 384;
constexpr int MAXRS = 4;

template <int BN, int PST, int OB = 1>
struct PLayout {
    static constexpr int A_BYTES = TBM * TBK * 2;
    static constexpr int B_BYTES = BN * TBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCH = BN / 64;                                   // 64-column chunks per tile
    static constexpr int NACC = BN <= 128 ? 4 : 2;                        // accumulators in TMEM
    static constexpr int TMEM_COLS = NACC * BN;                           // 256 or 512 columns (power of two)
    static constexpr int OBUF = PST * STAGE_BYTES;                        // output staging, OB [128 x 64] tiles per warpgroup
    static constexpr int RING = OBUF + 2 * OB * 16384;                    // residual / mask ring
    static constexpr int TOTAL = 227 * 1024;
    static constexpr int BAR_OFF = TOTAL - 1024 - 1024;                   // 1 KB alignment slack, 512 B of barriers, 2 x 64 bias floats
    static constexpr int BIAS_OFF = BAR_OFF + 512;
    static constexpr int RING_BYTES = BAR_OFF - RING;
    static_assert(RING_BYTES >= 0, "persistent GEMM: stages do not fit");
};

// fused epilogue arithmetic on 8 consecutive columns n..n+7 of row m (same order as igemm.cu)
__device__ __forceinline__ void epi_math8(float (&v)[8], const float (&res)[8], const float (&mk)[8], bool has_mask,
                                          const detrb_igemm_t &p, int m, int n, uint64_t seed, uint32_t thresh, float drop_scale,
                                          uint32_t sbias8, const uint2 &mbits, int c)
{
    if (p.bias) add_bias8(v, sbias8);
    if (!(p.drop_p > 0.f)) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] += res[i];
    }
    if (p.relu) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i], 0.f);
    }
    if (has_mask) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = mk[i] > 0.f ? v[i] * p.mask_scale : 0.f;
    }
    if (p.mask_bits) apply_bits8(v, mbits, c, p.mask_scale);
    if (p.sigmoid) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 1.f / (1.f + __expf(-v[i]));
    }
    if (p.drop_p > 0.f) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            bool k0, k1;
            dropout_keep2(dropout_bits(seed, p.site, (uint32_t)m, (uint32_t)((n + i) >> 1)), thresh, k0, k1);
            v[i] = (k0 ? v[i] * drop_scale : 0.f) + res[i];
            v[i + 1] = (k1 ? v[i + 1] * drop_scale : 0.f) + res[i + 1];
        }
    }
}
// this row's mask bits of tile `tile` (NCH chunks of 8 bytes), zero past the edges
template <int NCH>
__device__ __forceinline__ void load_tile_bits(uint2 (&b)[NCH], const detrb_igemm_t &p, int tile, int n_tiles, int n_tiles_n, int row)
{
#pragma unroll
    for (int i = 0; i < NCH; i++) b[i] = make_uint2(0u, 0u);
    if (tile >= n_tiles) return;
    const int m = (tile / n_tiles_n) * TBM + row, n0 = (tile % n_tiles_n) * (NCH * 64);
    if (m >= p.M) return;
    const uint8_t *mp = p.mask_bits + (size_t)m * p.ldmb + (n0 >> 3);
#pragma unroll
    for (int i = 0; i < NCH; i++)
        if (n0 + i * 64 < p.N) b[i] = ld_bits8(mp + 8 * i);
}

// OB = 2: two output staging tiles per warpgroup -- the TMA store of chunk i drains while chunk i+1 is computed.  Measured on the
// HBM-bound 1x1 layers (K <= 128): 15.37 vs 14.81 ms/step with the one-tile kernel, so only OB = 1 is instantiated.
template <int BN, int PST, bool IM2COL, int OB = 1>
__global__ void __launch_bounds__(NTHREADS_P, 1)
gemm_tcp_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
                const __grid_constant__ CUtensorMap map_m, const detrb_igemm_t p, const ConvAux aux,
                const int n_tiles_n, const int n_tiles, const int rs)
{
    using L = PLayout<BN, PST, OB>;
    constexpr int NACC = L::NACC;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (PST + s); };
    auto tmem_full = [&](int a) { return bar_base + 8u * (2 * PST + a); };
    auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * PST + NACC + a); };
    auto resid_full = [&](int i) { return bar_base + 8u * (2 * PST + 2 * NACC + i); };
    auto resid_empty = [&](int i) { return bar_base + 8u * (2 * PST + 2 * NACC + MAXRS + i); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * PST + 2 * NACC + 2 * MAXRS);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = p.K / TBK;
    const bool has_r = p.residual != nullptr, has_m = p.mask != nullptr, have_in = has_r || has_m;
    const uint32_t slot_bytes = 16384u * ((has_r ? 1u : 0u) + (has_m ? 1u : 0u));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_c);
        if (has_r) tma_prefetch_desc(&map_r);
        if (has_m) tma_prefetch_desc(&map_m);
        for (int s = 0; s < PST; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 8); }
        for (int i = 0; i < MAXRS; i++) { mbar_init(resid_full(i), 1); mbar_init(resid_empty(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();          // everything above (barriers, TMEM) overlapped the previous kernel's tail

    if (warp == 0) {
        // ===================== A / W producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * TBM, n0 = (tile % n_tiles_n) * BN;
                int w0 = 0, h0 = 0, img = 0;
                if (IM2COL) {
                    const int ohw = p.OH * p.OW;
                    img = m0 / ohw;
                    const int rem = m0 - img * ohw, oy = rem / p.OW, ox = rem - oy * p.OW;
                    const int st = p.mode == 0 ? p.stride : 1;
                    w0 = ox * st - aux.pad; h0 = oy * st - aux.pad;
                }
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
                    const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
                    if (IM2COL) {
                        const int k0 = kb * TBK, tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
                        const int kh = tap / p.KW, kw = tap - kh * p.KW;
                        tma_load_im2col(a_dst, &map_a, full_bar(stage), c0, w0, h0, img, (uint16_t)kw, (uint16_t)kh);
                        tma_load_2d(a_dst + L::A_BYTES, &map_b, full_bar(stage), aux.wtap[tap] * p.Cin + c0, n0);
                    } else {
                        tma_load_2d(a_dst, &map_a, full_bar(stage), p.a_kb_rows ? 0 : kb * TBK, m0 + kb * p.a_kb_rows);
                        tma_load_2d(a_dst + L::A_BYTES, &map_b, full_bar(stage), kb * TBK, n0);
                    }
                    if (++stage == PST) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TBM, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
                const int acc = it % NACC;
                mbar_wait(tmem_empty(acc), ((it / NACC) & 1) ^ 1);        // both epilogue warpgroups have drained this accumulator
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * L::STAGE_BYTES;
                    const uint64_t da = make_smem_desc(a_addr), db = make_smem_desc(a_addr + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < TBK / 16; k++)
                        tc_mma_f16(d_addr, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(empty_bar(stage));
                    if (++stage == PST) { stage = 0; phase ^= 1; }
                }
                tc_commit(tmem_full(acc));
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== residual / mask producer (one ring slot per 64-column chunk) =====================
        if (lane == 0 && have_in) {
            int g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * TBM, n0 = (tile % n_tiles_n) * BN;
                const int left = (p.N - n0 + 63) / 64, nch = left < L::NCH ? left : L::NCH;
                for (int cb = 0; cb < nch; cb++, g++) {
                    const int slot = g % rs;
                    mbar_wait(resid_empty(slot), ((g / rs) & 1) ^ 1);
                    mbar_expect_tx(resid_full(slot), slot_bytes);
                    const uint32_t dst = smem_base + L::RING + (uint32_t)slot * slot_bytes;
                    if (has_r) tma_load_2d(dst, &map_r, resid_full(slot), n0 + cb * 64, m0);
                    if (has_m) tma_load_2d(dst + (has_r ? 16384u : 0u), &map_m, resid_full(slot), n0 + cb * 64, m0);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: chunk g goes to warpgroup g % 2 =====================
        const int wg = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
        const bool leader = (q == 0 && lane == 0);
        const uint32_t obuf0 = smem_base + L::OBUF + (uint32_t)wg * (uint32_t)(OB * 16384);
        int oc = 0;                                             // chunks this warpgroup has produced (selects its staging tile)
        const uint32_t sbias = smem_base + L::BIAS_OFF + (uint32_t)wg * 256u;
        const uint32_t thresh = dropout_thresh16(p.drop_p);
        const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
        const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
        int it = 0, g = 0;
        const bool fast = !has_m && !p.sigmoid && p.N % 64 == 0;               // bias / residual / ReLU / bit masks / dropout: the straight-line chunk
        uint2 mb_cur[L::NCH], mb_nxt[L::NCH];                   // 1-bit mask of this row: current tile, next tile (prefetched)
        if (p.mask_bits) load_tile_bits<L::NCH>(mb_nxt, p, blockIdx.x, n_tiles, n_tiles_n, row);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int m0 = (tile / n_tiles_n) * TBM, n0 = (tile % n_tiles_n) * BN;
            const int acc = it % NACC;
            const int m = m0 + row;
            if (p.mask_bits) {
#pragma unroll
                for (int i = 0; i < L::NCH; i++) mb_cur[i] = mb_nxt[i];
                load_tile_bits<L::NCH>(mb_nxt, p, tile + (int)gridDim.x, n_tiles, n_tiles_n, row);
            }
            const int left = (p.N - n0 + 63) / 64, nch = left < L::NCH ? left : L::NCH;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            // every epilogue warp observes tmem_full before it arrives on tmem_empty below -- also a warpgroup without a chunk
            // in this tile -- so no warp can run ahead and arrive twice in one phase of tmem_empty
            mbar_wait(tmem_full(acc), (it / NACC) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < nch; cb++, g++) {
                if ((g & 1) != wg) continue;
                const int nb = n0 + cb * 64;
                int slot = 0;
                uint32_t rbuf = 0, mbuf = 0;
                if (have_in) {
                    slot = g % rs;
                    mbar_wait(resid_full(slot), (g / rs) & 1);
                    rbuf = smem_base + L::RING + (uint32_t)slot * slot_bytes;
                    mbuf = rbuf + (has_r ? 16384u : 0u);
                }
                if ((p.bias || fast) && q < 2) {                // this chunk's 64 bias values -> smem (visible after the barrier below)
                    const int t = q * 32 + lane;
                    const float b = (p.bias && nb + t < p.N) ? p.bias[nb + t] : 0.f;
                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbias + 4u * t), "f"(b) : "memory");
                }
                // the TMA store that last read this staging tile (OB chunks ago) must have finished reading it
                const uint32_t obuf = obuf0 + (uint32_t)((oc % OB) * 16384);
                oc++;
                if (leader) {
                    if (OB == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
                uint2 mbc = make_uint2(0u, 0u), ob = make_uint2(0u, 0u);
                if (p.mask_bits) {
                    mbc = mb_cur[0];
#pragma unroll
                    for (int i = 1; i < L::NCH; i++) if (cb == i) mbc = mb_cur[i];
                }
                if (fast) {                                     // bias / residual / ReLU / bit masks only: the straight-line chunk
                    uint32_t acc_r[64];
                    tc_ld64(t_addr + (uint32_t)(cb * 64), acc_r);
                    uint4 rr[8];
                    if (has_r) lds_row8(rr, rbuf + row_off, sw);
                    tc_wait_ld();
                    const bool relu = p.relu != 0;
                    if (p.drop_p > 0.f) {
                        EpiDrop dr;
                        dr.rowhash = dropout_rowhash(seed, p.site, (uint32_t)m); dr.thresh = thresh; dr.col0 = (uint32_t)nb; dr.scale = drop_scale;
                        if (has_r) epi_chunk_math<true, false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw, dr);
                        else epi_chunk_math<false, false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw, dr);
                    } else if (has_r) {
                        if (p.mask_bits) epi_chunk_math<true, true, false>(acc_r, rr, sbias, relu, mbc, p.mask_scale, ob, obuf + row_off, sw);
                        else if (p.out_bits) epi_chunk_math<true, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                        else epi_chunk_math<true, false, false>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                    } else {
                        if (p.mask_bits) epi_chunk_math<false, true, false>(acc_r, rr, sbias, relu, mbc, p.mask_scale, ob, obuf + row_off, sw);
                        else if (p.out_bits) epi_chunk_math<false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                        else epi_chunk_math<false, false, false>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                    }
                } else
#pragma unroll
                for (int c32 = 0; c32 < 2; c32++) {
                    uint32_t r[32];                            // two TMEM loads in flight per wait
                    tc_ld16(t_addr + (uint32_t)(cb * 64 + c32 * 32), *reinterpret_cast<uint32_t (*)[16]>(&r[0]));
                    tc_ld16(t_addr + (uint32_t)(cb * 64 + c32 * 32 + 16), *reinterpret_cast<uint32_t (*)[16]>(&r[16]));
                    tc_wait_ld();
#pragma unroll
                    for (int hf = 0; hf < 4; hf++) {
                        const int n = nb + c32 * 32 + hf * 8;
                        const uint32_t soff = row_off + (((uint32_t)(c32 * 4 + hf) ^ sw) << 4);
                        float v[8], res[8], mk[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) { v[i] = __uint_as_float(r[hf * 8 + i]); res[i] = 0.f; mk[i] = 1.f; }
                        if (has_r) {
                            uint4 u;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(rbuf + soff));
                            unpack8(u, res);
                        }
                        if (has_m) {
                            uint4 u;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(mbuf + soff));
                            unpack8(u, mk);
                        }
                        if (n < p.N) epi_math8(v, res, mk, has_m, p, m, n, seed, thresh, drop_scale, sbias + 4u * (uint32_t)(n - nb), mbc, c32 * 4 + hf);
                        if (p.out_bits) collect_bits8(v, ob, c32 * 4 + hf);
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(obuf + soff), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    }
                }
                if (p.out_bits && m < p.M) *reinterpret_cast<uint2 *>(p.out_bits + (size_t)m * p.ldob + (nb >> 3)) = ob;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
                if (leader) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 :: "l"(&map_c), "r"(obuf), "r"(nb), "r"(m0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (have_in && lane == 0) mbar_arrive(resid_empty(slot));      // all four warps are past their smem reads
            }
            // this warp is done with the accumulator of this tile (also when its warpgroup had no chunk in it)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(acc));
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)L::TMEM_COLS) : "memory");
    }
}

// ================================================================================================ CTA-pair variant
// gemm_pair_kernel: the persistent kernel above on a CTA PAIR (cluster of two SMs of one TPC, tcgen05 cta_group::2).  One pair owns a
// 256 x BN tile: each CTA stages ITS 128 rows of A and ITS half (BN / 2 weight rows) of W, the leader CTA's single thread issues
// 256 x BN x 16 MMAs that read both CTAs' shared memory and write both CTAs' tensor memory (128 lanes each).  Per SM the operand
// traffic of one k-step is 4 KB (A) + BN / 2 rows of W instead of BN rows: 64 B/clk at BN = 256 where the single-CTA 128 x 256
// tile needs 96 B/clk -- the shared-memory operand fetch (~64-70 B/clk measured) is what held that kernel at ~0.62 of the tensor
// peak.  Barrier protocol (CUTLASS sm100 2-SM pipelines): the smem-full barriers live in the LEADER (both CTAs' TMA loads signal
// them through the shared::cluster window, peer bit cleared; the leader's producer expects the bytes of both), smem-empty and
// tmem-full are multicast to both CTAs by tcgen05.commit, tmem-empty is the leader's and takes the arrivals of both CTAs'
// epilogue warps.  Epilogue: the straight-line chunk (bias / residual / ReLU / bit masks / dropout), per CTA as above.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;             // shared::cluster address of the same offset in CTA 0 of the pair

template <int BN, int PST>
struct PairLayout {
    static constexpr int A_BYTES = TBM * TBK * 2;
    static constexpr int B_BYTES = (BN / 2) * TBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCH = BN / 64;
    static constexpr int NACC = 512 / BN;                                 // 2 (BN = 256) or 4 (BN = 128) accumulators: all of TMEM
    static constexpr int OBUF = PST * STAGE_BYTES;                        // one [128 x 64] output staging tile per warpgroup
    static constexpr int RING = OBUF + 2 * 16384;                         // residual ring (two slots)
    static constexpr int TOTAL = 227 * 1024;
    static constexpr int BAR_OFF = TOTAL - 1024 - 1024;
    static constexpr int BIAS_OFF = BAR_OFF + 512;
    static_assert(BAR_OFF - RING >= 2 * 16384, "pair GEMM: stages do not fit");
};

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c, int w, int h, int n,
                                                     uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 :: "r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in CTA 0 of the pair (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BN, int PST, bool IM2COL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS_P, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
                 const detrb_igemm_t p, const ConvAux aux, const int n_tiles_n, const int n_tiles)
{
    using L = PairLayout<BN, PST>;
    constexpr int NACC = L::NACC, RS = 2;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (PST + s); };
    auto tmem_full = [&](int a) { return bar_base + 8u * (2 * PST + a); };
    auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * PST + NACC + a); };
    auto resid_full = [&](int i) { return bar_base + 8u * (2 * PST + 2 * NACC + i); };
    auto resid_empty = [&](int i) { return bar_base + 8u * (2 * PST + 2 * NACC + RS + i); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * PST + 2 * NACC + 2 * RS);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nk = p.K / TBK;
    const bool has_r = p.residual != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_c);
        if (has_r) tma_prefetch_desc(&map_r);
        for (int s = 0; s < PST; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 16); }
        for (int i = 0; i < RS; i++) { mbar_init(resid_full(i), 1); mbar_init(resid_empty(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    cluster_sync_all();                 // both CTAs' barriers initialised (the peer signals the leader's), TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===================== A / W producer (both CTAs: own 128 rows of A, own half of W) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile / n_tiles_n) * (2 * TBM) + (int)rank * TBM, n0 = (tile % n_tiles_n) * BN + (int)rank * (BN / 2);
                int w0 = 0, h0 = 0, img = 0;
                if (IM2COL) {
                    const int ohw = p.OH * p.OW;
                    img = m0 / ohw;
                    const int rem = m0 - img * ohw, oy = rem / p.OW, ox = rem - oy * p.OW;
                    const int st = p.mode == 0 ? p.stride : 1;
                    w0 = ox * st - aux.pad; h0 = oy * st - aux.pad;
                }
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
                    const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
                    if (IM2COL) {
                        const int k0 = kb * TBK, tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
                        const int kh = tap / p.KW, kw = tap - kh * p.KW;
                        tma_load_im2col_pair(a_dst, &map_a, full_bar(stage), c0, w0, h0, img, (uint16_t)kw, (uint16_t)kh);
                        tma_load_2d_pair(a_dst + L::A_BYTES, &map_b, full_bar(stage), aux.wtap[tap] * p.Cin + c0, n0);
                    } else {
                        tma_load_2d_pair(a_dst, &map_a, full_bar(stage), kb * TBK, m0);
                        tma_load_2d_pair(a_dst + L::A_BYTES, &map_b, full_bar(stage), kb * TBK, n0);
                    }
                    if (++stage == PST) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(2 * TBM, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, it++) {
                const int acc = it % NACC;
                mbar_wait(tmem_empty(acc), ((it / NACC) & 1) ^ 1);        // the epilogue warps of BOTH CTAs have drained this accumulator
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * L::STAGE_BYTES;
                    const uint64_t da = make_smem_desc(a_addr), db = make_smem_desc(a_addr + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < TBK / 16; k++)
                        tc_mma_f16_pair(d_addr, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit_pair(empty_bar(stage));
                    if (++stage == PST) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(tmem_full(acc));
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== residual producer (one ring slot per 64-column chunk of this CTA's 128 rows) =====================
        if (lane == 0 && has_r) {
            int g = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile / n_tiles_n) * (2 * TBM) + (int)rank * TBM, n0 = (tile % n_tiles_n) * BN;
                for (int cb = 0; cb < L::NCH; cb++, g++) {
                    const int slot = g % RS;
                    mbar_wait(resid_empty(slot), ((g / RS) & 1) ^ 1);
                    mbar_expect_tx(resid_full(slot), 16384u);
                    tma_load_2d(smem_base + L::RING + (uint32_t)slot * 16384u, &map_r, resid_full(slot), n0 + cb * 64, m0);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: chunk g goes to warpgroup g % 2 =====================
        const int wg = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
        const bool leader = (q == 0 && lane == 0);
        const uint32_t obuf = smem_base + L::OBUF + (uint32_t)wg * 16384u;
        const uint32_t sbias = smem_base + L::BIAS_OFF + (uint32_t)wg * 256u;
        const uint32_t thresh = dropout_thresh16(p.drop_p);
        const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
        const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
        int it = 0, g = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs, it++) {
            const int m0 = (tile / n_tiles_n) * (2 * TBM) + (int)rank * TBM, n0 = (tile % n_tiles_n) * BN;
            const int acc = it % NACC;
            const int m = m0 + row;
            uint2 mb_cur[L::NCH];
#pragma unroll
            for (int i = 0; i < L::NCH; i++) mb_cur[i] = make_uint2(0u, 0u);
            if (p.mask_bits && m < p.M) {
                const uint8_t *mp = p.mask_bits + (size_t)m * p.ldmb + (n0 >> 3);
#pragma unroll
                for (int i = 0; i < L::NCH; i++) mb_cur[i] = ld_bits8(mp + 8 * i);
            }
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            mbar_wait(tmem_full(acc), (it / NACC) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < L::NCH; cb++, g++) {
                if ((g & 1) != wg) continue;
                const int nb = n0 + cb * 64;
                int slot = 0;
                uint32_t rbuf = 0;
                if (has_r) {
                    slot = g % RS;
                    mbar_wait(resid_full(slot), (g / RS) & 1);
                    rbuf = smem_base + L::RING + (uint32_t)slot * 16384u;
                }
                if (q < 2) {                                    // this chunk's 64 bias values -> smem (visible after the barrier below)
                    const int t = q * 32 + lane;
                    const float b = (p.bias && nb + t < p.N) ? p.bias[nb + t] : 0.f;
                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbias + 4u * t), "f"(b) : "memory");
                }
                if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the last store has read the staging tile
                asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
                uint2 mbc = mb_cur[0], ob = make_uint2(0u, 0u);
#pragma unroll
                for (int i = 1; i < L::NCH; i++) if (cb == i) mbc = mb_cur[i];
                uint32_t acc_r[64];
                tc_ld64(t_addr + (uint32_t)(cb * 64), acc_r);
                uint4 rr[8];
                if (has_r) lds_row8(rr, rbuf + row_off, sw);
                tc_wait_ld();
                const bool relu = p.relu != 0;
                if (p.drop_p > 0.f) {
                    EpiDrop dr;
                    dr.rowhash = dropout_rowhash(seed, p.site, (uint32_t)m); dr.thresh = thresh; dr.col0 = (uint32_t)nb; dr.scale = drop_scale;
                    if (has_r) epi_chunk_math<true, false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw, dr);
                    else epi_chunk_math<false, false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw, dr);
                } else if (has_r) {
                    if (p.mask_bits) epi_chunk_math<true, true, false>(acc_r, rr, sbias, relu, mbc, p.mask_scale, ob, obuf + row_off, sw);
                    else if (p.out_bits) epi_chunk_math<true, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                    else epi_chunk_math<true, false, false>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                } else {
                    if (p.mask_bits) epi_chunk_math<false, true, false>(acc_r, rr, sbias, relu, mbc, p.mask_scale, ob, obuf + row_off, sw);
                    else if (p.out_bits) epi_chunk_math<false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                    else epi_chunk_math<false, false, false>(acc_r, rr, sbias, relu, mbc, 1.f, ob, obuf + row_off, sw);
                }
                if (p.out_bits && m < p.M) *reinterpret_cast<uint2 *>(p.out_bits + (size_t)m * p.ldob + (nb >> 3)) = ob;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
                if (leader) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 :: "l"(&map_c), "r"(obuf), "r"(nb), "r"(m0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (has_r && lane == 0) mbar_arrive(resid_empty(slot));
            }
            // this warp is done with the accumulator of this tile: one arrival on the LEADER's barrier (16 = 8 warps x 2 CTAs)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tmem_empty(acc));
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                 // no CTA of the pair leaves (or frees TMEM) while the other may still signal it or read its smem
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ================================================================================================ streaming variant
// The HBM-bound 1x1 layers of layer1 / layer2 (M = 133600 .. 534400 rows, K and N in 64 .. 512): 2-8 flops per byte, the job is
// to keep the memory system full.  One persistent CTA per SM owns a fixed column range of BN <= 256 columns and streams 128-row
// tiles of A through it:
//   * the WEIGHTS of the column range are loaded once and stay in shared memory (<= 64 KB) -- the other kernels re-fetch them per tile;
//   * a ring of `nst` 16 KB A stages runs ahead across tile boundaries (two or more tiles in flight);
//   * a ring of `rs` [128 x 64] slots carries each output chunk through its whole life IN PLACE: residual tile in by TMA (warp 3) ->
//     fused epilogue overwrites it -> TMA store from the same slot -> released when the store has read it (one chunk later, so the
//     epilogue warps never wait for a store); without a residual the slots are plain output staging;
//   * ReLU masks arrive as bits in registers, prefetched one tile ahead (no mask tiles in shared memory at all).
// warp 0: A (+W) producer, warp 1: MMA issuer (NACC accumulators rotate through TMEM), warp 2: TMEM allocator, warp 3: residual
// producer, warps 4-11: two epilogue warpgroups alternating over the 64-column chunks.
constexpr int ST_BAR_BYTES = 2048;                       // 512 B of barriers, up to 256 bias floats
constexpr int ST_MAX_NST = 12, ST_MAX_RS = 10;

template <int BN, bool HAS_R, bool MBITS, bool OBITS>
__global__ void __launch_bounds__(NTHREADS_P, 1)
gemm_stream_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r, const detrb_igemm_t p,
                   const int n_parts, const int m_tiles, const int nst, const int rs, const int diag)
{
    // diag (developer switch DETRB_STREAM_DIAG, wrong results, timing only): 1 no TMA stores, 2 no epilogue arithmetic, 4 no tcgen05.ld
    constexpr int NCH = BN / 64;
    constexpr int NACC = BN == 256 ? 2 : 4;
    constexpr int TMEM_COLS = NACC * BN;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nk = p.K / TBK;
    const uint32_t w_bytes = (uint32_t)(BN * TBK * 2) * (uint32_t)nk;
    const uint32_t a_base = smem_base + w_bytes, r_base = a_base + (uint32_t)nst * 16384u;
    const uint32_t bar_base = smem_base + (uint32_t)(227 * 1024 - 1024 - ST_BAR_BYTES);
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (ST_MAX_NST + s); };
    auto tmem_full = [&](int a) { return bar_base + 8u * (2 * ST_MAX_NST + a); };
    auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * ST_MAX_NST + 4 + a); };
    auto slot_full = [&](int i) { return bar_base + 8u * (2 * ST_MAX_NST + 8 + i); };
    auto slot_free = [&](int i) { return bar_base + 8u * (2 * ST_MAX_NST + 8 + ST_MAX_RS + i); };
    const uint32_t w_full = bar_base + 8u * (2 * ST_MAX_NST + 8 + 2 * ST_MAX_RS);
    const uint32_t tmem_slot = w_full + 8u;
    const uint32_t sbias = bar_base + 512u;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool has_r = HAS_R;
    const int part = blockIdx.x % n_parts, first_tile = blockIdx.x / n_parts, tile_step = gridDim.x / n_parts;
    const int n0 = part * BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_c);
        if (has_r) tma_prefetch_desc(&map_r);
        for (int s = 0; s < nst; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int a = 0; a < NACC; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 8); }
        for (int i = 0; i < rs; i++) { mbar_init(slot_full(i), 1); mbar_init(slot_free(i), 1); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===================== weights once, then the A stream =====================
        if (lane == 0) {
            mbar_expect_tx(w_full, w_bytes);
            for (int kb = 0; kb < nk; kb++) tma_load_2d(smem_base + (uint32_t)kb * (uint32_t)(BN * 128), &map_b, w_full, kb * TBK, n0);
            int stage = 0; uint32_t phase = 0;
            for (int t = first_tile; t < m_tiles; t += tile_step) {
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(a_empty(stage), phase ^ 1);
                    mbar_expect_tx(a_full(stage), 16384u);
                    // sliding-window A (the space-to-depth stem): k-block kb is the 64-element run a_kb_rows rows further down
                    tma_load_2d(a_base + (uint32_t)stage * 16384u, &map_a, a_full(stage), p.a_kb_rows ? 0 : kb * TBK,
                                t * TBM + kb * p.a_kb_rows);
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TBM, BN);
            mbar_wait(w_full, 0);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int t = first_tile; t < m_tiles; t += tile_step, it++) {
                const int acc = it % NACC;
                mbar_wait(tmem_empty(acc), ((it / NACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(a_full(stage), phase);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc(a_base + (uint32_t)stage * 16384u);
                    const uint64_t db = make_smem_desc(smem_base + (uint32_t)kb * (uint32_t)(BN * 128));
#pragma unroll
                    for (int k = 0; k < TBK / 16; k++)
                        tc_mma_f16(d_addr, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(a_empty(stage));
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
                tc_commit(tmem_full(acc));
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== residual producer: chunk g -> slot g % rs =====================
        if (lane == 0 && has_r) {
            int g = 0;
            for (int t = first_tile; t < m_tiles; t += tile_step) {
                for (int cb = 0; cb < NCH; cb++, g++) {
                    const int slot = g % rs;
                    mbar_wait(slot_free(slot), ((g / rs) & 1) ^ 1);
                    mbar_expect_tx(slot_full(slot), 16384u);
                    tma_load_2d(r_base + (uint32_t)slot * 16384u, &map_r, slot_full(slot), n0 + cb * 64, t * TBM);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: chunk g goes to warpgroup g % 2 =====================
        const int wg = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
        const bool leader = (q == 0 && lane == 0);
        // the CTA's column range is fixed: its bias values (zeros without a bias) go to shared memory once
        for (int t = threadIdx.x - 128; t < BN; t += 256)
            asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbias + 4u * (uint32_t)t), "f"(p.bias ? p.bias[n0 + t] : 0.f) : "memory");
        asm volatile("bar.sync 3, 256;" ::: "memory");
        const bool relu = p.relu != 0;
        const float mscale = p.mask_scale;
        const bool drop = p.drop_p > 0.f;
        const uint32_t thresh = dropout_thresh16(p.drop_p);
        const float drop_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
        const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
        uint2 mb_cur[NCH], mb_nxt[NCH];
        auto load_bits = [&](uint2 (&b)[NCH], int t) {
#pragma unroll
            for (int i = 0; i < NCH; i++) b[i] = make_uint2(0u, 0u);
            const int m = t * TBM + row;
            if (t >= m_tiles || m >= p.M) return;
            const uint8_t *mp = p.mask_bits + (size_t)m * p.ldmb + (n0 >> 3);
#pragma unroll
            for (int i = 0; i < NCH; i++) b[i] = ld_bits8(mp + 8 * i);
        };
        if (MBITS) load_bits(mb_nxt, first_tile);
        int it = 0, g = 0, prev_slot = -1;
        for (int t = first_tile; t < m_tiles; t += tile_step, it++) {
            const int m0 = t * TBM, m = m0 + row;
            const int acc = it % NACC;
            if (MBITS) {
#pragma unroll
                for (int i = 0; i < NCH; i++) mb_cur[i] = mb_nxt[i];
                load_bits(mb_nxt, t + tile_step);
            }
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            mbar_wait(tmem_full(acc), (it / NACC) & 1);
            tc_fence_after();
            bool released = false;
#pragma unroll 1
            for (int cb = 0; cb < NCH; cb++, g++) {
                if ((g & 1) != wg) continue;
                const int nb = n0 + cb * 64;
                const int slot = g % rs;
                const uint32_t sbuf = r_base + (uint32_t)slot * 16384u;
                // the whole chunk is straight-line code (no data-dependent or flag branches): every shared-memory and tensor-memory
                // load of the chunk is in flight before the first use -- with branches between the 8-column groups each group was a
                // serial load -> use chain at loaded shared-memory latency, and the epilogue, not HBM, set the pace (170 vs 102 us)
                uint32_t acc_r[64];
                tc_ld64(t_addr + (uint32_t)(cb * 64), acc_r);
                if (has_r) mbar_wait(slot_full(slot), (g / rs) & 1);           // residual tile landed
                else mbar_wait(slot_free(slot), ((g / rs) & 1) ^ 1);           // the store that last used this slot has read it
                uint4 rr[8];
                if (HAS_R) lds_row8(rr, sbuf + row_off, sw);
                uint2 mbc = make_uint2(0u, 0u), ob = make_uint2(0u, 0u);
                if (MBITS) {
                    mbc = mb_cur[0];
#pragma unroll
                    for (int i = 1; i < NCH; i++) if (cb == i) mbc = mb_cur[i];
                }
                tc_wait_ld();
                // this warpgroup's last chunk of the tile (chunks alternate between the two warpgroups): the accumulator goes back to
                // the MMA warp as soon as it is in registers, not after the stores
                if (cb + 2 >= NCH) {
                    released = true;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty(acc));
                }
                if (!MBITS && !OBITS && drop) {
                    EpiDrop dr;
                    dr.rowhash = dropout_rowhash(seed, p.site, (uint32_t)m); dr.thresh = thresh; dr.col0 = (uint32_t)nb; dr.scale = drop_scale;
                    epi_chunk_math<HAS_R, false, false, true>(acc_r, rr, sbias + 4u * (uint32_t)(cb * 64), relu, mbc, mscale, ob, sbuf + row_off, sw, dr);
                } else if (!(diag & 2)) epi_chunk_math<HAS_R, MBITS, OBITS>(acc_r, rr, sbias + 4u * (uint32_t)(cb * 64), relu, mbc, mscale, ob, sbuf + row_off, sw);
                if (OBITS && m < p.M) *reinterpret_cast<uint2 *>(p.out_bits + (size_t)m * p.ldob + (nb >> 3)) = ob;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
                if (leader) {
                    if (!(diag & 1))
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     :: "l"(&map_c), "r"(sbuf), "r"(nb), "r"(m0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (prev_slot >= 0) {                     // the previous chunk's store has read its slot by now (or we wait for it)
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        mbar_arrive(slot_free(prev_slot));
                    }
                    prev_slot = slot;
                }
            }
            if (!released) {                                  // a warpgroup without a chunk in this tile (BN = 64: every other tile)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty(acc));
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128-byte swizzle, zero OOB fill
bool make_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols = TBK,
              int swizzle_bytes = 128)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                                  : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static thread_local int g_tma_epilogue = 1;          // (the public switches are thread-local: detrb_bind installs a handle's values)
static int g_r_early = 1;            // early residual fetch: 0 never, 1 policy (launch_tc), 2 always      (env DETRB_R_EARLY)
static int g_deep_small = 1;         // one stage per k-block for K <= 256 on latency-bound grids           (env DETRB_DEEP_SMALL)
static int g_one_stage = 1;          // one-stage kernel for K = 64                                          (env DETRB_ONE_STAGE)
static thread_local int g_tc_persistent = 1;      // 1: auto policy (dispatch_tcp)
static int g_tc_pair = 0;            // CTA-pair kernel: 0 off, 1 for 256-wide tiles, 2 for 128- and 256-wide tiles   (env DETRB_PAIR)
static thread_local int g_tc_pair_forced = -1;    // detrb_set_tc_pair() overrides the environment
static long g_tcp_min_tiles = 64, g_tcp_min_tiles256 = 100, g_tcp_min_nk256 = 12;     // auto policy thresholds (env DETRB_TCP_MIN_TILES / DETRB_TCP_MIN_NK256)

struct ConvClass { ConvAux aux; int upper_w, upper_h, w_cols; };

template <int BN, int STAGES, bool IM2COL, bool PERSIST = false, int OB = 1, bool SPLIT = false, bool BITS_T = false>
int launch_tc(const detrb_igemm_t &p, cudaStream_t stream, const ConvClass *cls = nullptr)
{
    static_assert(!(PERSIST && SPLIT), "parity precision runs on the one-tile kernel");
    CUtensorMap ma, mb, ma2, mb2;
    ConvAux aux;
    aux.pad = 0;
    for (int i = 0; i < 16; i++) aux.wtap[i] = i;
    uint64_t w_cols = (uint64_t)p.K;                        // columns of the weight matrix the W tensor map spans
    if (IM2COL) {
        const int st = p.mode == 0 ? p.stride : 1;
        int lower_w, lower_h, upper_w, upper_h;
        if (cls) {
            // one parity class of a stride-2 data gradient: plain window walk over dY (pad 0), sparse tap map (see below)
            aux = cls->aux;
            lower_w = lower_h = 0; upper_w = cls->upper_w; upper_h = cls->upper_h;
            w_cols = (uint64_t)cls->w_cols;
        } else {
            // forward: window origin = out*stride - pad.  transposed (stride 1): convolution of dY with the flipped kernel, pad' = K-1-pad
            aux.pad = p.mode == 0 ? p.pad : (p.KH - 1 - p.pad);
            if (p.mode == 1) for (int i = 0; i < p.KH * p.KW; i++) aux.wtap[i] = p.KH * p.KW - 1 - i;
            // the bounding box must yield exactly OW x OH window origins: (I + upper - lower - 1) / st + 1 == O
            // (also covers asymmetric padding: the space-to-depth stem is a 4x4 window with 2 rows above and 1 below)
            lower_w = lower_h = -aux.pad;
            upper_w = (p.OW - 1) * st + 1 + lower_w - p.IW;
            upper_h = (p.OH - 1) * st + 1 + lower_h - p.IH;
        }
        if (upper_w > 0 || upper_h > 0 || upper_w < -16 || upper_h < -16)
            DETRB_FAIL(DETRB_E_SHAPE, "gemm_tc im2col: inconsistent conv geometry IH=%d IW=%d OH=%d OW=%d k=%dx%d s=%d p=%d", p.IH, p.IW,
                       p.OH, p.OW, p.KH, p.KW, st, aux.pad);
        int rc = detrb_make_im2col_map(&ma, p.A, p.batch, p.IH, p.IW, p.Cin, p.lda, lower_w, lower_h, upper_w, upper_h, st, TBM, 1, 64);
        if (rc) return rc;
        if (SPLIT) {
            rc = detrb_make_im2col_map(&ma2, p.A + p.split, p.batch, p.IH, p.IW, p.Cin, p.lda, lower_w, lower_h, upper_w, upper_h, st, TBM, 1, 64);
            if (rc) return rc;
        }
    } else if (p.a_kb_rows > 0) {
        // sliding-window operand: every row is one 64-element run, consecutive rows start lda elements apart (they overlap when
        // lda < 64); the map spans the rows the last k-block of the last tile row can reach
        const uint64_t rows = (uint64_t)p.M + (uint64_t)(p.K / TBK - 1) * (uint64_t)p.a_kb_rows;
        if (!make_map(&ma, p.A, rows, TBK, (uint64_t)p.lda, TBM) || (SPLIT && !make_map(&ma2, p.A + p.split, rows, TBK, (uint64_t)p.lda, TBM)))
            DETRB_FAIL(DETRB_E_CUDA, "gemm_tc: cuTensorMapEncodeTiled(sliding A) failed (rows=%llu lda=%d)", (unsigned long long)rows, p.lda);
    } else if (!make_map(&ma, p.A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.lda, TBM) ||
               (SPLIT && !make_map(&ma2, p.A + p.split, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.lda, TBM))) {
        DETRB_FAIL(DETRB_E_CUDA, "gemm_tc: cuTensorMapEncodeTiled(A) failed (M=%d K=%d lda=%d)", p.M, p.K, p.lda);
    }
    if (!make_map(&mb, p.W, (uint64_t)p.N, w_cols, (uint64_t)p.ldw, BN) ||
        (SPLIT && !make_map(&mb2, p.W + p.wsplit, (uint64_t)p.N, w_cols, (uint64_t)p.ldw, BN)))
        DETRB_FAIL(DETRB_E_CUDA, "gemm_tc: cuTensorMapEncodeTiled(W) failed (N=%d K=%d ldw=%d)", p.N, p.K, p.ldw);
    if (!SPLIT) { ma2 = ma; mb2 = mb; }
    // coalesced TMA epilogue whenever the output is a plain bf16 tile (no fp32 copy, scatter or read-modify-write)
    CUtensorMap mc = ma, mr = ma, mm = ma;
    int tma_epi = (!SPLIT && p.C && !p.Cf && p.out_stride <= 1 && !p.accumulate && g_tma_epilogue) ? 1 : 0;
    if (tma_epi) {
        bool ok = make_map(&mc, p.C, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldc, TBM);
        if (ok && p.residual) ok = make_map(&mr, p.residual, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldr, TBM);
        if (ok && p.mask) ok = make_map(&mm, p.mask, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldm, TBM);
        if (!ok) tma_epi = 0;
    }
    if constexpr (PERSIST) {
        using PL = PLayout<BN, STAGES, OB>;
        static detrb_per_device_flag pconfigured_dev; bool &pconfigured = pconfigured_dev.slot();      // the opt-in is per device
        static int num_sms = 148;
        if (!pconfigured) {
            DETRB_CUDA(cudaFuncSetAttribute((gemm_tcp_kernel<BN, STAGES, IM2COL, OB>), cudaFuncAttributeMaxDynamicSharedMemorySize, PL::TOTAL));
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
            pconfigured = true;
        }
        if (!tma_epi) DETRB_FAIL(DETRB_E_SHAPE, "persistent gemm_tc needs the TMA epilogue (plain bf16 output)");
        const int slot = 16384 * ((p.residual ? 1 : 0) + (p.mask ? 1 : 0));
        // even ring depth: chunk g lives in slot g % rs and is consumed by warpgroup g % 2, so each warpgroup owns its slots and
        // never observes a slot barrier more than one phase behind (mbarrier parity waits cannot tell phases two apart)
        int rs = slot ? (PL::RING_BYTES / slot) & ~1 : 0;
        if (rs > MAXRS) rs = MAXRS;
        if (slot && rs < 2) DETRB_FAIL(DETRB_E_SHAPE, "persistent gemm_tc: no room for the residual / mask ring");
        if constexpr (BN >= 128 && OB == 1) {
            // CTA-pair kernel (cta_group::2): 256 x BN tiles, the straight-line epilogue only
            static int pair_mode = -1;                     // env DETRB_PAIR: 0 off, 1 BN = 256, 2 BN = 128 and 256
            if (pair_mode < 0) { const char *e = getenv("DETRB_PAIR"); pair_mode = e ? atoi(e) : g_tc_pair; }
            const int mode = g_tc_pair_forced >= 0 ? g_tc_pair_forced : pair_mode;
            if (mode >= (BN == 256 ? 1 : 2) && !p.mask && !p.sigmoid && p.N % BN == 0 && p.K % TBK == 0 && !p.a_kb_rows && p.M > TBM) {
                constexpr int QST = BN == 256 ? 5 : 6;
                using QL = PairLayout<BN, QST>;
                static detrb_per_device_flag qconfigured_dev; bool &qconfigured = qconfigured_dev.slot();      // the opt-in is per device
                static int max_pairs = 74;
                if (!qconfigured) {
                    DETRB_CUDA(cudaFuncSetAttribute((gemm_pair_kernel<BN, QST, IM2COL>), cudaFuncAttributeMaxDynamicSharedMemorySize, QL::TOTAL));
                    cudaLaunchConfig_t qc = {};
                    qc.gridDim = dim3(2 * 74); qc.blockDim = dim3(NTHREADS_P); qc.dynamicSmemBytes = QL::TOTAL;
                    cudaLaunchAttribute qa[1];
                    qa[0].id = cudaLaunchAttributeClusterDimension;
                    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                    qc.attrs = qa; qc.numAttrs = 1;
                    int nc = 0;
                    if (cudaOccupancyMaxActiveClusters(&nc, (gemm_pair_kernel<BN, QST, IM2COL>), &qc) == cudaSuccess && nc > 0) max_pairs = nc;
                    else (void)cudaGetLastError();
                    qconfigured = true;
                }
                CUtensorMap mbh;
                if (!make_map(&mbh, p.W, (uint64_t)p.N, w_cols, (uint64_t)p.ldw, BN / 2))
                    DETRB_FAIL(DETRB_E_CUDA, "gemm_tc pair: cuTensorMapEncodeTiled(W half) failed (N=%d K=%d ldw=%d)", p.N, p.K, p.ldw);
                const int qtn = p.N / BN, qtiles = qtn * ceil_div(p.M, 2 * TBM);
                const int pairs = qtiles < max_pairs ? qtiles : max_pairs;
                DETRB_LAUNCH((gemm_pair_kernel<BN, QST, IM2COL>), dim3(2 * pairs), dim3(NTHREADS_P), QL::TOTAL, stream, ma, mbh, mc, mr, p, aux, qtn, qtiles);
                DETRB_CHECK_LAUNCH("gemm_pair_kernel");
                return DETRB_OK;
            }
        }
        const int ntn = ceil_div(p.N, BN), ntiles = ntn * ceil_div(p.M, TBM);
        const int grid_p = ntiles < num_sms ? ntiles : num_sms;
        DETRB_LAUNCH((gemm_tcp_kernel<BN, STAGES, IM2COL, OB>), dim3(grid_p), dim3(NTHREADS_P), PL::TOTAL, stream, ma, mb, mc, mr, mm, p, aux, ntn, ntiles, rs);
        DETRB_CHECK_LAUNCH("gemm_tcp_kernel");
        return DETRB_OK;
    } else {
    using L = SmemLayout<BN, STAGES>;
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute((gemm_tc_kernel<BN, STAGES, IM2COL, SPLIT, BITS_T>), cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, TBM));
    // residual tiles in their own shared memory, fetched at kernel start together with the operands: always for one-stage
    // kernels (the stage cannot hold residual + mask), for grids of at most two CTAs per SM (pure latency chains: occupancy is
    // irrelevant, every microsecond of the chain counts) -- elsewhere the extra 16-32 KB would cost a co-resident CTA
    int early = 0;
    if (tma_epi && p.residual) {
        const long ctas = (long)grid.x * grid.y;
        early = (g_r_early == 2 || (g_r_early == 1 && (STAGES == 1 || ctas <= 2 * 148))) ? 1 : 0;
        if (STAGES == 1 && p.mask) early = 1;
    }
    const size_t smem = early ? L::TOTAL : L::BASE;
    DETRB_LAUNCH((gemm_tc_kernel<BN, STAGES, IM2COL, SPLIT, BITS_T>), dim3(grid), dim3(NTHREADS_TC), smem, stream, ma, mb, mc, mr, mm, ma2, mb2, p, aux, tma_epi | (early << 1));
    DETRB_CHECK_LAUNCH("gemm_tc_kernel");
    return DETRB_OK;
    }
}

// streaming kernel: plain GEMM, bf16 TMA output, weights of one column range resident in shared memory
static thread_local int g_tc_stream = 1;          // 0 off, 1 auto (stream_pick), 2 wherever supported (tests)      (env DETRB_STREAM)
static int g_stream_nst = 0, g_stream_rs = 0;      // overrides (env DETRB_STREAM_NST / DETRB_STREAM_RS), 0 = policy
static int g_stream_diag = 0, g_stream_bn = 0;     // developer switches (env DETRB_STREAM_DIAG / DETRB_STREAM_BN)

// column width per CTA (0: not a streaming shape)
int stream_pick(const detrb_igemm_t &p)
{
    static thread_local bool env_read = false;     // per thread: g_tc_stream is
    if (!env_read) {
        env_read = true;
        if (const char *e = getenv("DETRB_STREAM")) g_tc_stream = atoi(e);
        if (const char *e = getenv("DETRB_STREAM_NST")) g_stream_nst = atoi(e);
        if (const char *e = getenv("DETRB_STREAM_RS")) g_stream_rs = atoi(e);
        if (const char *e = getenv("DETRB_STREAM_DIAG")) g_stream_diag = atoi(e);
        if (const char *e = getenv("DETRB_STREAM_BN")) g_stream_bn = atoi(e);
    }
    if (!g_tc_stream || p.split || p.mask || !p.C || p.Cf || p.out_stride > 1 || p.accumulate || !g_tma_epilogue) return 0;
    if (p.sigmoid || (p.mask_bits && p.out_bits) || (p.drop_p > 0.f && (p.mask_bits || p.out_bits))) return 0;   // bias, residual, ReLU, bit masks or dropout
    if (p.N % 64 != 0 || p.K % TBK != 0 || p.K > 512) return 0;
    int bn = 0;
    if (p.N == 64 || p.N == 128 || p.N == 256) bn = p.N;
    else if (p.N % 256 == 0 && p.N <= 2048) bn = 256;
    if (g_stream_bn && p.N % g_stream_bn == 0 && g_stream_bn < bn) bn = g_stream_bn;        // narrower column ranges (more parts)
    // the weights of one column range must fit 64 KB: K = 256 layers (layer3, N = 1024) take 128-column ranges, eight parts
    // (the encoder's FFN1, N = 2048: sixteen parts)
    if (bn == 256 && ((long)bn * p.K * 2 > 64 * 1024 || p.N > 1024) && p.N % 128 == 0 && p.N / 128 <= 16) bn = 128;
    if (!bn || (long)bn * p.K * 2 > 64 * 1024) return 0;
    if (g_tc_stream >= 2) return bn;
    // auto: the long HBM-bound streams (at least four tiles per CTA); shorter problems stay on the latency-oriented kernels
    const long tiles = (long)ceil_div(p.M, TBM) * (p.N / bn);
    // (K = 256 with eight column ranges -- layer3's 256 -> 1024 convs: the persistent kernel with 128-wide tiles beat this kernel
    //  while its A ring was 5 stages, 43.5 vs 54.7 us; with a whole-tile ring of 4 (launch_stream) this kernel takes 38 us.
    //  env DETRB_K256_TCP=1 sends the shape to the persistent kernel)
    static int k256_tcp = -1;
    if (k256_tcp < 0) { const char *e = getenv("DETRB_K256_TCP"); k256_tcp = e ? atoi(e) : 0; }
    if (k256_tcp && p.K == 256 && p.N / bn >= 8 && (k256_tcp >= 2 || p.N == 1024)) return 0;
    return tiles >= 4 * 148 ? bn : 0;
}

template <int BN, bool HAS_R, bool MBITS, bool OBITS>
int launch_stream(const detrb_igemm_t &p, cudaStream_t stream)
{
    CUtensorMap ma, mb, mc, mr;
    // sliding-window A: rows are overlapping 64-element runs lda elements apart (see launch_tc)
    const uint64_t a_rows = p.a_kb_rows ? (uint64_t)p.M + (uint64_t)(p.K / TBK - 1) * (uint64_t)p.a_kb_rows : (uint64_t)p.M;
    if (!make_map(&ma, p.A, a_rows, p.a_kb_rows ? (uint64_t)TBK : (uint64_t)p.K, (uint64_t)p.lda, TBM) ||
        !make_map(&mb, p.W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.ldw, BN) ||
        !make_map(&mc, p.C, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldc, TBM))
        DETRB_FAIL(DETRB_E_CUDA, "gemm_stream: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d)", p.M, p.N, p.K);
    mr = mc;
    if (p.residual && !make_map(&mr, p.residual, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldr, TBM))
        DETRB_FAIL(DETRB_E_CUDA, "gemm_stream: cuTensorMapEncodeTiled(residual) failed");
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    static int num_sms = 148;
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute((gemm_stream_kernel<BN, HAS_R, MBITS, OBITS>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int nk = p.K / TBK, n_parts = p.N / BN, m_tiles = ceil_div(p.M, TBM);
    // 16 KB units: 224 KB of data = weights + A stages + chunk slots.  A: two tiles in flight (at least 4 stages); the rest are slots
    const int units = (227 * 1024 - 1024 - ST_BAR_BYTES) / 16384 - ceil_div(BN * p.K * 2, 16384);
    int nst = g_stream_nst ? g_stream_nst : (2 * nk > 4 ? 2 * nk : 4);
    if (nst > ST_MAX_NST) nst = ST_MAX_NST;
    if (nst > units - 4) nst = units - 4;
    // a ring that is not a whole number of tiles measured much slower (K = 256, N = 1024: 5 stages 52 us, 4 stages 37 us): round
    // down to a multiple of the k-blocks per tile, the freed stages become chunk slots          (env DETRB_STREAM_WHOLE=0: off)
    static int whole = -1;
    if (whole < 0) { const char *e = getenv("DETRB_STREAM_WHOLE"); whole = e ? atoi(e) : 1; }
    if (whole && !g_stream_nst && nst > nk && nst % nk) nst -= nst % nk;
    int rs = g_stream_rs ? g_stream_rs : units - nst;
    if (rs > units - nst) rs = units - nst;
    rs &= ~1;
    if (rs > ST_MAX_RS) rs = ST_MAX_RS;
    if (nst < nk || nst < 1 || rs < 4) DETRB_FAIL(DETRB_E_SHAPE, "gemm_stream: no room for the rings (K=%d N=%d: %d stages, %d slots)", p.K, p.N, nst, rs);
    int ctas_per_part = num_sms / n_parts;
    if (ctas_per_part > m_tiles) ctas_per_part = m_tiles;
    DETRB_LAUNCH((gemm_stream_kernel<BN, HAS_R, MBITS, OBITS>), dim3(ctas_per_part * n_parts), dim3(NTHREADS_P), 227 * 1024, stream, ma, mb, mc, mr, p, n_parts, m_tiles,
                 nst, rs, g_stream_diag);
    DETRB_CHECK_LAUNCH("gemm_stream_kernel");
    return DETRB_OK;
}

template <int BN>
int dispatch_stream(const detrb_igemm_t &p, cudaStream_t stream)
{
    const bool r = p.residual != nullptr;
    if (p.mask_bits) return r ? launch_stream<BN, true, true, false>(p, stream) : launch_stream<BN, false, true, false>(p, stream);
    if (p.out_bits) return r ? launch_stream<BN, true, false, true>(p, stream) : launch_stream<BN, false, false, true>(p, stream);
    return r ? launch_stream<BN, true, false, false>(p, stream) : launch_stream<BN, false, false, false>(p, stream);
}

}  // namespace

extern "C" int detrb_set_tc_stream(int enable) { stream_pick(detrb_igemm_t{}); int old = g_tc_stream; g_tc_stream = enable; return old; }

bool detrb_make_tiled_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                          uint32_t box_cols, int swizzle_bytes)
{
    return make_map(map, base, rows, cols, ld, box_rows, box_cols, swizzle_bytes);
}

static bool aligned_epilogue(const detrb_igemm_t &p)
{
    if (p.N % 8 != 0 || p.ldw % 8 != 0 || ((uintptr_t)p.W & 15)) return false;
    if (p.mask_bits && (p.mask || p.N % 64 != 0 || p.ldmb % 8 != 0 || p.ldmb * 8 < p.N || ((uintptr_t)p.mask_bits & 7))) return false;
    if (p.out_bits && (p.N % 64 != 0 || p.ldob % 8 != 0 || p.ldob * 8 < p.N || ((uintptr_t)p.out_bits & 7))) return false;
    if (p.C && (p.ldc % 8 != 0 || ((uintptr_t)p.C & 15))) return false;
    if (p.Cf && (p.ldcf % 4 != 0 || ((uintptr_t)p.Cf & 15))) return false;
    if (p.residual && (p.ldr % 8 != 0 || ((uintptr_t)p.residual & 15))) return false;
    if (p.mask && (p.ldm % 8 != 0 || ((uintptr_t)p.mask & 15))) return false;
    if (p.bias && ((uintptr_t)p.bias & 15)) return false;
    return true;
}

// 0: not supported (falls through to igemm.cu), 1: plain GEMM (2-D TMA), 2: implicit-GEMM convolution (TMA im2col)
int detrb_gemm_tc_kind(const detrb_igemm_t &p)
{
    if (p.K % TBK != 0 || p.lda % 8 != 0 || ((uintptr_t)p.A & 15) || !aligned_epilogue(p)) return 0;
    const bool plain = p.KH == 1 && p.KW == 1 && p.stride == 1 && p.pad == 0 && p.Cin == p.K;   // mode 0 and 1 coincide
    if (plain) return get_encode_fn() != nullptr ? 1 : 0;
    if (p.Cin % TBK != 0 || p.K != p.KH * p.KW * p.Cin) return 0;
    if (p.mode == 1 && p.stride != 1) {                     // 3x3 / stride 2 / pad 1: four parity-class sub-convolutions
        if (p.stride == 2 && p.KH == 3 && p.KW == 3 && p.pad == 1 && p.out_stride <= 1 && !p.accumulate && !p.Cf &&
            get_encode_fn() != nullptr && detrb_get_im2col_encode() != nullptr) return 3;
        return 0;
    }
    if (p.KH > 16 || p.KW > 16 || p.stride > 8) return 0;
    if ((long)p.lda * 2 * p.IW >= (1l << 40)) return 0;
    return (get_encode_fn() != nullptr && detrb_get_im2col_encode() != nullptr) ? 2 : 0;
}

bool detrb_gemm_tc_supported(const detrb_igemm_t &p) { return detrb_gemm_tc_kind(p) != 0; }

static thread_local int g_tc_enabled = 1;      // validated on B200 (tests/test_gemm_tc_gpu.py): default on
extern "C" int detrb_set_tc(int enable) { int old = g_tc_enabled; g_tc_enabled = enable; return old; }
bool detrb_gemm_tc_enabled() { return g_tc_enabled != 0; }

extern "C" int detrb_set_tc_persistent(int enable)
{
    int old = g_tc_persistent;
    g_tc_persistent = enable;
    return old;
}
extern "C" int detrb_set_tc_pair(int mode) { int old = g_tc_pair_forced; g_tc_pair_forced = mode; return old; }
extern "C" int detrb_set_tc_tma_epilogue(int enable) { int old = g_tma_epilogue; g_tma_epilogue = enable; return old; }
static thread_local int g_tc_conv_enabled = 1;
extern "C" int detrb_set_tc_conv(int enable) { int old = g_tc_conv_enabled; g_tc_conv_enabled = enable; return old; }
bool detrb_gemm_tc_conv_enabled() { return g_tc_conv_enabled != 0; }

// persistent-kernel policy.  g_tc_persistent: 0 off, 1 auto (where it measured faster), 2 wherever it is supported (tests)
template <bool IM2COL>
static int dispatch_tcp(const detrb_igemm_t &p, int bn, cudaStream_t stream, const ConvClass *cls, bool *taken)
{
    *taken = false;
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char *e = getenv("DETRB_TCP_MIN_TILES")) g_tcp_min_tiles = atol(e);
        if (const char *e = getenv("DETRB_TCP_MIN_TILES256")) g_tcp_min_tiles256 = atol(e);
        if (const char *e = getenv("DETRB_TCP_MIN_NK256")) g_tcp_min_nk256 = atol(e);
        if (const char *e = getenv("DETRB_R_EARLY")) g_r_early = atoi(e);
        if (const char *e = getenv("DETRB_DEEP_SMALL")) g_deep_small = atoi(e);
        if (const char *e = getenv("DETRB_ONE_STAGE")) g_one_stage = atoi(e);
    }
    const bool tma_epi = p.C && !p.Cf && p.out_stride <= 1 && !p.accumulate && g_tma_epilogue;
    // (sliding-window stem: one-tile kernel measured faster, 198 vs 211 us.  Evict-first L2 hints on the read-once residual / mask /
    //  A streams were measured too: no gain, 14.26 vs 14.25 ms/step -- not kept)
    if (!g_tc_persistent || !tma_epi || p.a_kb_rows) return DETRB_OK;
    const int nk = p.K / TBK;
    const bool both = p.residual && p.mask;
    const long mt = ceil_div(p.M, TBM);
    int pick = 0;
    if (g_tc_persistent >= 2) {
        pick = bn ? bn : ((p.N % 256 == 0 && nk > 4) ? 256 : (p.N >= 128 && nk > 2) ? 128 : 64);
        if (pick == 256 && both) pick = 128;        // 3 x 48 KB stages leave room for one residual+mask slot only
    } else if (bn == 0) {
        // measured on B200 (profiles/r01_tcp_sweep.log): long k-loops are compute bound and gain 1.2-2x from the persistent
        // kernel (256-wide tiles once the k-loop amortises the wider epilogue); the HBM-bound 1x1 layers (K <= 256) are faster
        // on the one-tile kernel, whose 4 co-resident CTAs give the epilogue 16 warps per SM
        if (nk >= g_tcp_min_nk256 && p.N % 256 == 0 && !both && mt * (p.N / 256) >= g_tcp_min_tiles256) pick = 256;
        else if (nk > 4 && p.N >= 128 && mt * ceil_div(p.N, 128) >= g_tcp_min_tiles) pick = 128;
        else if (nk == 4 && p.N >= 1024 && p.N % 128 == 0 && !p.mask && mt * (p.N / 128) >= 4 * 148) pick = 128;    // see stream_pick
        else if (nk == 4 && p.N == 64 && !p.residual && !p.mask && mt >= 4 * 148) pick = 64;
    }
    if (!pick) return DETRB_OK;
    *taken = true;
    if (pick == 256)
        return (p.residual || p.mask) ? launch_tc<256, 3, IM2COL, true>(p, stream, cls) : launch_tc<256, 4, IM2COL, true>(p, stream, cls);
    if (pick == 128) return launch_tc<128, 4, IM2COL, true>(p, stream, cls);
    return launch_tc<64, 4, IM2COL, true>(p, stream, cls);
}

template <bool IM2COL>
static int dispatch_tc(const detrb_igemm_t &p, int bn, cudaStream_t stream, const ConvClass *cls = nullptr)
{
    if (p.split) {            // parity precision: one-tile kernel, three k passes, direct-store epilogue
        const bool wide = bn == 128 || (bn == 0 && p.N >= 128);
        return wide ? launch_tc<128, 3, IM2COL, false, 1, true>(p, stream, cls) : launch_tc<64, 3, IM2COL, false, 1, true>(p, stream, cls);
    }
    if constexpr (!IM2COL) {
        if (bn == 0) {
            const int sbn = stream_pick(p);
            if (sbn) return sbn == 256 ? dispatch_stream<256>(p, stream) : sbn == 128 ? dispatch_stream<128>(p, stream) : dispatch_stream<64>(p, stream);
        }
    }
    bool taken = false;
    int rc = dispatch_tcp<IM2COL>(p, bn, stream, cls, &taken);
    if (taken || rc) return rc;
    if (bn == 0) {
        // 64-wide tiles when there are too few 128-wide ones to fill the machine, and for very short k-loops (K <= 128):
        // those tiles are pure latency chains (load -> 4-8 MMAs -> epilogue) and 4 small CTAs per SM overlap better than 3 large
        const long tiles128 = (long)ceil_div(p.N, 128) * ceil_div(p.M, TBM);
        bn = (p.N >= 128 && tiles128 >= 148 && p.K > 128) ? 128 : 64;
    }
    // 1-bit ReLU masks: the kernels that carry the bit code (the full-size backbone layers run on the streaming / halo / persistent
    // kernels; this is the path of small images, strided data gradients and scatter-accumulate shortcuts)
    if (p.mask_bits || p.out_bits)
        return bn == 128 ? launch_tc<128, 3, IM2COL, false, 1, false, true>(p, stream, cls) : launch_tc<64, 3, IM2COL, false, 1, false, true>(p, stream, cls);
    // short k-loops (K <= 256) are latency bound: 2 stages -> 64 / 48 KB of smem -> 3-4 co-resident CTAs per SM hide each other
    const int nk = p.K / TBK;
    const bool shallow = nk <= 4;
    if constexpr (!IM2COL) {
        const long ctas = (long)ceil_div(p.N, bn) * ceil_div(p.M, TBM);
        // K = 64: the whole k-loop is one stage -> 24 KB + the residual tile: five co-resident CTAs instead of four
        if (nk == 1 && g_one_stage && (bn == 64 || bn == 0)) {
            // (128- and 256-wide one-stage tiles measured no faster: 14.80 vs 14.81 ms/step)
            return launch_tc<64, 1, false>(p, stream, cls);
        }
        // at most two CTAs per SM: nothing to overlap with -> fetch every k-block at once (one stage each)
        if (g_deep_small && shallow && nk > 2 && ctas <= 2 * 148)
            return bn == 128 ? launch_tc<128, 4, false>(p, stream, cls) : launch_tc<64, 4, false>(p, stream, cls);
        // long k-loops on less than one CTA per SM (the decoder's K = 2048 linears, 28 CTAs): the loop runs at the speed of the
        // bytes in flight per CTA -> eight stages (192 KB) instead of three
        if (g_deep_small && nk >= 8 && ctas <= 148 && bn == 64) return launch_tc<64, 8, false>(p, stream, cls);
    }
    if (bn == 128) return shallow ? launch_tc<128, 2, IM2COL>(p, stream, cls) : launch_tc<128, 3, IM2COL>(p, stream, cls);
    return shallow ? launch_tc<64, 2, IM2COL>(p, stream, cls) : launch_tc<64, 3, IM2COL>(p, stream, cls);
}

// Data gradient of a 3x3 / stride-2 / pad-1 convolution as four dense stride-1 sub-convolutions over dY, one per output
// parity class (py, px): pixel (2a+py, 2b+px) only sees the taps kh with (py + 1 - kh) even -- {1} for py = 0, {2, 0} for
// py = 1 (dY rows a, a+1) -- so no multiply-by-zero work is issued (the generic transposed gather wastes 75 %).
// can the scatter go through the dense workspace (detrb_igemm_t.scratch) and one coalesced pass?
static bool scratch_scatter_ok(const detrb_igemm_t &p)
{
    return p.scratch && !p.split && p.C && !p.Cf && !p.mask && !p.out_bits && !p.residual && !p.sigmoid && !(p.drop_p > 0.f) && !p.bias &&
           !p.relu && p.mask_scale == 1.f && (((uintptr_t)p.scratch) & 15) == 0 && p.N % 8 == 0 && p.ldc % 8 == 0 && (((uintptr_t)p.C) & 15) == 0;
}

static int strided_dgrad_tc(const detrb_igemm_t &p, cudaStream_t stream)
{
    static const int kmap[2][2] = {{1, -1}, {2, 0}};       // [parity][window offset] -> original tap index along that axis
    // with a workspace: the four classes are written densely (fast TMA-store kernels) and scattered by one coalesced pass that also
    // applies the bit mask; without: each class scatters its rows from the GEMM epilogue (16 bytes per thread and row)
    const bool compact = scratch_scatter_ok(p) && !p.accumulate;
    const bf16 *cls_out[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t soff = 0;
    for (int py = 0; py < 2; py++)
        for (int px = 0; px < 2; px++) {
            const int ny = 1 + py, nx = 1 + px;
            const int A = (p.OH - py + 1) / 2, Bc = (p.OW - px + 1) / 2;       // rows / cols of this class in the dX grid
            if (A <= 0 || Bc <= 0) continue;
            detrb_igemm_t q = p;
            q.mode = 0; q.stride = 1; q.pad = 0; q.KH = ny; q.KW = nx;
            q.OH = A; q.OW = Bc; q.M = p.batch * A * Bc; q.K = ny * nx * p.Cin;
            q.out_stride = 2; q.SH = p.OH; q.SW = p.OW;
            const size_t off = (size_t)py * p.OW + px;                         // first pixel of the class
            if (compact) {
                q.out_stride = 1; q.SH = q.SW = 0; q.mask_bits = nullptr; q.scratch = nullptr;
                q.C = p.scratch + soff; q.ldc = p.N;
                cls_out[py * 2 + px] = reinterpret_cast<const bf16 *>(q.C);
                soff += (size_t)q.M * p.N;
            } else
            if (q.C) q.C = q.C + off * q.ldc;
            if (q.Cf) q.Cf = q.Cf + off * q.ldcf;
            if (q.mask) q.mask = q.mask + off * q.ldm;
            if (q.mask_bits) q.mask_bits = q.mask_bits + off * q.ldmb;
            if (q.out_bits) q.out_bits = q.out_bits + off * q.ldob;
            if (q.residual) q.residual = q.residual + off * q.ldr;
            ConvClass cls;
            cls.aux.pad = 0;
            for (int i = 0; i < 16; i++) cls.aux.wtap[i] = 0;
            for (int ty = 0; ty < ny; ty++)
                for (int tx = 0; tx < nx; tx++) cls.aux.wtap[ty * nx + tx] = kmap[py][ty] * 3 + kmap[px][tx];
            cls.upper_h = A - p.IH; cls.upper_w = Bc - p.IW;
            cls.w_cols = p.K;
            int rc = dispatch_tc<true>(q, 0, stream, &cls);
            if (rc) return rc;
        }
    if (compact)
        return detrb_scatter_s2(cls_out, reinterpret_cast<bf16 *>(p.C), p.ldc, p.mask_bits, p.ldmb, p.batch, p.OH, p.OW, p.N, 0, stream);
    return DETRB_OK;
}

// out_stride == 2 on a plain GEMM (the data gradient of a 1x1 / stride-2 shortcut, scatter-accumulated into the even pixels): with a
// workspace the GEMM writes densely and one coalesced pass adds it to C
static int strided_plain_tc(const detrb_igemm_t &p, cudaStream_t stream)
{
    detrb_igemm_t q = p;
    q.out_stride = 1; q.SH = q.SW = 0; q.accumulate = 0; q.mask_bits = nullptr; q.scratch = nullptr;
    q.C = p.scratch; q.ldc = p.N;
    int rc = dispatch_tc<false>(q, 0, stream);
    if (rc) return rc;
    const bf16 *cls_out[4] = {reinterpret_cast<const bf16 *>(p.scratch), nullptr, nullptr, nullptr};
    return detrb_scatter_s2(cls_out, reinterpret_cast<bf16 *>(p.C), p.ldc, p.mask_bits, p.ldmb, p.batch, p.SH, p.SW, p.N, p.accumulate, stream);
}

int detrb_gemm_tc(const detrb_igemm_t &p, cudaStream_t stream)
{
    const int kind = detrb_gemm_tc_kind(p);
    if (kind == 1) {
        if (p.out_stride == 2 && !p.a_kb_rows && scratch_scatter_ok(p) && p.OH == (p.SH + 1) / 2 && p.OW == (p.SW + 1) / 2 &&
            p.M == p.batch * p.OH * p.OW)
            return strided_plain_tc(p, stream);
        return dispatch_tc<false>(p, 0, stream);
    }
    if (kind == 2) {
        if (detrb_conv_halo_supported(p)) return detrb_conv_halo(p, stream);      // 3x3, 64 -> 64 channels: halo-reusing row kernel (conv_halo.cu)
        return dispatch_tc<true>(p, 0, stream);
    }
    if (kind == 3) return strided_dgrad_tc(p, stream);
    DETRB_FAIL(DETRB_E_SHAPE, "detrb_gemm_tc: unsupported problem");
}

// standalone entry for tests / microbenchmarks: forces the tcgen05 path (error if unsupported); bn = 64 | 128 | 0 (auto)
extern "C" int detrb_gemm_tc_force(const detrb_igemm_t *pp, int bn, detrb_stream_t stream)
{
    if (!pp) DETRB_FAIL(DETRB_E_BADARG, "detrb_gemm_tc_force: null params");
    detrb_igemm_t p = *pp;
    if (p.out_stride < 1) p.out_stride = 1;
    const int kind = detrb_gemm_tc_kind(p);
    if (kind == 0) DETRB_FAIL(DETRB_E_SHAPE, "detrb_gemm_tc_force: problem not supported by the tcgen05 path");
    if (bn != 0 && bn != 64 && bn != 128 && bn != 256) DETRB_FAIL(DETRB_E_BADARG, "detrb_gemm_tc_force: bn must be 0, 64, 128 or 256");
    if (bn == 256 && g_tc_persistent < 2) DETRB_FAIL(DETRB_E_BADARG, "detrb_gemm_tc_force: bn 256 exists only in the persistent kernel");
    if (kind == 3) return strided_dgrad_tc(p, (cudaStream_t)stream);
    return kind == 1 ? dispatch_tc<false>(p, bn, (cudaStream_t)stream) : dispatch_tc<true>(p, bn, (cudaStream_t)stream);
}

#ifdef DETRB_TRACE
extern "C" int detrb_trace_read(unsigned long long *out16, int reset)
{
    DETRB_CUDA(cudaDeviceSynchronize());
    DETRB_CUDA(cudaMemcpyFromSymbol(out16, g_trace, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        DETRB_CUDA(cudaMemcpyToSymbol(g_trace, z, sizeof(z)));
    }
    return DETRB_OK;
}
#endif
