// attention_tc.cu -- Blackwell-native multi-head attention core (head_dim = 32), forward:
//     O = dropout(softmax(scale * Q K^T)) V      per (batch, head);  no S x S tensor in HBM   (transformer.py:308-345)
// TMA (cp.async.bulk.tensor, 64-byte swizzle) stages the Q / K / V head slices ([rows x 32] bf16 = 64-byte rows) into shared
// memory, one elected thread issues tcgen05.mma for both products, the scores and the P V partial products live in TENSOR MEMORY:
//
//   warp 0      TMA producer: Q tile once, then a ring of KVST (K_j, V_j) tiles of 64 keys
//   warp 1      TMEM allocator + MMA issuer:  S_j = Q K_j^T  (M = 128 queries, N = 64 keys, K = 32: two UMMAs, both operands K-major)
//                                             PV_j = P_j V_j (M = 128, N = 32, K = 64: four UMMAs; P_j K-major from shared memory,
//                                                             V_j MN-major -- the [key][channel] tile exactly as TMA lands it)
//               S and PV are double-buffered in TMEM (2 x 64 + 2 x 32 columns), S_{j+1} is issued before PV_j: the tensor pipe runs
//               ahead of the softmax
//   warps 2-5   softmax: thread = query row (TMEM lane).  tcgen05.ld brings the 64 scores of the row into registers; running
//               maximum / sum in the log2 domain (one FFMA + one ex2.approx per score), dropout by AND-masks on the packed bf16
//               probabilities (same counter-based stream as the mma.sync kernels of attention.cu), P_j -> shared memory in the
//               128-byte-swizzled K-major layout the UMMA reads; the running output row (32 fp32 registers) is rescaled and
//               PV_{j-1} added from TMEM one tile late, so the P V product never stalls the softmax
// No row reductions across threads (no shuffles), no ldmatrix, no register-fragment MMAs: per score the SM issues ~10 instructions
// instead of ~25, which is what bounds attention at head_dim 32 (128 tensor FLOP per score).  Two CTAs are co-resident per SM
// (256 TMEM columns, 75 KB shared memory, <= 168 registers).
#include "tc_common.cuh"

// waits of the single-thread producer / MMA warps: relaxed polling (DETRB_ATTN_SPIN builds keep the plain spin for A/B runs)
#ifdef DETRB_ATTN_SPIN
#define MBAR_WAIT_CTRL(bar, parity) mbar_wait(bar, parity)
#else
#define MBAR_WAIT_CTRL(bar, parity) mbar_wait_relaxed(bar, parity, 32u)
#endif

namespace {

constexpr int DH = 32;
constexpr int BQ = 128;                       // queries per CTA  = UMMA M
constexpr int BKV = 64;                       // keys per tile    = UMMA N of S, K of P V
constexpr int KVST = 4;                       // (K, V) pipeline stages
constexpr int Q_BYTES = BQ * DH * 2;          // 8 KB
constexpr int KT_BYTES = BKV * DH * 2;        // 4 KB (K tile; the V tile is the same size)
constexpr int P_BYTES = BQ * BKV * 2;         // 16 KB
constexpr int OFF_Q = 0;
constexpr int OFF_KV = OFF_Q + Q_BYTES;
constexpr int OFF_P = OFF_KV + KVST * 2 * KT_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int SMEM_FWD = OFF_BAR + 256 + 1024;             // + slack for the 1024-byte alignment of the dynamic smem base
constexpr int NTHREADS = 192;
constexpr int TMEM_COLS = 256;                             // S: 2 x 64 columns at 0 / 64, PV: 2 x 32 columns at 128 / 160
constexpr float LOG2E = 1.4426950408889634f;

// [rows x 32] bf16 tile, 64-byte rows, 64-byte swizzle (8-row atoms of 512 B): K-major operand (rows = M / N index, the 32 channels = K)
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// the same tile read as an MN-major operand (rows = K index, the 32 channels = N): 8-row (K) groups 512 B apart (SBO); one 32-wide MN block
__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// [rows x 64] bf16 tile, 128-byte rows, 128-byte swizzle, K-major (the P tile)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16, D = f32, A = B = bf16; N >> 3 at [17,23), M >> 4 at [24,29); bit 16: B is MN-major
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(NTHREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const detrb_attn_fwd_t p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem0 + OFF_BAR;
    const uint32_t q_full = bar0;
    auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar0 + 8u * (1 + KVST + s); };
    auto s_full = [&](int b) { return bar0 + 8u * (1 + 2 * KVST + b); };
    auto s_empty = [&](int b) { return bar0 + 8u * (3 + 2 * KVST + b); };
    auto p_full = [&](int b) { return bar0 + 8u * (5 + 2 * KVST + b); };
    auto p_empty = [&](int b) { return bar0 + 8u * (7 + 2 * KVST + b); };
    auto pv_full = [&](int b) { return bar0 + 8u * (9 + 2 * KVST + b); };
    auto pv_empty = [&](int b) { return bar0 + 8u * (11 + 2 * KVST + b); };
    const uint32_t tmem_slot = bar0 + 8u * (13 + 2 * KVST);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int nkt = (p.Lk + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < KVST; s++) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        for (int i = 0; i < 2; i++) {
            mbar_init(s_full(i), 1); mbar_init(s_empty(i), 4);
            mbar_init(p_full(i), 4); mbar_init(p_empty(i), 1);
            mbar_init(pv_full(i), 1); mbar_init(pv_empty(i), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, Q_BYTES);
            tma_load_2d(smem0 + OFF_Q, &map_q, q_full, h * DH, b * p.Lq + q0);
            for (int j = 0; j < nkt; j++) {
                const int st = j % KVST;
                MBAR_WAIT_CTRL(kv_empty(st), ((j / KVST) & 1) ^ 1);
                mbar_expect_tx(kv_full(st), 2 * KT_BYTES);
                const uint32_t dst = smem0 + OFF_KV + (uint32_t)st * (2 * KT_BYTES);
                // rows beyond this batch's Lk keys belong to the next batch (or are zero-filled past the end of the tensor):
                // their scores are masked to -inf below, so what they hold never matters (they are finite either way)
                tma_load_2d(dst, &map_k, kv_full(st), h * DH, b * p.Lk + j * BKV);
                tma_load_2d(dst + KT_BYTES, &map_v, kv_full(st), h * DH, b * p.Lk + j * BKV);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t IDESC_S = idesc_f16(BQ, BKV, 0), IDESC_PV = idesc_f16(BQ, DH, 1);
            MBAR_WAIT_CTRL(q_full, 0);
            const uint64_t dq = desc_k_sw64(smem0 + OFF_Q);
            auto issue_s = [&](int j) {
                const int st = j % KVST, sb = j & 1;
                MBAR_WAIT_CTRL(kv_full(st), (j / KVST) & 1);
                MBAR_WAIT_CTRL(s_empty(sb), ((j >> 1) & 1) ^ 1);              // the softmax warps have read S_{j-2} out of this buffer
                tc_fence_after();
                const uint64_t dk = desc_k_sw64(smem0 + OFF_KV + (uint32_t)st * (2 * KT_BYTES));
#pragma unroll
                for (int k = 0; k < DH / 16; k++)                        // 16 channels = 32 B further inside the 64-byte row
                    tc_mma_f16(tmem_base + (uint32_t)(sb * BKV), dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), IDESC_S, k != 0);
                tc_commit(s_full(sb));
            };
            issue_s(0);
            for (int j = 0; j < nkt; j++) {
                if (j + 1 < nkt) issue_s(j + 1);                         // scores of the next tile while the softmax works on this one
                const int st = j % KVST, pb = j & 1;
                MBAR_WAIT_CTRL(p_full(pb), (j >> 1) & 1);                     // P_j is in shared memory
                MBAR_WAIT_CTRL(pv_empty(pb), ((j >> 1) & 1) ^ 1);             // PV_{j-2} has been read out of this buffer
                tc_fence_after();
                const uint64_t dp = desc_k_sw128(smem0 + OFF_P + (uint32_t)pb * P_BYTES);
                const uint64_t dv = desc_mn_sw64(smem0 + OFF_KV + (uint32_t)st * (2 * KT_BYTES) + KT_BYTES);
#pragma unroll
                for (int k = 0; k < BKV / 16; k++)                       // 16 keys: 32 B along a P row, 2 swizzle atoms (1024 B) down the V tile
                    tc_mma_f16(tmem_base + (uint32_t)(2 * BKV + pb * DH), dp + (uint64_t)(k * 2), dv + (uint64_t)(k * (1024 >> 4)), IDESC_PV, k != 0);
                tc_commit(pv_full(pb));
                tc_commit(kv_empty(st));
                tc_commit(p_empty(pb));
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax: thread = query row =====================
        const int qr = warp & 3;                                         // TMEM lane quarter of this warp
        const int row = qr * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qr * 32) << 16);
        const bool drop = p.drop_p > 0.f;
        const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
        const float sl2 = p.scale * LOG2E;
        const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
        const uint32_t rh = attn_drop_rowhash(seed, p.site, (uint32_t)((b * p.H + h) * p.Lq + q0 + row));
        const uint32_t prow = (uint32_t)row * 128u, psw = (uint32_t)(row & 7);
        float o[DH];
#pragma unroll
        for (int i = 0; i < DH; i++) o[i] = 0.f;
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < nkt; j++) {
            const int sb = j & 1;
            mbar_wait(s_full(sb), (j >> 1) & 1);
            tc_fence_after();
            uint32_t s[BKV];
            tc_ld32(lane_addr + (uint32_t)(sb * BKV), *reinterpret_cast<uint32_t (*)[32]>(&s[0]));
            tc_ld32(lane_addr + (uint32_t)(sb * BKV + 32), *reinterpret_cast<uint32_t (*)[32]>(&s[32]));
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(sb));                     // the scores are in registers: S_{j+2} may overwrite the buffer
            const int kbase = j * BKV;
            if (kbase + BKV > p.Lk) {                                    // keys beyond Lk exist only in the last tile
#pragma unroll
                for (int c = 0; c < BKV; c++)
                    if (kbase + c >= p.Lk) s[c] = 0xff800000u;           // -inf
            }
            float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains (the softmax warps are latency bound)
#pragma unroll
            for (int c = 0; c < BKV; c += 4) {
                tm[0] = fmaxf(tm[0], __uint_as_float(s[c])); tm[1] = fmaxf(tm[1], __uint_as_float(s[c + 1]));
                tm[2] = fmaxf(tm[2], __uint_as_float(s[c + 2])); tm[3] = fmaxf(tm[3], __uint_as_float(s[c + 3]));
            }
            const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
            const float mnew = fmaxf(m, tmax * sl2);                     // log2 domain (sl2 > 0: max commutes with the scaling)
            const float alpha = ex2(m - mnew);                           // ex2(-inf) = 0 on the first tile
            const float mneg = -mnew;
            float rs[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[BKV / 2];                                        // packed bf16 pairs (keys 2i, 2i+1)
#pragma unroll
            for (int c = 0; c < BKV; c += 4) {
                const float p0 = ex2(fmaf(__uint_as_float(s[c]), sl2, mneg));
                const float p1 = ex2(fmaf(__uint_as_float(s[c + 1]), sl2, mneg));
                const float p2 = ex2(fmaf(__uint_as_float(s[c + 2]), sl2, mneg));
                const float p3 = ex2(fmaf(__uint_as_float(s[c + 3]), sl2, mneg));
                rs[0] += p0; rs[1] += p1; rs[2] += p2; rs[3] += p3;
                pk[c >> 1] = pack_bf16x2(p0, p1);
                pk[(c >> 1) + 1] = pack_bf16x2(p2, p3);
            }
            l = l * alpha + ((rs[0] + rs[1]) + (rs[2] + rs[3]));
            if (drop) {                                                  // the 1/(1-p) rescale is folded into the final normalisation
#pragma unroll
                for (int g16 = 0; g16 < BKV / 16; g16++) {
                    // one word per key pair (k, k+8) of a 16-key group: low field = key 16g + i, high field = key 16g + 8 + i
                    uint32_t kw[8];
                    const uint32_t pair0 = (uint32_t)((kbase >> 4) + g16) * 8u;
#pragma unroll
                    for (int i = 0; i < 8; i++) kw[i] = attn_drop_keepbits(attn_drop_word(rh, pair0 + i), thresh2);
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        pk[g16 * 8 + (i >> 1)] &= attn_drop_mask2_lo(kw[i], kw[i + 1]);          // keys 16g + i, 16g + i + 1
                        pk[g16 * 8 + 4 + (i >> 1)] &= attn_drop_mask2_hi(kw[i], kw[i + 1]);      // keys 16g + 8 + i, 16g + 9 + i
                    }
                }
            }
            // P_j -> shared memory (K-major, 128-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ (r % 8))
            mbar_wait(p_empty(sb), ((j >> 1) & 1) ^ 1);                  // PV_{j-2} has consumed this buffer
            const uint32_t pbuf = smem0 + OFF_P + (uint32_t)sb * P_BYTES + prow;
#pragma unroll
            for (int c8 = 0; c8 < 8; c8++)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(pbuf + (((uint32_t)c8 ^ psw) << 4)),
                             "r"(pk[4 * c8]), "r"(pk[4 * c8 + 1]), "r"(pk[4 * c8 + 2]), "r"(pk[4 * c8 + 3]) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy writes -> visible to the UMMA
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(sb));
            // the previous tile's P V product joins the running output one tile late: o = (o + PV_{j-1}) * alpha_j
            if (j > 0) {
                const int pb = (j - 1) & 1;
                mbar_wait(pv_full(pb), ((j - 1) >> 1) & 1);
                tc_fence_after();
                uint32_t r[DH];
                tc_ld32(lane_addr + (uint32_t)(2 * BKV + pb * DH), r);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pv_empty(pb));
#pragma unroll
                for (int i = 0; i < DH; i++) o[i] = (o[i] + __uint_as_float(r[i])) * alpha;
            }
            m = mnew;
        }
        {
            const int pb = (nkt - 1) & 1;
            mbar_wait(pv_full(pb), ((nkt - 1) >> 1) & 1);
            tc_fence_after();
            uint32_t r[DH];
            tc_ld32(lane_addr + (uint32_t)(2 * BKV + pb * DH), r);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < DH; i++) o[i] += __uint_as_float(r[i]);
        }
        const int q = q0 + row;
        if (q < p.Lq) {
            const float inv = (drop ? 1.f / (1.f - p.drop_p) : 1.f) / l;
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(p.O) + ((size_t)b * p.Lq + q) * p.ldo + h * DH);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 v;
                v.x = pack_bf16x2(o[8 * i] * inv, o[8 * i + 1] * inv); v.y = pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
                v.z = pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv); v.w = pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
                dst[i] = v;
            }
            if (p.lse) p.lse[((size_t)b * p.H + h) * p.Lq + q] = (m + log2f(l)) * (1.f / LOG2E);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}


// ================================================================================================ backward
// Two kernels with the forward kernel's anatomy (TMA producer warp, MMA issuer warp, four thread-per-row warps) and only the
// operand layouts the forward kernel uses: K-major 64-byte-swizzled [rows x 32] tiles from TMA, K-major 128-byte-swizzled
// [128 x 64] tiles written by the row threads, MN-major reads of the [row][channel] tiles.  delta = rowsum(dO * O) comes from
// attn_delta_kernel (attention.cu); the 1/(1-p) dropout rescale and the score scale are folded into dss / the final dV.
//
//   attn_bwd_dq_tc_kernel   CTA = 128 queries (thread = query row: lse, delta, the dropout row hash are per-thread scalars), loop
//                           over 64-key tiles:  S = Q K_j^T, dP = dO V_j^T (TMEM), dS = P * (dP - delta) -> shared memory,
//                           dQ += dS K_j accumulated in TMEM over all key tiles
//   attn_bwd_dkv_tc_kernel  CTA = 128 keys (thread = key row), loop over 64-query tiles:  S^T = K Q_j^T, dP^T = V dO_j^T (TMEM),
//                           P^T and dS^T -> shared memory, dV += P^T dO_j and dK += dS^T Q_j accumulated in TMEM; the per-query
//                           values (lse, delta, row hash) of a tile are staged in shared memory and read as broadcasts
constexpr int BSTG = 4;                                   // pipeline stages of the streamed operand pair
constexpr int TILE128 = BQ * DH * 2;                      // [128 x 32] bf16 tile: 8 KB
constexpr int TILE64 = BKV * DH * 2;                      // [64 x 32] bf16 tile: 4 KB
constexpr int DS_BYTES = BQ * BKV * 2;                    // [128 x 64] bf16 tile: 16 KB
// dQ kernel: Q, dO (resident) | BSTG x (K_j, V_j) | 2 x dS | barriers
constexpr int DQ_OFF_Q = 0, DQ_OFF_DO = TILE128, DQ_OFF_KV = 2 * TILE128, DQ_OFF_DS = DQ_OFF_KV + BSTG * 2 * TILE64;
constexpr int DQ_OFF_BAR = DQ_OFF_DS + 2 * DS_BYTES;
constexpr int SMEM_DQ = DQ_OFF_BAR + 256 + 1024;
// dK/dV kernel: K, V (resident) | BSTG x (Q_j, dO_j) | P^T, dS^T | 2 x per-query arrays (lse, delta, row hash) | barriers
constexpr int DKV_OFF_K = 0, DKV_OFF_V = TILE128, DKV_OFF_QDO = 2 * TILE128, DKV_OFF_P = DKV_OFF_QDO + BSTG * 2 * TILE64;
constexpr int DKV_OFF_DS = DKV_OFF_P + DS_BYTES, DKV_OFF_ARR = DKV_OFF_DS + DS_BYTES, DKV_OFF_BAR = DKV_OFF_ARR + 2 * 3 * BKV * 4;
constexpr int SMEM_DKV = DKV_OFF_BAR + 256 + 1024;

// one swizzled 128-byte row (64 bf16 as 32 packed words) of a K-major [128 x 64] tile: chunk c of row r lives at chunk c ^ (r % 8)
__device__ __forceinline__ void st_row_half(uint32_t row_addr, uint32_t sw, int half, const uint32_t (&pk)[16]) {
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++)
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(row_addr + ((((uint32_t)(half * 4 + c4)) ^ sw) << 4)),
                     "r"(pk[4 * c4]), "r"(pk[4 * c4 + 1]), "r"(pk[4 * c4 + 2]), "r"(pk[4 * c4 + 3]) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const detrb_attn_bwd_t p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem0 + DQ_OFF_BAR;
    const uint32_t q_full = bar0;
    auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar0 + 8u * (1 + BSTG + s); };
    const uint32_t sdp_full = bar0 + 8u * (1 + 2 * BSTG), sdp_empty = bar0 + 8u * (2 + 2 * BSTG), dq_full = bar0 + 8u * (3 + 2 * BSTG);
    auto ds_full = [&](int b) { return bar0 + 8u * (4 + 2 * BSTG + b); };
    auto ds_empty = [&](int b) { return bar0 + 8u * (6 + 2 * BSTG + b); };
    const uint32_t tmem_slot = bar0 + 8u * (8 + 2 * BSTG);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DQ = 128;             // TMEM columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int nkt = (p.Lk + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < BSTG; s++) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        mbar_init(sdp_full, 1); mbar_init(sdp_empty, 4); mbar_init(dq_full, 1);
        for (int i = 0; i < 2; i++) { mbar_init(ds_full(i), 4); mbar_init(ds_empty(i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
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
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * TILE128);
            tma_load_2d(smem0 + DQ_OFF_Q, &map_q, q_full, h * DH, b * p.Lq + q0);
            tma_load_2d(smem0 + DQ_OFF_DO, &map_do, q_full, h * DH, b * p.Lq + q0);
            for (int j = 0; j < nkt; j++) {
                const int st = j % BSTG;
                MBAR_WAIT_CTRL(kv_empty(st), ((j / BSTG) & 1) ^ 1);
                mbar_expect_tx(kv_full(st), 2 * TILE64);
                const uint32_t dst = smem0 + DQ_OFF_KV + (uint32_t)st * (2 * TILE64);
                tma_load_2d(dst, &map_k, kv_full(st), h * DH, b * p.Lk + j * BKV);
                tma_load_2d(dst + TILE64, &map_v, kv_full(st), h * DH, b * p.Lk + j * BKV);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC_S = idesc_f16(BQ, BKV, 0), IDESC_DQ = idesc_f16(BQ, DH, 1);
            MBAR_WAIT_CTRL(q_full, 0);
            const uint64_t dq_ = desc_k_sw64(smem0 + DQ_OFF_Q), ddo = desc_k_sw64(smem0 + DQ_OFF_DO);
            auto issue_sdp = [&](int j) {                                // S = Q K_j^T and dP = dO V_j^T (single-buffered in TMEM)
                const int st = j % BSTG;
                MBAR_WAIT_CTRL(kv_full(st), (j / BSTG) & 1);
                MBAR_WAIT_CTRL(sdp_empty, (j & 1) ^ 1);                       // the row threads hold tile j-1 in registers
                tc_fence_after();
                const uint32_t kv = smem0 + DQ_OFF_KV + (uint32_t)st * (2 * TILE64);
                const uint64_t dk = desc_k_sw64(kv), dv = desc_k_sw64(kv + TILE64);
#pragma unroll
                for (int k = 0; k < DH / 16; k++) tc_mma_f16(tmem_base + COL_S, dq_ + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), IDESC_S, k != 0);
#pragma unroll
                for (int k = 0; k < DH / 16; k++) tc_mma_f16(tmem_base + COL_DP, ddo + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), IDESC_S, k != 0);
                tc_commit(sdp_full);
            };
            issue_sdp(0);
            for (int j = 0; j < nkt; j++) {
                if (j + 1 < nkt) issue_sdp(j + 1);
                const int st = j % BSTG, db = j & 1;
                MBAR_WAIT_CTRL(ds_full(db), (j >> 1) & 1);
                tc_fence_after();
                const uint64_t dds = desc_k_sw128(smem0 + DQ_OFF_DS + (uint32_t)db * DS_BYTES);
                const uint64_t dkm = desc_mn_sw64(smem0 + DQ_OFF_KV + (uint32_t)st * (2 * TILE64));
#pragma unroll
                for (int k = 0; k < BKV / 16; k++)                       // dQ += dS_j K_j (K_j read MN-major: [key][channel])
                    tc_mma_f16(tmem_base + COL_DQ, dds + (uint64_t)(k * 2), dkm + (uint64_t)(k * (1024 >> 4)), IDESC_DQ, (j | k) != 0);
                tc_commit(kv_empty(st));
                tc_commit(ds_empty(db));
            }
            tc_commit(dq_full);
        }
        __syncwarp();
    } else {
        const int qr = warp & 3, row = qr * 32 + lane, q = q0 + row;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qr * 32) << 16);
        const bool drop = p.drop_p > 0.f, qok = q < p.Lq;
        const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
        const float drop_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
        const float sl2 = p.scale * LOG2E, dss = p.scale * drop_scale;
        const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
        const size_t rowg = ((size_t)b * p.H + h) * p.Lq + q;
        const uint32_t rh = attn_drop_rowhash(seed, p.site, (uint32_t)rowg);
        const float lse2 = qok ? -p.lse[rowg] * LOG2E : -INFINITY;       // negated, log2 domain (rows beyond Lq: p = 0)
        const float dlt = qok ? p.delta[rowg] * p.scale : 0.f;           // scale * delta
        const uint32_t prow = (uint32_t)row * 128u, psw = (uint32_t)(row & 7);
        for (int j = 0; j < nkt; j++) {
            const int db = j & 1, kbase = j * BKV;
            mbar_wait(sdp_full, j & 1);
            tc_fence_after();
            uint32_t pk[2][16];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t s[32], dp[32];
                tc_ld32(lane_addr + COL_S + (uint32_t)(half * 32), s);
                tc_ld32(lane_addr + COL_DP + (uint32_t)(half * 32), dp);
                tc_wait_ld();
                if (half == 1) {                                         // both halves are in registers: the next tile may be computed
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sdp_empty);
                }
#pragma unroll
                for (int g16 = 0; g16 < 2; g16++) {
                    uint32_t kw[8];
                    if (drop) {
                        const uint32_t pair0 = (uint32_t)((kbase >> 4) + half * 2 + g16) * 8u;
#pragma unroll
                        for (int i = 0; i < 8; i++) kw[i] = attn_drop_keepbits(attn_drop_word(rh, pair0 + i), thresh2);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float ds2[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int c = g16 * 16 + i + e;              // column within this half; key = kbase + half*32 + c
                            const float pv = ex2(fmaf(__uint_as_float(s[c]), sl2, lse2));
                            uint32_t dpu = dp[c];
                            if (drop) dpu &= ((i + e) & 8) ? attn_drop_mask_hi(kw[(i + e) & 7]) : attn_drop_mask_lo(kw[(i + e) & 7]);
                            float dsv = pv * fmaf(__uint_as_float(dpu), dss, -dlt);      // scale * dS
                            // keys beyond Lk (last tile): their K rows are not this problem's; exp(0 - lse) may even overflow
                            if (kbase + half * 32 + c >= p.Lk) dsv = 0.f;
                            ds2[e] = dsv;
                        }
                        pk[half][(g16 * 16 + i) >> 1] = pack_bf16x2(ds2[0], ds2[1]);
                    }
                }
            }
            mbar_wait(ds_empty(db), ((j >> 1) & 1) ^ 1);
            const uint32_t dst = smem0 + DQ_OFF_DS + (uint32_t)db * DS_BYTES + prow;
            st_row_half(dst, psw, 0, pk[0]);
            st_row_half(dst, psw, 1, pk[1]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_full(db));
        }
        mbar_wait(dq_full, 0);
        tc_fence_after();
        uint32_t r[DH];
        tc_ld32(lane_addr + COL_DQ, r);
        tc_wait_ld();
        if (qok) {
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(p.dQ) + ((size_t)b * p.Lq + q) * p.lddq + h * DH);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1]));
                v.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3]));
                v.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5]));
                v.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7]));
                dst[i] = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

__global__ void __launch_bounds__(NTHREADS, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                       const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const detrb_attn_bwd_t p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem0 + DKV_OFF_BAR;
    const uint32_t kv_full = bar0;
    auto q_full = [&](int s) { return bar0 + 8u * (1 + s); };
    auto q_empty = [&](int s) { return bar0 + 8u * (1 + BSTG + s); };
    const uint32_t sdp_full = bar0 + 8u * (1 + 2 * BSTG), sdp_empty = bar0 + 8u * (2 + 2 * BSTG), dkv_full = bar0 + 8u * (3 + 2 * BSTG);
    const uint32_t pds_full = bar0 + 8u * (4 + 2 * BSTG), pds_empty = bar0 + 8u * (5 + 2 * BSTG);
    const uint32_t tmem_slot = bar0 + 8u * (6 + 2 * BSTG);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DV = 128, COL_DK = 160;
    float *arr = reinterpret_cast<float *>(smem_raw + (smem0 + DKV_OFF_ARR - smem_u32(smem_raw)));   // [2][3][64]: lse2 | dlt | row hash

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int nqt = (p.Lq + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
        mbar_init(kv_full, 1);
        for (int s = 0; s < BSTG; s++) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
        mbar_init(sdp_full, 1); mbar_init(sdp_empty, 4); mbar_init(dkv_full, 1);
        mbar_init(pds_full, 4); mbar_init(pds_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
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
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * TILE128);
            tma_load_2d(smem0 + DKV_OFF_K, &map_k, kv_full, h * DH, b * p.Lk + k0);
            tma_load_2d(smem0 + DKV_OFF_V, &map_v, kv_full, h * DH, b * p.Lk + k0);
            for (int j = 0; j < nqt; j++) {
                const int st = j % BSTG;
                MBAR_WAIT_CTRL(q_empty(st), ((j / BSTG) & 1) ^ 1);
                mbar_expect_tx(q_full(st), 2 * TILE64);
                const uint32_t dst = smem0 + DKV_OFF_QDO + (uint32_t)st * (2 * TILE64);
                tma_load_2d(dst, &map_q, q_full(st), h * DH, b * p.Lq + j * BKV);
                tma_load_2d(dst + TILE64, &map_do, q_full(st), h * DH, b * p.Lq + j * BKV);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC_S = idesc_f16(BQ, BKV, 0), IDESC_D = idesc_f16(BQ, DH, 1);
            MBAR_WAIT_CTRL(kv_full, 0);
            const uint64_t dk = desc_k_sw64(smem0 + DKV_OFF_K), dv = desc_k_sw64(smem0 + DKV_OFF_V);
            auto issue_sdp = [&](int j) {                                // S^T = K Q_j^T and dP^T = V dO_j^T
                const int st = j % BSTG;
                MBAR_WAIT_CTRL(q_full(st), (j / BSTG) & 1);
                MBAR_WAIT_CTRL(sdp_empty, (j & 1) ^ 1);
                tc_fence_after();
                const uint32_t qd = smem0 + DKV_OFF_QDO + (uint32_t)st * (2 * TILE64);
                const uint64_t dq_ = desc_k_sw64(qd), ddo = desc_k_sw64(qd + TILE64);
#pragma unroll
                for (int k = 0; k < DH / 16; k++) tc_mma_f16(tmem_base + COL_S, dk + (uint64_t)(k * 2), dq_ + (uint64_t)(k * 2), IDESC_S, k != 0);
#pragma unroll
                for (int k = 0; k < DH / 16; k++) tc_mma_f16(tmem_base + COL_DP, dv + (uint64_t)(k * 2), ddo + (uint64_t)(k * 2), IDESC_S, k != 0);
                tc_commit(sdp_full);
            };
            issue_sdp(0);
            for (int j = 0; j < nqt; j++) {
                if (j + 1 < nqt) issue_sdp(j + 1);
                const int st = j % BSTG;
                MBAR_WAIT_CTRL(pds_full, j & 1);
                tc_fence_after();
                const uint32_t qd = smem0 + DKV_OFF_QDO + (uint32_t)st * (2 * TILE64);
                const uint64_t dp_ = desc_k_sw128(smem0 + DKV_OFF_P), dds = desc_k_sw128(smem0 + DKV_OFF_DS);
                const uint64_t dqm = desc_mn_sw64(qd), ddom = desc_mn_sw64(qd + TILE64);
#pragma unroll
                for (int k = 0; k < BKV / 16; k++)                       // dV += P^T dO_j
                    tc_mma_f16(tmem_base + COL_DV, dp_ + (uint64_t)(k * 2), ddom + (uint64_t)(k * (1024 >> 4)), IDESC_D, (j | k) != 0);
#pragma unroll
                for (int k = 0; k < BKV / 16; k++)                       // dK += dS^T Q_j
                    tc_mma_f16(tmem_base + COL_DK, dds + (uint64_t)(k * 2), dqm + (uint64_t)(k * (1024 >> 4)), IDESC_D, (j | k) != 0);
                tc_commit(q_empty(st));
                tc_commit(pds_empty);
            }
            tc_commit(dkv_full);
        }
        __syncwarp();
    } else {
        const int qr = warp & 3, row = qr * 32 + lane, key = k0 + row;
        const int te = threadIdx.x - 64;                                 // 0..127 among the row threads
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qr * 32) << 16);
        const bool drop = p.drop_p > 0.f;
        const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
        const float drop_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
        const float sl2 = p.scale * LOG2E, dss = p.scale * drop_scale;
        const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
        // this thread's key is one field of the dropout word (key pair (k, k+8) of a 16-key group): fixed pair index, fixed field
        const uint32_t kpair = attn_drop_pair((uint32_t)key);
        const uint32_t fsel = ((key >> 3) & 1) ? 0xBBBBu : 0x9999u;      // PRMT selector: sign of byte 3 (high field) / byte 1 (low field)
        const uint32_t prow = (uint32_t)row * 128u, psw = (uint32_t)(row & 7);
        const size_t rowbase = ((size_t)b * p.H + h) * p.Lq;
        for (int j = 0; j < nqt; j++) {
            // per-query values of this tile -> shared memory (double-buffered; one named barrier per tile among the 128 row threads)
            float *a_lse = arr + (j & 1) * 3 * BKV, *a_dlt = a_lse + BKV;
            uint32_t *a_rh = reinterpret_cast<uint32_t *>(a_dlt + BKV);
            if (te < BKV) {
                const int q = j * BKV + te;
                const bool ok = q < p.Lq;
                a_lse[te] = ok ? -p.lse[rowbase + q] * LOG2E : -INFINITY;   // queries beyond Lq: p = 0
                a_dlt[te] = ok ? p.delta[rowbase + q] * p.scale : 0.f;
                a_rh[te] = attn_drop_rowhash(seed, p.site, (uint32_t)(rowbase + q));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(sdp_full, j & 1);
            tc_fence_after();
            uint32_t ppk[2][16], dpk[2][16];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t s[32], dp[32];
                tc_ld32(lane_addr + COL_S + (uint32_t)(half * 32), s);
                tc_ld32(lane_addr + COL_DP + (uint32_t)(half * 32), dp);
                tc_wait_ld();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sdp_empty);
                }
#pragma unroll
                for (int c4 = 0; c4 < 32; c4 += 4) {
                    const float4 l4 = *reinterpret_cast<const float4 *>(a_lse + half * 32 + c4);
                    const float4 d4 = *reinterpret_cast<const float4 *>(a_dlt + half * 32 + c4);
                    const uint4 h4 = *reinterpret_cast<const uint4 *>(a_rh + half * 32 + c4);
                    const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq4[4] = {d4.x, d4.y, d4.z, d4.w};
                    const uint32_t hq[4] = {h4.x, h4.y, h4.z, h4.w};
                    float pd[4], ds[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float pv = ex2(fmaf(__uint_as_float(s[c4 + e]), sl2, lq[e]));
                        uint32_t mk = 0xffffffffu;
                        if (drop) { const uint32_t kb = attn_drop_keepbits(attn_drop_word(hq[e], kpair), thresh2); mk = prmt(kb, kb, fsel); }
                        pd[e] = __uint_as_float(__float_as_uint(pv) & mk);                       // dropped probability (for dV), unscaled
                        ds[e] = pv * fmaf(__uint_as_float(dp[c4 + e] & mk), dss, -dq4[e]);     // scale * dS^T
                    }
                    ppk[half][c4 >> 1] = pack_bf16x2(pd[0], pd[1]); ppk[half][(c4 >> 1) + 1] = pack_bf16x2(pd[2], pd[3]);
                    dpk[half][c4 >> 1] = pack_bf16x2(ds[0], ds[1]); dpk[half][(c4 >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
                }
            }
            mbar_wait(pds_empty, (j & 1) ^ 1);                           // the dV / dK products of tile j-1 have read the tiles
            st_row_half(smem0 + DKV_OFF_P + prow, psw, 0, ppk[0]);
            st_row_half(smem0 + DKV_OFF_P + prow, psw, 1, ppk[1]);
            st_row_half(smem0 + DKV_OFF_DS + prow, psw, 0, dpk[0]);
            st_row_half(smem0 + DKV_OFF_DS + prow, psw, 1, dpk[1]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full);
        }
        mbar_wait(dkv_full, 0);
        tc_fence_after();
        uint32_t rv[DH], rk[DH];
        tc_ld32(lane_addr + COL_DV, rv);
        tc_ld32(lane_addr + COL_DK, rk);
        tc_wait_ld();
        if (key < p.Lk) {
            uint4 *dstv = reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(p.dV) + ((size_t)b * p.Lk + key) * p.lddv + h * DH);
            uint4 *dstk = reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(p.dK) + ((size_t)b * p.Lk + key) * p.lddk + h * DH);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 v, k;
                v.x = pack_bf16x2(__uint_as_float(rv[8 * i]) * drop_scale, __uint_as_float(rv[8 * i + 1]) * drop_scale);
                v.y = pack_bf16x2(__uint_as_float(rv[8 * i + 2]) * drop_scale, __uint_as_float(rv[8 * i + 3]) * drop_scale);
                v.z = pack_bf16x2(__uint_as_float(rv[8 * i + 4]) * drop_scale, __uint_as_float(rv[8 * i + 5]) * drop_scale);
                v.w = pack_bf16x2(__uint_as_float(rv[8 * i + 6]) * drop_scale, __uint_as_float(rv[8 * i + 7]) * drop_scale);
                k.x = pack_bf16x2(__uint_as_float(rk[8 * i]), __uint_as_float(rk[8 * i + 1]));
                k.y = pack_bf16x2(__uint_as_float(rk[8 * i + 2]), __uint_as_float(rk[8 * i + 3]));
                k.z = pack_bf16x2(__uint_as_float(rk[8 * i + 4]), __uint_as_float(rk[8 * i + 5]));
                k.w = pack_bf16x2(__uint_as_float(rk[8 * i + 6]), __uint_as_float(rk[8 * i + 7]));
                dstv[i] = v;
                dstk[i] = k;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

}  // namespace

static thread_local int g_attn_tc = 1;
extern "C" int detrb_set_tc_attn(int enable) { int old = g_attn_tc; g_attn_tc = enable; return old; }
bool detrb_attn_tc_enabled() { return g_attn_tc != 0; }

bool detrb_attn_fwd_tc_supported(const detrb_attn_fwd_t &p)
{
    if (p.split) return false;
    if (p.ldq % 8 || p.ldk % 8 || p.ldv % 8 || p.ldo % 8) return false;
    if (((uintptr_t)p.Q | (uintptr_t)p.K | (uintptr_t)p.V | (uintptr_t)p.O) & 15) return false;
    if ((long long)p.B * p.Lq >= (1ll << 31) || (long long)p.B * p.Lk >= (1ll << 31)) return false;
    return true;
}

int detrb_attn_fwd_tc(const detrb_attn_fwd_t &p, cudaStream_t stream)
{
    CUtensorMap mq, mk, mv;
    const uint64_t cols = (uint64_t)p.H * DH;
    if (!detrb_make_tiled_map(&mq, p.Q, (uint64_t)p.B * p.Lq, cols, (uint64_t)p.ldq, BQ, DH, 64) ||
        !detrb_make_tiled_map(&mk, p.K, (uint64_t)p.B * p.Lk, cols, (uint64_t)p.ldk, BKV, DH, 64) ||
        !detrb_make_tiled_map(&mv, p.V, (uint64_t)p.B * p.Lk, cols, (uint64_t)p.ldv, BKV, DH, 64))
        DETRB_FAIL(DETRB_E_CUDA, "attn_fwd_tc: cuTensorMapEncodeTiled failed (B=%d H=%d Lq=%d Lk=%d)", p.B, p.H, p.Lq, p.Lk);
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
        configured = true;
    }
    dim3 grid(ceil_div(p.Lq, BQ), p.H, p.B);
    DETRB_LAUNCH(attn_fwd_tc_kernel, dim3(grid), dim3(NTHREADS), SMEM_FWD, stream, mq, mk, mv, p);
    DETRB_CHECK_LAUNCH("attn_fwd_tc_kernel");
    return DETRB_OK;
}

bool detrb_attn_bwd_tc_supported(const detrb_attn_bwd_t &p)
{
    if (p.split) return false;
    if (p.ldq % 8 || p.ldk % 8 || p.ldv % 8 || p.lddo % 8 || p.lddq % 8 || p.lddk % 8 || p.lddv % 8) return false;
    if (((uintptr_t)p.Q | (uintptr_t)p.K | (uintptr_t)p.V | (uintptr_t)p.dO | (uintptr_t)p.dQ | (uintptr_t)p.dK | (uintptr_t)p.dV) & 15) return false;
    if ((long long)p.B * p.Lq >= (1ll << 31) || (long long)p.B * p.Lk >= (1ll << 31)) return false;
    return true;
}

// dK/dV then dQ (delta is computed by the caller: attn_delta_kernel)
int detrb_attn_bwd_tc(const detrb_attn_bwd_t &p, cudaStream_t stream)
{
    const uint64_t cols = (uint64_t)p.H * DH;
    const uint64_t rq = (uint64_t)p.B * p.Lq, rk = (uint64_t)p.B * p.Lk;
    CUtensorMap q128, do128, k64, v64, q64, do64, k128, v128;
    if (!detrb_make_tiled_map(&q128, p.Q, rq, cols, (uint64_t)p.ldq, BQ, DH, 64) || !detrb_make_tiled_map(&do128, p.dO, rq, cols, (uint64_t)p.lddo, BQ, DH, 64) ||
        !detrb_make_tiled_map(&k64, p.K, rk, cols, (uint64_t)p.ldk, BKV, DH, 64) || !detrb_make_tiled_map(&v64, p.V, rk, cols, (uint64_t)p.ldv, BKV, DH, 64) ||
        !detrb_make_tiled_map(&q64, p.Q, rq, cols, (uint64_t)p.ldq, BKV, DH, 64) || !detrb_make_tiled_map(&do64, p.dO, rq, cols, (uint64_t)p.lddo, BKV, DH, 64) ||
        !detrb_make_tiled_map(&k128, p.K, rk, cols, (uint64_t)p.ldk, BQ, DH, 64) || !detrb_make_tiled_map(&v128, p.V, rk, cols, (uint64_t)p.ldv, BQ, DH, 64))
        DETRB_FAIL(DETRB_E_CUDA, "attn_bwd_tc: cuTensorMapEncodeTiled failed (B=%d H=%d Lq=%d Lk=%d)", p.B, p.H, p.Lq, p.Lk);
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ));
        DETRB_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV));
        configured = true;
    }
    const int parts = p.parts ? p.parts : 7;
    if (parts & 2) {
        DETRB_LAUNCH(attn_bwd_dkv_tc_kernel, dim3(ceil_div(p.Lk, BQ), p.H, p.B), dim3(NTHREADS), SMEM_DKV, stream, q64, do64, k128, v128, p);
        DETRB_CHECK_LAUNCH("attn_bwd_dkv_tc_kernel");
    }
    if (parts & 4) {
        DETRB_LAUNCH(attn_bwd_dq_tc_kernel, dim3(ceil_div(p.Lq, BQ), p.H, p.B), dim3(NTHREADS), SMEM_DQ, stream, q128, do128, k64, v64, p);
        DETRB_CHECK_LAUNCH("attn_bwd_dq_tc_kernel");
    }
    return DETRB_OK;
}
