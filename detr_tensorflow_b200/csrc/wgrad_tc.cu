// wgrad_tc.cu -- weight gradients on tcgen05:   dW[n, k] += rowscale[n] * sum_m dY[m, n] * gather(A)[m, k]
//
// The reduction runs over pixels m, i.e. over the *rows* of both operands as they sit in memory ([pixel][channel]).  TMA
// lands [64 pixels x 64 channels] boxes (128-byte swizzle) and the UMMA descriptors declare both operands MN-major
// (a_major = b_major = 1): no transposition anywhere.  The gathered operand uses the same TMA im2col tensor map as the forward
// convolution (one box per 64-channel slice of one filter tap), so 1x1, 3x3, strided convolutions and Linear layers share the
// kernel.  D = 128 (out channels) x 128 (k = tap*Cin + c) fp32 in TMEM; CTAs split the pixel range and merge with fp32 atomics.
//
// Replaces wgrad.cu (mma.sync) for every layer whose Cin is a multiple of 64.  The 7x7 stem arrives as a plain problem too: its
// space-to-depth image is a sliding-window operand (detrb_wgrad_t.a_kb_rows: 64-column block j of pixel m is the 64-element
// run j * a_kb_rows pixels further down, rows overlap), K = 4 window rows x 4 taps x 16 channels = 256, and k_mask drops the
// columns that do not exist in the 7x7x3 kernel.
#include "tc_common.cuh"

namespace {

constexpr int WN = 128;                 // dW rows per CTA (output channels)  = UMMA M
constexpr int WK = 128;                 // dW cols per CTA (tap*Cin + c)      = UMMA N
constexpr int WP = 64;                  // pixels per pipeline stage (4 UMMAs of K = 16)
constexpr int BOX_BYTES = WP * 128;     // one [64 pixels x 64 channels] bf16 box
constexpr int WTHREADS = 192;
// Tile variants <NA, KT>: NA accumulators of 128 output channels x KT k-columns per CTA.  <1,128> is the base tile (32 KB per
// 64-pixel stage, two CTAs per SM).  On the deep layers the base tile is bound by the L2 -> shared-memory path (64 FLOP per
// staged byte: ~15 TB/s for ~830 TFLOP/s, tests/time_wgrad_shapes.py); the wider tiles stage fewer bytes per FLOP -- <1,256> and
// <2,128> 85 FLOP/B (48 KB stages), <2,256> 128 FLOP/B (64 KB stages, all 512 TMEM columns) -- on one CTA per SM.
__host__ __device__ constexpr int wgrad_stage_bytes(int na, int kt) { return (2 * na + kt / 64) * BOX_BYTES; }
__host__ __device__ constexpr int wgrad_stages(int na, int kt) { return na * kt == 128 ? 3 : na * kt == 256 ? 4 : 3; }
__host__ __device__ constexpr int wgrad_smem(int na, int kt) { return wgrad_stages(na, kt) * wgrad_stage_bytes(na, kt) + 256 + 1024; }

// MN-major 128B-swizzled operand: 64-element (128 B) rows, one per K index (pixel); 8-pixel swizzle atoms 1024 B apart
// (SBO); the next 64-channel block of the MN dimension lives one box further (LBO = BOX_BYTES).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(BOX_BYTES >> 4) << 16;                   // leading byte offset: next 64-channel block
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset: next 8 pixels
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t wgrad_idesc(int kt) { return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kt >> 3) << 17) | ((uint32_t)(WN >> 4) << 24); }

// SPLIT (parity precision, include/detrb.h): dY and A are bf16 pairs (lo planes through map_y2 / map_x2); the pixel range is
// walked three times -- dY_hi^T A_hi, dY_lo^T A_hi, dY_hi^T A_lo -- into the same TMEM accumulator.
template <bool IM2COL, bool SPLIT, int NA, int KT>
__global__ void __launch_bounds__(WTHREADS)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_y2, const __grid_constant__ CUtensorMap map_x2, const detrb_wgrad_t p,
                const int pix_per_split, const int stem_mask, const int interleave)
{
    constexpr int NBY = 2 * NA, NBX = KT / 64;                          // dY / gathered boxes per stage
    constexpr int STG = wgrad_stages(NA, KT);
    constexpr int STG_BYTES = wgrad_stage_bytes(NA, KT);
    constexpr int TCOLS = NA * KT;                                      // TMEM columns: accumulator a at column a * KT
    constexpr uint32_t IDESC = wgrad_idesc(KT);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STG * STG_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STG + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STG);
    const uint32_t tmem_slot = bar_base + 8u * (2 * STG + 1);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * KT, n0 = blockIdx.y * (WN * NA);
    // pixel blocks of this split: a contiguous range, or (interleave) every gridDim.z-th 64-pixel block -- then all CTAs together sweep
    // the pixel range front to back like the data-gradient kernel that runs beside this one on the main stream and reads the same
    // dY: whichever of the two comes second finds it in L2 (the reduction over pixels does not care about the order)
    const int m_begin = interleave ? (int)blockIdx.z * WP : (int)blockIdx.z * pix_per_split;
    const int m_step = interleave ? (int)gridDim.z * WP : WP;
    const int m_end = interleave ? p.M : min(p.M, m_begin + pix_per_split);
    const int nsteps1 = (m_end - m_begin + m_step - 1) / m_step;    // >= 1 by construction of the grid
    const int nsteps = SPLIT ? 3 * nsteps1 : nsteps1;

    // bias gradient (column sums of dY) fused: the k-tile-0 CTAs' otherwise idle epilogue warps add up the dY boxes of every stage
    const bool do_bias = p.dbias != nullptr && blockIdx.x == 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_y);
        tma_prefetch_desc(&map_x);
        if (SPLIT) { tma_prefetch_desc(&map_y2); tma_prefetch_desc(&map_x2); }
        for (int s = 0; s < STG; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), do_bias ? 5 : 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)TCOLS) : "memory");
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
            // the 64-wide k blocks of this tile: (tap, channel offset); blocks past K (a partial last tile) gather tap 0 -- their
            // columns are dropped in the epilogue
            int tap[NBX], cc[NBX];
#pragma unroll
            for (int j = 0; j < NBX; j++) {
                const int k = k0 + j * 64;
                tap[j] = k < p.K ? k / p.Cin : 0; cc[j] = k < p.K ? k - tap[j] * p.Cin : 0;
            }
            const int ohw = p.OH * p.OW;
            int stage = 0; uint32_t phase = 0;
            for (int st = 0; st < nsteps; st++) {
                const int part = SPLIT ? st / nsteps1 : 0;
                const CUtensorMap *py = (SPLIT && part == 1) ? &map_y2 : &map_y;
                const CUtensorMap *px = (SPLIT && part == 2) ? &map_x2 : &map_x;
                const int m = m_begin + (st - part * nsteps1) * m_step;
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), STG_BYTES);
                const uint32_t dst = smem_base + stage * STG_BYTES;
#pragma unroll
                for (int j = 0; j < NBY; j++)
                    tma_load_2d(dst + j * BOX_BYTES, py, full_bar(stage), n0 + 64 * j, m);       // dY[m.., n0 + 64 j ..): zero fill past N
                if (IM2COL) {
                    const int img = m / ohw, rem = m - img * ohw, oy = rem / p.OW, ox = rem - oy * p.OW;
                    const int w0 = ox * p.stride - p.pad, h0 = oy * p.stride - p.pad;
#pragma unroll
                    for (int j = 0; j < NBX; j++) {
                        const int kh = tap[j] / p.KW, kw = tap[j] - kh * p.KW;
                        tma_load_im2col(dst + (NBY + j) * BOX_BYTES, px, full_bar(stage), cc[j], w0, h0, img, (uint16_t)kw, (uint16_t)kh);
                    }
                } else {
                    // sliding-window A (a_kb_rows > 0): 64-column block j is the 64-element run j * a_kb_rows rows further down
                    const int kb = k0 / 64, sl = p.a_kb_rows;
#pragma unroll
                    for (int j = 0; j < NBX; j++)
                        tma_load_2d(dst + (NBY + j) * BOX_BYTES, px, full_bar(stage), sl ? 0 : k0 + 64 * j, m + (kb + j) * sl);
                }
                if (++stage == STG) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int st = 0; st < nsteps; st++) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t base = smem_base + stage * STG_BYTES;
                const uint64_t db = make_desc_mn(base + NBY * BOX_BYTES);
#pragma unroll
                for (int ks = 0; ks < WP / 16; ks++)      // 16 pixels = 2 swizzle atoms = 2048 B further down the box
#pragma unroll
                    for (int a = 0; a < NA; a++)          // accumulator a: output channels n0 + 128 a .. (dY boxes 2a, 2a + 1)
                        tc_mma_f16(tmem_base + (uint32_t)(a * KT), make_desc_mn(base + a * 2 * BOX_BYTES) + (uint64_t)(ks * (2048 >> 4)),
                                   db + (uint64_t)(ks * (2048 >> 4)), IDESC, (st | ks) != 0);
                tc_commit(empty_bar(stage));
                if (++stage == STG) { stage = 0; phase ^= 1; }
            }
            tc_commit(tmem_full_bar);
        }
        __syncwarp();
    } else {
        if (do_bias) {
            // thread -> channel pair 2cp, 2cp+1 of the CTA's 128 and one half of the stage's 64 pixel rows.  dY box layout: pixel
            // row r at r*128 B, its 16-byte chunk c (8 channels) at chunk c ^ (r % 8) (128-byte swizzle)
            const int te = threadIdx.x - 64, cp = te & 63, half = te >> 6;
            const uint32_t box_off = (uint32_t)(cp >> 5) * BOX_BYTES, chunk = (uint32_t)((cp & 31) >> 2), within = (uint32_t)(cp & 3) * 4u;
            float s0[NA], s1[NA];
#pragma unroll
            for (int a = 0; a < NA; a++) { s0[a] = 0.f; s1[a] = 0.f; }
            int stage = 0; uint32_t phase = 0;
            for (int st = 0; st < nsteps; st++) {
                mbar_wait(full_bar(stage), phase);
                const uint32_t base = smem_base + stage * STG_BYTES + box_off + within;
                const bool count = !SPLIT || st < 2 * nsteps1;            // the third pass stages dY_hi again
#pragma unroll
                for (int a = 0; a < NA; a++) {
#pragma unroll 8
                    for (int r = half * 32; count && r < half * 32 + 32; r++) {
                        uint32_t u;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(base + (uint32_t)(a * 2 * BOX_BYTES) + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
                        const float2 f = unpack_bf16x2(u);
                        s0[a] += f.x; s1[a] += f.y;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar(stage));            // this warp is done reading the stage
                if (++stage == STG) { stage = 0; phase ^= 1; }
            }
#pragma unroll
            for (int a = 0; a < NA; a++) {
                const int nb = n0 + a * WN + 2 * cp;
                if (nb < p.N) atomicAdd(p.dbias + nb, s0[a] * (p.rowscale ? p.rowscale[nb] : 1.f));
                if (nb + 1 < p.N) atomicAdd(p.dbias + nb + 1, s1[a] * (p.rowscale ? p.rowscale[nb + 1] : 1.f));
            }
        }
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const bool vec_ok = (p.ldw % 4 == 0) && (((uintptr_t)p.dW & 15) == 0);
#pragma unroll 1
        for (int ac = 0; ac < NA * KT; ac += 16) {
            const int a = ac / KT, c0 = ac - a * KT;
            const int n = n0 + a * WN + q * 32 + lane;
            const bool row_ok = n < p.N;
            const float sc = (row_ok && p.rowscale) ? p.rowscale[n] : 1.f;
            float *drow = p.dW + (size_t)(row_ok ? n : 0) * p.ldw;
            uint32_t r[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ac, r);
            tc_wait_ld();
            if (!row_ok) continue;
#pragma unroll
            for (int i4 = 0; i4 < 16; i4 += 4) {
                // 4 consecutive gradient elements of one row: one vector reduction (red.global.add.v4.f32) instead of 4 atomics --
                // the L2 atomic units, not HBM, bound this kernel on the small layer1/2 weight matrices
                float v[4]; bool ok[4]; bool all_ok = true;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = k0 + c0 + i4 + i;
                    ok[i] = k < p.K;
                    if (stem_mask) {
                        // space-to-depth stem: column k = (ta, tb, (ry*2+rx)*3 + c) stands for the 7x7 tap (2ta+ry-1, 2tb+rx-1);
                        // taps outside 0..6 and the 4 padding channels do not exist in the reference kernel: no gradient
                        const int ch = k & 15, tb = (k >> 4) & 3, ta = k >> 6;
                        const int ry = ch / 6, rx = (ch / 3) & 1, kh = 2 * ta + ry - 1, kw = 2 * tb + rx - 1;
                        ok[i] = ok[i] && ch < 12 && kh >= 0 && kh <= 6 && kw >= 0 && kw <= 6;
                    }
                    v[i] = __uint_as_float(r[i4 + i]) * sc;
                    all_ok = all_ok && ok[i];
                }
                float *dst = drow + k0 + c0 + i4;
                if (all_ok && vec_ok) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) if (ok[i]) atomicAdd(dst + i, v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TCOLS) : "memory");
    }
}

// ================================================================================================ narrow variant (N <= 64)
// Layers with at most 64 output channels (the stem, layer1's 256 -> 64 and 3x3 64 -> 64 convs: the largest pixel counts of the network)
// waste half of the 128-row UMMA of the kernel above (dY is the M operand there), and its k-tiles each re-read dY.  Here the roles
// are swapped and one CTA owns ALL of K:   D_a[128 k-columns x 64 out channels] += X_a^T[128 x 16 px] . dY[16 px x 64]   for the
// NA = ceil(K / 128) accumulators a (64 TMEM columns each).  Per 64-pixel stage: ONE dY box and K / 64 gathered boxes (nine taps of the
// 3x3: 80 KB), every MMA does useful work (M = 128, N = 64), dY is read once.  CTAs split the pixel range, fp32 red.add merge; in
// the epilogue a warp's lanes hold consecutive k of one output channel: coalesced reductions.
constexpr int NARROW_MAX_BOXES = 9;
constexpr uint32_t WIDESC_NARROW = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

template <bool IM2COL>
__global__ void __launch_bounds__(WTHREADS)
wgrad_narrow_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x, const detrb_wgrad_t p,
                    const int pix_per_split, const int stem_mask, const int nstages, const int tmem_cols)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nb = p.K / 64, na = (nb + 1) / 2;                        // gathered boxes per stage, accumulators
    const uint32_t stage_bytes = (uint32_t)(1 + nb) * BOX_BYTES;
    const uint32_t bar_base = smem_base + (uint32_t)nstages * stage_bytes + BOX_BYTES;      // (+ one box: the odd last accumulator's upper half reads past the stage)
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * 8;
    const uint32_t tmem_slot = bar_base + 8u * 9;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_begin = (int)blockIdx.x * pix_per_split;
    const int m_end = min(p.M, m_begin + pix_per_split);
    const int nsteps = (m_end - m_begin + WP - 1) / WP;                // >= 1 by construction of the grid
    const bool do_bias = p.dbias != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_y);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < nstages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), do_bias ? 5 : 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)tmem_cols) : "memory");
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
            const int ohw = p.OH * p.OW;
            int stage = 0; uint32_t phase = 0;
            for (int st = 0; st < nsteps; st++) {
                const int m = m_begin + st * WP;
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), stage_bytes);
                const uint32_t dst = smem_base + (uint32_t)stage * stage_bytes;
                tma_load_2d(dst, &map_y, full_bar(stage), 0, m);                              // dY[m .. m+64, 0 .. 64)
                if (IM2COL) {
                    const int img = m / ohw, rem = m - img * ohw, oy = rem / p.OW, ox = rem - oy * p.OW;
                    const int w0 = ox * p.stride - p.pad, h0 = oy * p.stride - p.pad;
                    for (int j = 0; j < nb; j++) {
                        const int k = j * 64, tap = k / p.Cin, cc = k - tap * p.Cin;
                        const int kh = tap / p.KW, kw = tap - kh * p.KW;
                        tma_load_im2col(dst + (uint32_t)(1 + j) * BOX_BYTES, &map_x, full_bar(stage), cc, w0, h0, img, (uint16_t)kw, (uint16_t)kh);
                    }
                } else {
                    const int sl = p.a_kb_rows;                                              // sliding-window A: block j is j * sl rows further down
                    for (int j = 0; j < nb; j++)
                        tma_load_2d(dst + (uint32_t)(1 + j) * BOX_BYTES, &map_x, full_bar(stage), sl ? 0 : j * 64, m + j * sl);
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int st = 0; st < nsteps; st++) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t base = smem_base + (uint32_t)stage * stage_bytes;
                const uint64_t dy = make_desc_mn(base);
#pragma unroll
                for (int ks = 0; ks < WP / 16; ks++) {        // 16 pixels = 2 swizzle atoms = 2048 B further down every box
                    const uint64_t off = (uint64_t)(ks * (2048 >> 4));
                    for (int a = 0; a < na; a++)
                        tc_mma_f16(tmem_base + (uint32_t)(a * 64), make_desc_mn(base + (uint32_t)(1 + 2 * a) * BOX_BYTES) + off, dy + off,
                                   WIDESC_NARROW, (st | ks) != 0);
                }
                tc_commit(empty_bar(stage));
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            tc_commit(tmem_full_bar);
        }
        __syncwarp();
    } else {
        const int te = threadIdx.x - 64;                                 // 0 .. 127
        if (do_bias) {
            // thread -> channel pair (2cp, 2cp+1) and one quarter of the stage's 64 pixel rows; dY box: pixel row r at r * 128 B, its
            // 16-byte chunk c (8 channels) at chunk c ^ (r % 8)
            const int cp = te & 31, quarter = te >> 5;
            const uint32_t chunk = (uint32_t)(cp >> 2), within = (uint32_t)(cp & 3) * 4u;
            float s0 = 0.f, s1 = 0.f;
            int stage = 0; uint32_t phase = 0;
            for (int st = 0; st < nsteps; st++) {
                mbar_wait(full_bar(stage), phase);
                const uint32_t base = smem_base + (uint32_t)stage * stage_bytes + within;
#pragma unroll 8
                for (int r = quarter * 16; r < quarter * 16 + 16; r++) {
                    uint32_t u;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(base + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
                    const float2 f = unpack_bf16x2(u);
                    s0 += f.x; s1 += f.y;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar(stage));
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            const int c0 = 2 * cp;
            if (c0 < p.N) atomicAdd(p.dbias + c0, s0 * (p.rowscale ? p.rowscale[c0] : 1.f));
            if (c0 + 1 < p.N) atomicAdd(p.dbias + c0 + 1, s1 * (p.rowscale ? p.rowscale[c0 + 1] : 1.f));
        }
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        for (int a = 0; a < na; a++) {
            const int k = a * 128 + q * 32 + lane;                   // this thread's TMEM lane = one k column of dW
            bool k_ok = k < p.K;
            if (stem_mask) {
                const int ch = k & 15, tb = (k >> 4) & 3, ta = k >> 6;
                const int ry = ch / 6, rx = (ch / 3) & 1, kh = 2 * ta + ry - 1, kw = 2 * tb + rx - 1;
                k_ok = k_ok && ch < 12 && kh >= 0 && kh <= 6 && kw >= 0 && kw <= 6;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t r[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c0), r);
                tc_wait_ld();
                if (!k_ok) continue;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int n = c0 + i;
                    if (n < p.N) atomicAdd(p.dW + (size_t)n * p.ldw + k, __uint_as_float(r[i]) * (p.rowscale ? p.rowscale[n] : 1.f));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

}  // namespace


bool detrb_wgrad_tc_supported(const detrb_wgrad_t &p)
{
    if (p.Cin % 64 != 0 || p.K != p.KH * p.KW * p.Cin) return false;
    if (p.lda % 8 != 0 || p.ldy % 8 != 0 || ((uintptr_t)p.A & 15) || ((uintptr_t)p.dY & 15)) return false;
    if (p.KH > 16 || p.KW > 16 || p.stride > 8) return false;
    return detrb_get_im2col_encode() != nullptr;
}

// small token counts (transformer: M = 800 / 8400) are latency bound and faster on the mma.sync split kernel
bool detrb_wgrad_tc_profitable(const detrb_wgrad_t &p)
{
    // large pixel counts (backbone) and the encoder-sized linears (M = 8400); the decoder's (M = 800) stay on the mma.sync split kernel
    static long min_m = -1, min_nk = -1;                 // env overrides for tuning runs
    if (min_m < 0) {
        const char *e1 = getenv("DETRB_WGRAD_TC_MIN_M"), *e2 = getenv("DETRB_WGRAD_TC_MIN_NK");
        min_m = e1 ? atol(e1) : 512;
        min_nk = e2 ? atol(e2) : (1l << 16);
    }
    return p.M >= 16384 || (p.M >= min_m && (long)p.N * p.K >= min_nk);
}

static int g_wgrad_tile = 0;            // developer switch, see detrb_wgrad_tc (detrb_set_wgrad_tile / env DETRB_WGRAD_TILE)
extern "C" int detrb_set_wgrad_tile(int mode) { int old = g_wgrad_tile; g_wgrad_tile = mode; return old; }

// which tile for which layer (tests/time_wgrad_shapes.py on B200, isolated launches)
static void wgrad_tile_policy(const detrb_wgrad_t &p, int *na, int *kt)
{
    // profiles/r02_wgrad_tiles.log: launched alone the 128 x 256 tile is 5-15 % faster on the backbone's deep layers (layer2 3x3,
    // layer3 / layer4 1x1, layer3 3x3), the 256-channel tiles (NA = 2) are slower everywhere.  Inside the train step, where the
    // weight gradients share the GPU with the data-gradient chain of the main stream, the wide tile changes nothing (11.46 / 11.45 ms
    // base tile vs 11.48 / 11.48 ms, profiles/r02_wgrad_tile_step_ab.log): the base tile (two CTAs per SM) stays the policy, the
    // wide tiles stay selectable (detrb_set_wgrad_tile, tests/test_gemm_tc_gpu.py::test_wgrad_tc_wide_tiles)
    (void)p;
    *na = 1; *kt = WK;
}

int detrb_wgrad_tc(const detrb_wgrad_t &p, cudaStream_t stream)
{
    const bool plain = p.KH == 1 && p.KW == 1 && p.stride == 1 && p.pad == 0;
    CUtensorMap my, mx, my2, mx2;
    const bool sp = p.split != 0;
    if (!detrb_make_tiled_map(&my, p.dY, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldy, WP, 64) ||
        (sp && !detrb_make_tiled_map(&my2, p.dY + p.split, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldy, WP, 64)))
        DETRB_FAIL(DETRB_E_CUDA, "wgrad_tc: tensor map for dY failed (M=%d N=%d ldy=%d)", p.M, p.N, p.ldy);
    if (plain && p.a_kb_rows > 0) {
        const uint64_t rows = (uint64_t)p.M + (uint64_t)(p.K / 64 - 1) * (uint64_t)p.a_kb_rows;
        if (p.K % WK != 0 || !detrb_make_tiled_map(&mx, p.A, rows, 64, (uint64_t)p.lda, WP, 64) ||
            (sp && !detrb_make_tiled_map(&mx2, p.A + p.split, rows, 64, (uint64_t)p.lda, WP, 64)))
            DETRB_FAIL(DETRB_E_CUDA, "wgrad_tc: tensor map for the sliding-window A failed (rows=%llu K=%d lda=%d)", (unsigned long long)rows, p.K, p.lda);
    } else if (plain) {
        if (!detrb_make_tiled_map(&mx, p.A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.lda, WP, 64) ||
            (sp && !detrb_make_tiled_map(&mx2, p.A + p.split, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.lda, WP, 64)))
            DETRB_FAIL(DETRB_E_CUDA, "wgrad_tc: tensor map for A failed (M=%d K=%d lda=%d)", p.M, p.K, p.lda);
    } else {
        const int lower = -p.pad;
        const int upper_w = (p.OW - 1) * p.stride + 1 + lower - p.IW, upper_h = (p.OH - 1) * p.stride + 1 + lower - p.IH;
        if (upper_w > 0 || upper_h > 0 || upper_w < -16 || upper_h < -16)
            DETRB_FAIL(DETRB_E_SHAPE, "wgrad_tc: inconsistent conv geometry");
        int rc = detrb_make_im2col_map(&mx, p.A, p.batch, p.IH, p.IW, p.Cin, p.lda, lower, lower, upper_w, upper_h, p.stride, WP, 1, 64);
        if (rc) return rc;
        if (sp) {
            rc = detrb_make_im2col_map(&mx2, p.A + p.split, p.batch, p.IH, p.IW, p.Cin, p.lda, lower, lower, upper_w, upper_h, p.stride, WP, 1, 64);
            if (rc) return rc;
        }
    }
    if (!sp) { my2 = my; mx2 = mx; }
    static int narrow_on = -1;                       // env DETRB_WGRAD_NARROW=0: the general kernel everywhere
    if (narrow_on < 0) { const char *e = getenv("DETRB_WGRAD_NARROW"); narrow_on = e ? atoi(e) : 1; }
    // mode 1: the plain operands only (K = 256: layer1's 256 -> 64 convs, the stem); mode 2: also the gathered 3x3 (measured SLOWER
    // than the general kernel, 117 vs 100 us: nine im2col boxes per stage leave room for two stages only and the loop runs at TMA
    // latency -- the fix is the halo scheme of conv_halo.cu, whole rows staged once, taps as shifted views; not built)
    if (narrow_on && (plain ? p.K >= 128 : narrow_on >= 2) && !sp && p.N <= 64 && p.K % 64 == 0 && p.K / 64 <= NARROW_MAX_BOXES && p.M >= 148 * WP) {
        const int nb = p.K / 64, na = (nb + 1) / 2;
        const int stage_bytes = (1 + nb) * BOX_BYTES;
        int nstages = (227 * 1024 - 2048 - 1024 - BOX_BYTES) / stage_bytes;
        if (nstages > 4) nstages = 4;
        const int tmem_cols = na * 64 <= 64 ? 64 : na * 64 <= 128 ? 128 : na * 64 <= 256 ? 256 : 512;
        const size_t smem = (size_t)nstages * stage_bytes + BOX_BYTES + 256 + 1024;
        static detrb_per_device_flag nconfigured_dev; bool &nconfigured = nconfigured_dev.slot();      // the opt-in is per device
        if (!nconfigured) {
            DETRB_CUDA(cudaFuncSetAttribute((wgrad_narrow_kernel<false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            DETRB_CUDA(cudaFuncSetAttribute((wgrad_narrow_kernel<true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            nconfigured = true;
        }
        int nsplit = 148;
        int pps = ceil_div(ceil_div(p.M, nsplit), WP) * WP;
        nsplit = ceil_div(p.M, pps);
        if (nstages >= 2) {
            if (plain) DETRB_LAUNCH((wgrad_narrow_kernel<false>), dim3(nsplit), dim3(WTHREADS), smem, stream, my, mx, p, pps, p.k_mask ? 1 : 0, nstages, tmem_cols);
            else DETRB_LAUNCH((wgrad_narrow_kernel<true>), dim3(nsplit), dim3(WTHREADS), smem, stream, my, mx, p, pps, 0, nstages, tmem_cols);
            DETRB_CHECK_LAUNCH("wgrad_narrow_kernel");
            return DETRB_OK;
        }
    }
    // tile variant <NA, KT> (see the top of the file).  g_wgrad_tile: 0 = policy, 1 = base tile everywhere, 2 / 3 / 4 = <1,256> / <2,128> /
    // <2,256> wherever the shape allows (tests, tuning runs; env DETRB_WGRAD_TILE presets it)
    static bool env_read = false;
    if (!env_read) { env_read = true; if (const char *e = getenv("DETRB_WGRAD_TILE")) g_wgrad_tile = atoi(e); }
    int na = 1, kt = WK;
    if (!sp && !p.a_kb_rows && !p.k_mask) {
        const bool wide_k = p.K >= 256, wide_n = p.N >= 256;
        if (g_wgrad_tile == 0) wgrad_tile_policy(p, &na, &kt);
        else if (g_wgrad_tile == 2 && wide_k) kt = 256;
        else if (g_wgrad_tile == 3 && wide_n) na = 2;
        else if (g_wgrad_tile == 4) { if (wide_k) kt = 256; if (wide_n) na = 2; }
    }
    const int tiles = ceil_div(p.K, kt) * ceil_div(p.N, WN * na);
    // one wave: 2 CTAs per SM of the base tile, 1 CTA per SM of the wider ones (fewer fp32 atomics per gradient element) -- rounded
    // DOWN: 36 tiles x 9 splits = 324 CTAs on 296 slots ran a second, nearly empty wave (ncu, round 2: SMs active 56 % of the 3x3
    // 256->256 kernel's duration)
    int splits = (na * kt == 128 ? 148 * 2 : 148) / tiles;
    const int max_splits = ceil_div(p.M, WP * 4);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int pix_per_split = ceil_div(ceil_div(p.M, splits), WP) * WP;
    splits = ceil_div(p.M, pix_per_split);
    dim3 grid(ceil_div(p.K, kt), ceil_div(p.N, WN * na), splits);
    const int stem_mask = p.k_mask ? 1 : 0;
    // interleaved pixel blocks (env DETRB_WGRAD_INTERLEAVE=1): measured on the full step, no difference (651.2 / 651.5 vs 647.4 / 653.2
    // img/s) -- the two kernels do not stay in lockstep -- so contiguous ranges remain the default
    static int interleave = -1;
    if (interleave < 0) { const char *e = getenv("DETRB_WGRAD_INTERLEAVE"); interleave = e ? atoi(e) : 0; }
#define DETRB_WGRAD_LAUNCH(IM2COL_, SPLIT_, NA_, KT_, MASK_)                                                                            \
    do {                                                                                                                                \
        static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      /* the opt-in is per device */     \
        if (!configured) {                                                                                                              \
            DETRB_CUDA(cudaFuncSetAttribute((wgrad_tc_kernel<IM2COL_, SPLIT_, NA_, KT_>), cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                            wgrad_smem(NA_, KT_)));                                                                     \
            configured = true;                                                                                                          \
        }                                                                                                                               \
        DETRB_LAUNCH((wgrad_tc_kernel<IM2COL_, SPLIT_, NA_, KT_>), dim3(grid), dim3(WTHREADS), wgrad_smem(NA_, KT_), stream, my, mx, my2, mx2, \
                     p, pix_per_split, MASK_, interleave);                                                                              \
    } while (0)
    if (sp) {
        if (plain) DETRB_WGRAD_LAUNCH(false, true, 1, 128, stem_mask);
        else DETRB_WGRAD_LAUNCH(true, true, 1, 128, 0);
    } else if (plain) {
        if (na == 1 && kt == 128) DETRB_WGRAD_LAUNCH(false, false, 1, 128, stem_mask);
        else if (na == 1) DETRB_WGRAD_LAUNCH(false, false, 1, 256, 0);
        else if (kt == 128) DETRB_WGRAD_LAUNCH(false, false, 2, 128, 0);
        else DETRB_WGRAD_LAUNCH(false, false, 2, 256, 0);
    } else {
        if (na == 1 && kt == 128) DETRB_WGRAD_LAUNCH(true, false, 1, 128, 0);
        else if (na == 1) DETRB_WGRAD_LAUNCH(true, false, 1, 256, 0);
        else if (kt == 128) DETRB_WGRAD_LAUNCH(true, false, 2, 128, 0);
        else DETRB_WGRAD_LAUNCH(true, false, 2, 256, 0);
    }
#undef DETRB_WGRAD_LAUNCH
    DETRB_CHECK_LAUNCH("wgrad_tc_kernel");
    return DETRB_OK;
}

static thread_local int g_wgrad_tc_enabled = 1;     // validated on B200 (tests/test_gemm_tc_gpu.py::test_wgrad_tc)
extern "C" int detrb_set_tc_wgrad(int enable) { int old = g_wgrad_tc_enabled; g_wgrad_tc_enabled = enable; return old; }
bool detrb_wgrad_tc_enabled() { return g_wgrad_tc_enabled != 0; }

// test entry: force the tcgen05 weight-gradient kernel
extern "C" int detrb_wgrad_tc_force(const detrb_wgrad_t *pp, detrb_stream_t stream)
{
    if (!pp) DETRB_FAIL(DETRB_E_BADARG, "detrb_wgrad_tc_force: null params");
    if (!detrb_wgrad_tc_supported(*pp)) DETRB_FAIL(DETRB_E_SHAPE, "detrb_wgrad_tc_force: problem not supported");
    return detrb_wgrad_tc(*pp, (cudaStream_t)stream);
}
