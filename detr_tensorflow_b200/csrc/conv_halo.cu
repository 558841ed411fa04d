// conv_halo.cu -- 3x3 / stride 1 / pad 1 convolution with 64 input and 64 output channels, forward and data gradient: the conv2 of
// every layer1 bottleneck (resnet_backbone.py:116-136 at 200 x 334 pixels, M = 534 400 rows at batch 8).
//
// As an implicit GEMM through TMA im2col (gemm_tc.cu) this layer reads its input NINE times from L2 (616 MB per launch through the
// L2 -> SM path: 97 us, L2-bandwidth bound, tensor pipe 21 %).  Here every input pixel enters shared memory ONCE:
//   * a persistent CTA owns a run of consecutive image rows of one 128-pixel-wide column strip; per output row it loads ONE new
//     input row segment (130 pixels x 64 channels = 16.6 KB, TMA tiled mode, zero fill outside the image = the padding) into a ring
//     of seven row slots;
//   * the nine filter taps are nine VIEWS of the three resident rows: a row slot is a K-major 128-byte-swizzled operand tile of 130
//     pixel rows, and tap (kh, kw) is the UMMA descriptor of row slot y + kh - 1 advanced by kw pixel rows (kw * 128 bytes; the
//     swizzle follows the absolute address, so the shifted view needs nothing else) -- 36 tcgen05.mma (128 x 64 x 16) per output
//     row, no copies, no im2col;
//   * the 72 KB of weights stay resident in shared memory; 4 accumulators rotate through tensor memory; the epilogue (bias = folded
//     BN shift, ReLU, 1-bit ReLU masks in / out) is the straight-line chunk of tc_common.cuh; results leave by TMA store.
// Algorithmic bytes: input + output once (137 MB per launch at batch 8) + 1.6 % halo columns + two halo rows per run.
#include "tc_common.cuh"

namespace {

constexpr int HC = 64;                              // channels in = channels out
constexpr int ROW_PIX = TBM + 2;                    // 128 outputs + one halo pixel on each side
constexpr int ROW_TX = ROW_PIX * HC * 2;            // bytes one row load delivers (16 640)
constexpr int ROW_BYTES = 17 * 1024;                // slot pitch: the next multiple of the 1024-byte swizzle period
constexpr int NR = 7;                               // input-row ring: 3 rows in use + 4 in flight (5 slots left the MMA warp waiting for rows: 65 vs 35 us of MMA time)
constexpr int RS = 2;                               // output staging slots: one per epilogue warpgroup (it works on every other unit)
constexpr int W_TAP_BYTES = HC * 128;               // one tap of the filter: [64 out] x [64 in] bf16 = 8 KB
constexpr int W_BYTES = 9 * W_TAP_BYTES;            // 72 KB
constexpr int NACC = 4;
constexpr int OFF_ROWS = W_BYTES, OFF_OUT = OFF_ROWS + NR * ROW_BYTES, OFF_BAR = OFF_OUT + RS * 16384, OFF_BIAS = OFF_BAR + 256;
constexpr int SMEM_BYTES = OFF_BIAS + HC * 4 + 1024;            // + slack for the 1024-byte alignment of the dynamic smem base
constexpr int NTHREADS = 384;
static_assert(SMEM_BYTES <= 227 * 1024, "conv_halo: shared memory");

struct HaloGeom { int B, H, W, nseg, units; int wtap[9]; };

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c, int w, int h, int n) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_y, const detrb_igemm_t p, const HaloGeom geo, const int base_offset_mode, const int diag)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem0 + OFF_BAR;
    auto row_full = [&](int s) { return bar0 + 8u * s; };
    auto row_empty = [&](int s) { return bar0 + 8u * (NR + s); };
    auto tmem_full = [&](int a) { return bar0 + 8u * (2 * NR + a); };
    auto tmem_empty = [&](int a) { return bar0 + 8u * (2 * NR + NACC + a); };
    auto slot_free = [&](int i) { return bar0 + 8u * (2 * NR + 2 * NACC + i); };
    const uint32_t w_full = bar0 + 8u * (2 * NR + 2 * NACC + RS);
    const uint32_t tmem_slot = w_full + 8u;
    const uint32_t sbias = smem0 + OFF_BIAS;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's run of units; unit u = (strip, y), strip = (image, 128-pixel column segment)
    const int u_begin = (int)((long long)blockIdx.x * geo.units / gridDim.x), u_end = (int)((long long)(blockIdx.x + 1) * geo.units / gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_y);
        for (int s = 0; s < NR; s++) { mbar_init(row_full(s), 1); mbar_init(row_empty(s), 1); }
        for (int a = 0; a < NACC; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 8); }
        for (int i = 0; i < RS; i++) mbar_init(slot_free(i), 1);
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)(NACC * HC)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===================== weights once, then one new input row per unit (three at the start of a run) =====================
        if (elect_one()) {
            mbar_expect_tx(w_full, W_BYTES);
            for (int t = 0; t < 9; t++) tma_load_2d(smem0 + (uint32_t)(t * W_TAP_BYTES), &map_w, w_full, t * HC, 0);
            int seq = 0;
            auto load_row = [&](int x0, int yrow, int b) {
                const int slot = seq % NR;
                mbar_wait(row_empty(slot), ((seq / NR) & 1) ^ 1);
                mbar_expect_tx(row_full(slot), ROW_TX);
                tma_load_4d(smem0 + OFF_ROWS + (uint32_t)(slot * ROW_BYTES), &map_x, row_full(slot), 0, x0 - 1, yrow, b);
                seq++;
            };
            for (int u = u_begin; u < u_end; u++) {
                const int strip = u / geo.H, y = u - strip * geo.H;
                const int b = strip / geo.nseg, x0 = (strip - b * geo.nseg) * TBM;
                if (u == u_begin || y == 0) { load_row(x0, y - 1, b); load_row(x0, y, b); }
                load_row(x0, y + 1, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: nine taps = nine views of the three resident rows =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(TBM, HC);
            mbar_wait(w_full, 0);
            int top = 0, next_seq = 0, it = 0;
            for (int u = u_begin; u < u_end; u++, it++) {
                const int strip = u / geo.H, y = u - strip * geo.H;
                const bool fresh = (u == u_begin || y == 0);
                if (fresh) { top = next_seq; next_seq += 3; } else { top += 1; next_seq += 1; }
                for (int r = fresh ? 0 : 2; r < 3; r++) {              // rows that this unit is the first to use
                    const int s = top + r;
                    mbar_wait(row_full(s % NR), (s / NR) & 1);
                }
                const int acc = it % NACC;
                mbar_wait(tmem_empty(acc), ((it / NACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * HC);
#pragma unroll
                for (int kh = 0; kh < 3; kh++) {
                    const uint32_t row_addr = smem0 + OFF_ROWS + (uint32_t)(((top + kh) % NR) * ROW_BYTES);
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) {
                        if ((diag & 2) && (kh | kw)) continue;                                   // developer switch: one tap only
                        const uint32_t a_addr = row_addr + ((diag & 1) ? 0u : (uint32_t)(kw * 128));   // developer switch: aligned views only
                        // the 128-byte swizzle is a function of the ABSOLUTE shared-memory address bits [7,10) -- for the TMA write and for
                        // the UMMA read alike -- so a start address shifted by whole 128-byte rows reads the right chunks with the
                        // descriptor's base-offset field left at 0 (measured on B200: bit-exact against the im2col kernel; setting the
                        // field to (addr >> 7) & 7 gives wrong results)
                        uint64_t da = make_smem_desc(a_addr);
                        if (base_offset_mode) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
                        const uint64_t db = make_smem_desc(smem0 + (uint32_t)(geo.wtap[kh * 3 + kw] * W_TAP_BYTES));
#pragma unroll
                        for (int k = 0; k < HC / 16; k++)
                            tc_mma_f16(d_addr, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kh | kw | k) != 0);
                    }
                }
                tc_commit(tmem_full(acc));
                // rows nobody needs any more: the top row, and all three when the run ends here
                const bool run_ends = (u + 1 == u_end) || (y + 1 == geo.H);
                tc_commit(row_empty(top % NR));
                if (run_ends) { tc_commit(row_empty((top + 1) % NR)); tc_commit(row_empty((top + 2) % NR)); }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: unit g goes to warpgroup g % 2 =====================
        const int wg = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
        const bool leader = (q == 0 && lane == 0);
        for (int t = threadIdx.x - 128; t < HC; t += 256)
            asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbias + 4u * (uint32_t)t), "f"(p.bias ? p.bias[t] : 0.f) : "memory");
        asm volatile("bar.sync 3, 256;" ::: "memory");
        const bool relu = p.relu != 0;
        const bool mbits = p.mask_bits != nullptr, obits = p.out_bits != nullptr;
        auto pixel = [&](int u, int &x0, int &y, int &b) -> long long {       // flat NHWC pixel index of this thread's row of unit u (-1: outside)
            const int strip = u / geo.H;
            y = u - strip * geo.H;
            b = strip / geo.nseg;
            x0 = (strip - b * geo.nseg) * TBM;
            return (x0 + row < geo.W) ? ((long long)(b * geo.H + y) * geo.W + x0 + row) : -1ll;
        };
        uint2 mb_cur = make_uint2(0u, 0u), mb_nxt = make_uint2(0u, 0u);
        auto load_bits = [&](int u) -> uint2 {
            int x0, y, b;
            if (u >= u_end) return make_uint2(0u, 0u);
            const long long m = pixel(u, x0, y, b);
            return m >= 0 ? ld_bits8(p.mask_bits + (size_t)m * p.ldmb) : make_uint2(0u, 0u);
        };
        if (mbits) mb_nxt = load_bits(u_begin);
        int it = 0;
        for (int u = u_begin; u < u_end; u++, it++) {
            const int acc = it % NACC;
            if (mbits) { mb_cur = mb_nxt; mb_nxt = load_bits(u + 1); }
            mbar_wait(tmem_full(acc), (it / NACC) & 1);
            tc_fence_after();
            if ((it & 1) == wg) {
                int x0, y, b;
                const long long m = pixel(u, x0, y, b);
                const int slot = it % RS;
                const uint32_t sbuf = smem0 + OFF_OUT + (uint32_t)slot * 16384u;
                uint32_t acc_r[64];
                tc_ld64(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * HC), acc_r);
                mbar_wait(slot_free(slot), ((it / RS) & 1) ^ 1);       // the store that last used this slot has read it
                uint4 rr[8];
                uint2 ob = make_uint2(0u, 0u);
                tc_wait_ld();
                // the accumulator is in registers: hand it back to the MMA warp now, not after the store (the leader waits for the
                // TMA store to read the slot, which would hold the accumulator for another microsecond)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty(acc));
                if (diag & 4) {}                                                                  // developer switch: no epilogue arithmetic
                else if (mbits) epi_chunk_math<false, true, false>(acc_r, rr, sbias, relu, mb_cur, p.mask_scale, ob, sbuf + row_off, sw);
                else if (obits) epi_chunk_math<false, false, true>(acc_r, rr, sbias, relu, mb_cur, 1.f, ob, sbuf + row_off, sw);
                else epi_chunk_math<false, false, false>(acc_r, rr, sbias, relu, mb_cur, 1.f, ob, sbuf + row_off, sw);
                if (obits && m >= 0) *reinterpret_cast<uint2 *>(p.out_bits + (size_t)m * p.ldob) = ob;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
                if (leader) {
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 :: "l"(&map_y), "r"(sbuf), "r"(0), "r"(x0), "r"(y), "r"(b) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the warpgroup's only slot: free as soon as the store has read it
                    mbar_arrive(slot_free(slot));
                }
            } else {
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(NACC * HC)) : "memory");
    }
}

// ================================================================================================ row-stationary variant
// The kernel above issues 36 MMAs of 128 x 64 x 16 per output row: each fetches 4 KB of pixels + 2 KB of weights for 32 cycles of
// tensor-pipe work, and at the ~64 B/clk the SM delivers operands that is 86 cycles per MMA (ncu + diag runs: MMA bound at 52 us).
// Here an INPUT row is pushed through the tensor core once against all the taps that use it: input row r is tap row kh = 0 of output
// row r+1, kh = 1 of output row r and kh = 2 of output row r-1, so one MMA of N = 192 with the weights stacked [W(2,kw); W(1,kw);
// W(0,kw)] adds its contribution to THREE accumulators at once -- 12 MMAs (10 KB per 3x the work) instead of 36 per row.  The
// accumulators form a ring of eight 64-column slots in tensor memory (output row i of the CTA lives in slot i % 8; a triple that
// wraps around the ring is issued as two MMAs), every MMA accumulates, and a slot is zeroed by the epilogue warpgroup that drained
// it (tcgen05.st) before the MMA warp may use it again.  Same accumulation order per output element as above (kh, kw, c): bit-exact.
constexpr int NSLOT = 8;                            // accumulator ring (8 x 64 = all 512 TMEM columns)
// input-row ring: a row is needed only while its own twelve MMAs run (seven 16 640-byte slots at 128-byte alignment were tried: wrong
// results at full size -- the TMA write does want the 1024-byte alignment of the swizzle period -- and no faster)
constexpr int NR2 = 5;
constexpr int ROW_PITCH2 = ROW_BYTES;
constexpr int RS2 = 4;                              // output staging tiles: two per epilogue warpgroup, released one store later
constexpr int OFF_ROWS2 = W_BYTES, OFF_OUT2 = OFF_ROWS2 + NR2 * ROW_PITCH2, OFF_BAR2 = OFF_OUT2 + RS2 * 16384, OFF_BIAS2 = OFF_BAR2 + 256;
constexpr int SMEM_BYTES2 = OFF_BIAS2 + HC * 4 + 1024;

__device__ __forceinline__ void tmem_zero64(uint32_t taddr) {
    const uint32_t z = 0u;
#pragma unroll
    for (int c = 0; c < 64; c += 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                     :: "r"(taddr + (uint32_t)c), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_halo192_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ CUtensorMap map_y, const detrb_igemm_t p, const HaloGeom geo, const int diag)
{   // diag (developer switch DETRB_HALO_DIAG, wrong results, timing only): 1 = MMAs of kw 0 only, 2 = no epilogue arithmetic, 4 = N <= 64 per MMA
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem0 + OFF_BAR2;
    auto row_full = [&](int s) { return bar0 + 8u * s; };
    auto row_empty = [&](int s) { return bar0 + 8u * (NR2 + s); };
    auto tmem_full = [&](int a) { return bar0 + 8u * (2 * NR2 + a); };
    auto tmem_empty = [&](int a) { return bar0 + 8u * (2 * NR2 + NSLOT + a); };
    auto slot_free = [&](int i) { return bar0 + 8u * (2 * NR2 + 2 * NSLOT + i); };
    const uint32_t w_full = bar0 + 8u * (2 * NR2 + 2 * NSLOT + RS2);
    const uint32_t tmem_slot = w_full + 8u;
    const uint32_t sbias = smem0 + OFF_BIAS2;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = geo.H;
    const int u_begin = (int)((long long)blockIdx.x * geo.units / gridDim.x), u_end = (int)((long long)(blockIdx.x + 1) * geo.units / gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_y);
        for (int s = 0; s < NR2; s++) { mbar_init(row_full(s), 1); mbar_init(row_empty(s), 1); }
        for (int a = 0; a < NSLOT; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 4); }
        for (int i = 0; i < RS2; i++) mbar_init(slot_free(i), 1);
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)(NSLOT * HC)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===================== weights once (stacked per kw: kh = 2, 1, 0), then every input row of every run once =====================
        if (elect_one()) {
            mbar_expect_tx(w_full, W_BYTES);
            for (int kw = 0; kw < 3; kw++)
                for (int j = 0; j < 3; j++)                 // j = 2 - kh
                    tma_load_2d(smem0 + (uint32_t)((kw * 3 + j) * W_TAP_BYTES), &map_w, w_full, geo.wtap[(2 - j) * 3 + kw] * HC, 0);
            int seq = 0;
            for (int u = u_begin; u < u_end;) {
                const int strip = u / H, y0 = u - strip * H, y1 = min(H, y0 + (u_end - u));
                const int b = strip / geo.nseg, x0 = (strip - b * geo.nseg) * TBM;
                for (int r = max(0, y0 - 1); r <= min(H - 1, y1); r++, seq++) {
                    const int slot = seq % NR2;
                    mbar_wait(row_empty(slot), ((seq / NR2) & 1) ^ 1);
                    mbar_expect_tx(row_full(slot), ROW_TX);
                    tma_load_4d(smem0 + OFF_ROWS2 + (uint32_t)(slot * ROW_PITCH2), &map_x, row_full(slot), 0, x0 - 1, r, b);
                }
                u += y1 - y0;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: one input row -> up to three accumulators =====================
        if (elect_one()) {
            mbar_wait(w_full, 0);
            int seq = 0, oi_run = 0;                        // input rows consumed; index (within this CTA) of the run's first output row
            for (int u = u_begin; u < u_end;) {
                const int strip = u / H, y0 = u - strip * H, y1 = min(H, y0 + (u_end - u));
                const int r_last = min(H - 1, y1);
                for (int r = max(0, y0 - 1); r <= r_last; r++, seq++) {
                    const int rslot = seq % NR2;
                    mbar_wait(row_full(rslot), (seq / NR2) & 1);
                    const int lo = max(r - 1, y0), hi = min(r + 1, y1 - 1);    // output rows this input row contributes to
                    // an output row gets its first contribution from input row max(o - 1, 0): its slot must be drained and zeroed
                    for (int o = lo; o <= hi; o++)
                        if (r == max(o - 1, 0)) {
                            const int oi = oi_run + (o - y0);
                            mbar_wait(tmem_empty(oi % NSLOT), (oi / NSLOT) & 1);
                        }
                    tc_fence_after();
                    const uint32_t row_addr = smem0 + OFF_ROWS2 + (uint32_t)(rslot * ROW_PITCH2);
                    const int oi_lo = oi_run + (lo - y0), s_lo = oi_lo % NSLOT, nrows = hi - lo + 1;
                    const int n1 = min(nrows, NSLOT - s_lo), n2 = nrows - n1;   // rows before / after the ring wraps
                    const int j0 = lo - (r - 1);                                // first block of the stack [W(2,kw); W(1,kw); W(0,kw)]
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) {
                        if ((diag & 1) && kw) continue;
                        const uint64_t da = make_smem_desc(row_addr + (uint32_t)(kw * 128));
                        const uint32_t wb = smem0 + (uint32_t)((kw * 3 + j0) * W_TAP_BYTES);
                        const uint64_t db1 = make_smem_desc(wb), db2 = make_smem_desc(wb + (uint32_t)(n1 * W_TAP_BYTES));
#pragma unroll
                        for (int k = 0; k < HC / 16; k++) {
                            tc_mma_f16(tmem_base + (uint32_t)(s_lo * HC), da + (uint64_t)(k * 2), db1 + (uint64_t)(k * 2), make_idesc(TBM, (diag & 4) ? 64 : 64 * n1), 1u);
                            if (n2 > 0) tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), make_idesc(TBM, 64 * n2), 1u);
                        }
                    }
                    tc_commit(row_empty(rslot));                                // this input row is done with
                    if (r - 1 >= y0) tc_commit(tmem_full((oi_run + (r - 1 - y0)) % NSLOT));        // output row r-1 is complete
                    if (r == r_last && r <= y1 - 1) tc_commit(tmem_full((oi_run + (r - y0)) % NSLOT));   // bottom image row: row r too
                }
                oi_run += y1 - y0;
                u += y1 - y0;
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: output row i of the CTA goes to warpgroup i % 2 (slots i % 8: each warpgroup owns four) =====================
        const int wg = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
        const bool leader = (q == 0 && lane == 0);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int t = threadIdx.x - 128; t < HC; t += 256)
            asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbias + 4u * (uint32_t)t), "f"(p.bias ? p.bias[t] : 0.f) : "memory");
        asm volatile("bar.sync 3, 256;" ::: "memory");
        // all accumulators start at zero
        for (int a = wg; a < NSLOT; a += 2) {
            tmem_zero64(lane_addr + (uint32_t)(a * HC));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(a));
        }
        const bool relu = p.relu != 0;
        const bool mbits = p.mask_bits != nullptr, obits = p.out_bits != nullptr;
        int prev_st = -1;
        auto pixel = [&](int u, int &x0, int &y, int &b) -> long long {       // flat NHWC pixel index of this thread's row of unit u (-1: outside)
            const int strip = u / H;
            y = u - strip * H;
            b = strip / geo.nseg;
            x0 = (strip - b * geo.nseg) * TBM;
            return (x0 + row < geo.W) ? ((long long)(b * H + y) * geo.W + x0 + row) : -1ll;
        };
        auto load_bits = [&](int u) -> uint2 {                                 // mask bits of this warpgroup's next row, one row ahead
            int x0, y, b;
            if (u >= u_end) return make_uint2(0u, 0u);
            const long long m = pixel(u, x0, y, b);
            return m >= 0 ? ld_bits8(p.mask_bits + (size_t)m * p.ldmb) : make_uint2(0u, 0u);
        };
        uint2 mb_nxt = mbits ? load_bits(u_begin + wg) : make_uint2(0u, 0u);
        int oi = 0, nstore = 0;
        for (int u = u_begin; u < u_end; u++, oi++) {
            if ((oi & 1) != wg) continue;
            int x0, y, b;
            const long long m = pixel(u, x0, y, b);
            uint2 mbc = mb_nxt, ob = make_uint2(0u, 0u);
            if (mbits) mb_nxt = load_bits(u + 2);
            const int a = oi % NSLOT;
            mbar_wait(tmem_full(a), (oi / NSLOT) & 1);
            tc_fence_after();
            uint32_t acc_r[64];
            tc_ld64(lane_addr + (uint32_t)(a * HC), acc_r);
            const int st = 2 * (nstore & 1) + wg;                  // this warpgroup's two staging tiles alternate
            const uint32_t sbuf = smem0 + OFF_OUT2 + (uint32_t)st * 16384u;
            mbar_wait(slot_free(st), ((nstore >> 1) & 1) ^ 1);     // the store that last used this tile has read it
            tc_wait_ld();
            tmem_zero64(lane_addr + (uint32_t)(a * HC));           // drained: zero it and hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(a));
            uint4 rr[8];
            if (diag & 2) {}
            else if (mbits) epi_chunk_math<false, true, false>(acc_r, rr, sbias, relu, mbc, p.mask_scale, ob, sbuf + row_off, sw);
            else if (obits) epi_chunk_math<false, false, true>(acc_r, rr, sbias, relu, mbc, 1.f, ob, sbuf + row_off, sw);
            else epi_chunk_math<false, false, false>(acc_r, rr, sbias, relu, mbc, 1.f, ob, sbuf + row_off, sw);
            if (obits && m >= 0) *reinterpret_cast<uint2 *>(p.out_bits + (size_t)m * p.ldob) = ob;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
            if (leader) {
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             :: "l"(&map_y), "r"(sbuf), "r"(0), "r"(x0), "r"(y), "r"(b) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (prev_st >= 0) {                                   // the previous row's store has read its tile by now (or we wait for it)
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    mbar_arrive(slot_free(prev_st));
                }
                prev_st = st;
            }
            nstore++;
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(NSLOT * HC)) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
// dense NHWC bf16 tensor [B, H, W, 64] as a 4-D tiled map (C, W, H, B); box = one row segment of box_w pixels, 128-byte swizzle,
// zero fill outside the tensor (= the convolution's padding, also for negative coordinates)
bool make_nhwc_map(CUtensorMap *map, const void *base, int B, int H, int W, int box_w)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)HC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)HC * 2, (cuuint64_t)W * HC * 2, (cuuint64_t)H * W * HC * 2};
    cuuint32_t box[4] = {(cuuint32_t)HC, (cuuint32_t)box_w, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

thread_local int g_halo = 1;          // 0 off, 1 on                                                  (env DETRB_HALO)
int g_halo_bo = 0;            // descriptor base-offset field for the shifted tap views: must stay 0 (see the kernel)   (env DETRB_HALO_BO, developer switch)
int g_halo_diag = 0;          // timing-only switches (wrong results)                                                    (env DETRB_HALO_DIAG)
int g_halo_192 = 1;           // row-stationary N = 192 variant (conv3x3_halo192_kernel)                                  (env DETRB_HALO_192)
void read_env()
{
    static thread_local bool done = false;     // per thread: g_halo is (the developer switches below are process-wide, re-reading them is idempotent)
    if (done) return;
    done = true;
    if (const char *e = getenv("DETRB_HALO")) g_halo = atoi(e);
    if (const char *e = getenv("DETRB_HALO_BO")) g_halo_bo = atoi(e);
    if (const char *e = getenv("DETRB_HALO_DIAG")) g_halo_diag = atoi(e);
    if (const char *e = getenv("DETRB_HALO_192")) g_halo_192 = atoi(e);
}

}  // namespace

extern "C" int detrb_set_tc_halo(int enable) { read_env(); int old = g_halo; g_halo = enable; return old; }

bool detrb_conv_halo_supported(const detrb_igemm_t &p)
{
    read_env();
    if (!g_halo || p.split || p.a_kb_rows) return false;
    if (p.KH != 3 || p.KW != 3 || p.stride != 1 || p.pad != 1 || p.Cin != HC || p.N != HC || p.K != 9 * HC) return false;
    if (p.lda != HC || p.ldc != HC || p.ldw != 9 * HC || p.IH != p.OH || p.IW != p.OW || p.M != p.batch * p.OH * p.OW) return false;
    if (!p.C || p.Cf || p.out_stride > 1 || p.accumulate || p.residual || p.mask || p.sigmoid || p.drop_p > 0.f) return false;
    if (p.mask_bits && (p.out_bits || p.ldmb != HC / 8 || ((uintptr_t)p.mask_bits & 7))) return false;
    if (p.out_bits && (p.ldob != HC / 8 || ((uintptr_t)p.out_bits & 7))) return false;
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.C & 15) || ((uintptr_t)p.W & 15)) return false;
    return encode_fn() != nullptr;
}

int detrb_conv_halo(const detrb_igemm_t &p, cudaStream_t stream)
{
    CUtensorMap mx, mw, my;
    if (!make_nhwc_map(&mx, p.A, p.batch, p.IH, p.IW, ROW_PIX) || !make_nhwc_map(&my, p.C, p.batch, p.OH, p.OW, TBM) ||
        !detrb_make_tiled_map(&mw, p.W, (uint64_t)HC, (uint64_t)(9 * HC), (uint64_t)p.ldw, HC, 64))
        DETRB_FAIL(DETRB_E_CUDA, "conv_halo: cuTensorMapEncodeTiled failed (B=%d H=%d W=%d)", p.batch, p.IH, p.IW);
    HaloGeom geo;
    geo.B = p.batch; geo.H = p.OH; geo.W = p.OW; geo.nseg = ceil_div(p.OW, TBM);
    geo.units = geo.B * geo.nseg * geo.H;
    // forward: tap (kh, kw) pairs with weight tap kh*3+kw; data gradient (mode 1, stride 1): the flipped kernel
    for (int i = 0; i < 9; i++) geo.wtap[i] = p.mode == 1 ? 8 - i : i;
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    static int num_sms = 148;
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute(conv3x3_halo192_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES2));
        DETRB_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int grid = geo.units < num_sms ? geo.units : num_sms;
    if (g_halo_192) {
        DETRB_LAUNCH(conv3x3_halo192_kernel, dim3(grid), dim3(NTHREADS), SMEM_BYTES2, stream, mx, mw, my, p, geo, g_halo_diag);
        DETRB_CHECK_LAUNCH("conv3x3_halo192_kernel");
        return DETRB_OK;
    }
    DETRB_LAUNCH(conv3x3_halo_kernel, dim3(grid), dim3(NTHREADS), SMEM_BYTES, stream, mx, mw, my, p, geo, g_halo_bo, g_halo_diag);
    DETRB_CHECK_LAUNCH("conv3x3_halo_kernel");
    return DETRB_OK;
}
