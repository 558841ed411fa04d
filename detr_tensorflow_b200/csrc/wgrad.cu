// wgrad.cu -- weight gradients of conv / linear layers:
//     dW[n, k] += rowscale[n] * sum_m dY[m, n] * gather(A)[m, k]          (fp32, atomics across pixel splits)
// The reduction runs over the pixel/token dimension m, so both operands are staged [pixel][channel] in
// shared memory and read through ldmatrix.trans (mma.sync bf16, fp32 accumulate).
// Replaces the kernel gradients of tape.gradient (training.py:23 / optimizers.py:115).
#include "common.cuh"

namespace {

constexpr int TN = 128;            // rows of dW per CTA (output channels)
constexpr int TK = 128;            // cols of dW per CTA (taps*Cin)
constexpr int BP = 32;             // pixels per pipeline stage
constexpr int LDT = 128 + 8;       // smem row stride (bf16): 272 B
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;

template <bool STEM>
__global__ void __launch_bounds__(NTHREADS)
wgrad_kernel(const detrb_wgrad_t p, int pix_per_split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16 *sY = reinterpret_cast<bf16 *>(smem_raw);          // [STAGES][BP][LDT]  dY tile  (pixel, n)
    bf16 *sA = sY + STAGES * BP * LDT;                      // [STAGES][BP][LDT]  A tile   (pixel, k)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wn = warp >> 2, wk = warp & 3;                // 2 (n) x 4 (k) warps; warp tile 64 x 32
    const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
    const int m_begin = blockIdx.z * pix_per_split;
    const int m_end = min(p.M, m_begin + pix_per_split);
    if (m_begin >= m_end) return;
    // parity precision (p.split): three passes over the pixel range -- dY_hi^T A_hi, dY_lo^T A_hi, dY_hi^T A_lo
    const int nsteps1 = (m_end - m_begin + BP - 1) / BP;
    const int nsteps = p.split ? 3 * nsteps1 : nsteps1;

    const bf16 *A = reinterpret_cast<const bf16 *>(p.A);
    const bf16 *dY = reinterpret_cast<const bf16 *>(p.dY);
    const int ohw = p.OH * p.OW;
    const int npad = (p.N + 7) & ~7;

    // ---- fixed per-thread column assignment
    // general: 16 chunks of 8 k per pixel row, 2 pixel rows per thread; stem: 32 chunks of 4 k (one tap), 4 rows
    constexpr int A_ROWS = STEM ? 4 : 2;
    constexpr int A_ROW_STEP = STEM ? 8 : 16;
    const int a_chunk = STEM ? (tid & 31) : (tid & 15);
    const int a_row0 = STEM ? (tid >> 5) : (tid >> 4);
    const int ka = k0 + a_chunk * (STEM ? 4 : 8);
    bool ka_ok = ka < p.K;
    int a_kh = 0, a_kw = 0, a_c = 0;
    if (ka_ok) {
        if (STEM) {
            // 7x7 stem stored as 7 x 8 taps x 4 channels: the 8th tap (and the 4th channel) are padding
            int tap = ka >> 2; a_kh = tap >> 3; a_kw = tap & 7; a_c = 0;
            if (a_kw >= 7) ka_ok = false;
        } else { int tap = ka / p.Cin; a_c = ka - tap * p.Cin; a_kh = tap / p.KW; a_kw = tap - a_kh * p.KW; }
    }
    const int y_chunk = tid & 15, y_row0 = tid >> 4;
    const int ny = n0 + y_chunk * 8;
    const bool ny_ok = ny < npad;

    auto load_stage = [&](int stage, int step) {
        bf16 *y_dst = sY + stage * BP * LDT;
        bf16 *a_dst = sA + stage * BP * LDT;
        const int part = step / nsteps1;
        const long long y_off = part == 1 ? p.split : 0, a_off = part == 2 ? p.split : 0;
        const int mb = m_begin + (step - part * nsteps1) * BP;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            int r = y_row0 + i * 16;
            int m = mb + r;
            bool ok = ny_ok && m < m_end;
            const bf16 *src = ok ? dY + y_off + (size_t)m * p.ldy + ny : dY;
            cp_async16(smem_u32(y_dst + r * LDT + y_chunk * 8), src, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < A_ROWS; i++) {
            int r = a_row0 + i * A_ROW_STEP;
            int m = mb + r;
            bool ok = ka_ok && m < m_end;
            const bf16 *src = A;
            if (ok) {
                int b = m / ohw, rem = m - b * ohw;
                int oy = rem / p.OW, ox = rem - oy * p.OW;
                int iy = oy * p.stride - p.pad + a_kh, ix = ox * p.stride - p.pad + a_kw;
                ok = iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
                if (ok) src = A + a_off + (((size_t)b * p.IH + iy) * p.IW + ix) * p.lda + a_c;
            }
            if (STEM) cp_async8(smem_u32(a_dst + r * LDT + a_chunk * 4), src, ok ? 8 : 0);
            else      cp_async16(smem_u32(a_dst + r * LDT + a_chunk * 8), src, ok ? 16 : 0);
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[i][j][k] = 0.f;
    const bool do_bias = p.dbias != nullptr && blockIdx.x == 0 && tid < TN;     // one k-tile column of CTAs owns the bias
    float bias_acc = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nsteps) load_stage(s, s);
        cp_async_commit();
    }
    for (int step = 0; step < nsteps; step++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = step + STAGES - 1;
            if (nxt < nsteps) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const bf16 *y_s = sY + (step % STAGES) * BP * LDT;
        const bf16 *a_s = sA + (step % STAGES) * BP * LDT;
        if (do_bias && step < 2 * nsteps1) {              // bias gradient = column sums of dY (hi and lo passes): free ride on the staged tile
#pragma unroll 8
            for (int pix = 0; pix < BP; pix++) bias_acc += __bfloat162float(y_s[pix * LDT + tid]);
        }
#pragma unroll
        for (int kk = 0; kk < BP / 16; kk++) {
            uint32_t af[4][4], bfr[4][2];
            const int pb = kk * 16;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int pix = pb + (lane & 7) + ((lane >> 4) << 3);
                int col = wn * 64 + i * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(af[i][0], af[i][1], af[i][2], af[i][3], smem_u32(y_s + pix * LDT + col));
            }
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                int pix = pb + (lane & 7) + ((lane >> 3) & 1) * 8;
                int col = wk * 32 + j * 8 + (lane >> 4) * 8;
                ldmatrix_x4_trans(bfr[j][0], bfr[j][1], bfr[j + 1][0], bfr[j + 1][1], smem_u32(a_s + pix * LDT + col));
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) mma_bf16_16816(acc[i][j], af[i], bfr[j]);
        }
    }
    cp_async_wait<0>();

    if (do_bias && n0 + tid < p.N) atomicAdd(p.dbias + n0 + tid, bias_acc * (p.rowscale ? p.rowscale[n0 + tid] : 1.f));
    const int g = lane >> 2, t = lane & 3;
    const bool vec_ok = (p.ldw % 2 == 0) && (((uintptr_t)p.dW & 7) == 0);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int n = n0 + wn * 64 + i * 16 + g + h * 8;
            if (n >= p.N) continue;
            float sc = p.rowscale ? p.rowscale[n] : 1.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int k = k0 + wk * 32 + j * 8 + t * 2;
                float *dst = p.dW + (size_t)n * p.ldw + k;
                if (k + 1 < p.K && vec_ok) {            // two adjacent columns: one vector reduction
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(dst), "f"(acc[i][j][h * 2 + 0] * sc), "f"(acc[i][j][h * 2 + 1] * sc) : "memory");
                } else {
                    if (k < p.K)     atomicAdd(dst,     acc[i][j][h * 2 + 0] * sc);
                    if (k + 1 < p.K) atomicAdd(dst + 1, acc[i][j][h * 2 + 1] * sc);
                }
            }
        }
}

}  // namespace

extern "C" int detrb_wgrad(const detrb_wgrad_t *pp, detrb_stream_t stream_)
{
    if (!pp) DETRB_FAIL(DETRB_E_BADARG, "detrb_wgrad: null params");
    detrb_wgrad_t p = *pp;
    cudaStream_t stream = (cudaStream_t)stream_;
    DETRB_REQUIRE(p.A && p.dY && p.dW, "detrb_wgrad: null pointer");
    DETRB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "detrb_wgrad: empty problem");
    DETRB_REQUIRE(p.M == p.batch * p.OH * p.OW, "detrb_wgrad: M=%d != batch*OH*OW", p.M);
    DETRB_REQUIRE(p.ldy >= ((p.N + 7) & ~7) && p.ldy % 8 == 0, "detrb_wgrad: ldy=%d must cover N=%d rounded to 8", p.ldy, p.N);
    DETRB_REQUIRE(p.split >= 0 && p.split % 8 == 0, "detrb_wgrad: split=%lld", (long long)p.split);
    if (p.a_kb_rows || p.k_mask) {                             // sliding-window A / stem column mask: tcgen05 kernel only
        DETRB_REQUIRE(p.a_kb_rows >= 0 && p.KH == 1 && p.KW == 1 && p.stride == 1 && p.pad == 0 && detrb_wgrad_tc_supported(p),
                      "detrb_wgrad: a_kb_rows / k_mask need plain geometry and the tcgen05 kernel");
        return detrb_wgrad_tc(p, stream);
    }
    if (detrb_wgrad_tc_enabled() && detrb_wgrad_tc_supported(p) && detrb_wgrad_tc_profitable(p)) {           // tcgen05 / TMA im2col / TMEM
        return detrb_wgrad_tc(p, stream);                      // bias gradient fused (k-tile 0 CTAs)
    }
    const bool stem = (p.Cin == 4);
    if (stem) DETRB_REQUIRE(p.KW == 8 && p.K == p.KH * 32 && p.lda == 4 && !p.split, "detrb_wgrad: stem geometry");
    else DETRB_REQUIRE(p.Cin % 8 == 0 && p.K == p.KH * p.KW * p.Cin && p.lda % 8 == 0, "detrb_wgrad: Cin=%d K=%d lda=%d", p.Cin, p.K, p.lda);

    const int tiles = ceil_div(p.K, TK) * ceil_div(p.N, TN);
    int splits = ceil_div(148 * 3, tiles);
    int max_splits = ceil_div(p.M, BP * 4);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int pix_per_split = ceil_div(ceil_div(p.M, splits), BP) * BP;
    splits = ceil_div(p.M, pix_per_split);
    constexpr int smem = STAGES * 2 * BP * LDT * (int)sizeof(bf16);
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        DETRB_CUDA(cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(ceil_div(p.K, TK), ceil_div(p.N, TN), splits);
    if (stem) DETRB_LAUNCH((wgrad_kernel<true>), dim3(grid), dim3(NTHREADS), smem, stream, p, pix_per_split);
    else      DETRB_LAUNCH((wgrad_kernel<false>), dim3(grid), dim3(NTHREADS), smem, stream, p, pix_per_split);
    DETRB_CHECK_LAUNCH("wgrad_kernel");
    return DETRB_OK;                                       // bias gradient fused (column sums of the staged dY tiles)
}
