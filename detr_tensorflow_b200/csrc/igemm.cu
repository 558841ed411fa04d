// igemm.cu -- implicit-GEMM convolution / linear layer on the legacy tensor path (mma.sync bf16,
// fp32 accumulate), cp.async multistage pipeline.  This is the *general* kernel: it handles every
// gather the DETR train step needs (1x1, 3x3, 7x7-stem, strided, transposed/data-gradient) and every
// epilogue.  The tcgen05/TMA kernels in gemm_tc.cu take over the shapes they support.
//
// Replaces: Conv2D / ZeroPadding2D / FrozenBatchNorm2D / ReLU / residual add
// (networks/resnet_backbone.py:20-26,116-136), Linear (custom_layers.py:49-50), all tf.matmul's of
// transformer.py:294-347 and their data gradients.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int LDS = BK + 8;          // smem row stride (bf16): 80 B -> conflict-free ldmatrix
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;

struct RowInfo {
    const bf16 *base;   // image base pointer (batch b), nullptr if the row is out of range
    int y0, x0;
};

template <int BN, bool STEM>
__global__ void __launch_bounds__(NTHREADS)
igemm_kernel(const detrb_igemm_t p)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    constexpr int WN = (BN == 128) ? 4 : 2;      // warps along N
    constexpr int WM = 8 / WN;                   // warps along M
    constexpr int WTM = BM / WM;                 // warp tile M (64 or 32)
    constexpr int WTN = BN / WN;                 // warp tile N (32)
    constexpr int MT = WTM / 16;                 // m16 tiles per warp
    constexpr int NT = WTN / 8;                  // n8 tiles per warp (4)
    constexpr int A_ROWS = STEM ? 4 : 2;         // A rows handled per thread
    constexpr int B_ROWS = BN / 64;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16 *sA = reinterpret_cast<bf16 *>(smem_raw);
    bf16 *sB = sA + STAGES * BM * LDS;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    const bf16 *A = reinterpret_cast<const bf16 *>(p.A);
    const bf16 *W = reinterpret_cast<const bf16 *>(p.W);

    // ---- per-thread gather setup
    RowInfo rows[A_ROWS];
    const int a_chunk = STEM ? (tid & 7) : (tid & 3);
    const int a_row0 = STEM ? (tid >> 3) : (tid >> 2);
    constexpr int A_ROW_STEP = STEM ? 32 : 64;
    const int ohw = p.OH * p.OW;
#pragma unroll
    for (int i = 0; i < A_ROWS; i++) {
        int m = m0 + a_row0 + i * A_ROW_STEP;
        if (m < p.M) {
            int b = m / ohw, rem = m - b * ohw;
            int oy = rem / p.OW, ox = rem - oy * p.OW;
            rows[i].base = A + (size_t)b * p.IH * p.IW * p.lda;
            if (p.mode == 0) { rows[i].y0 = oy * p.stride - p.pad; rows[i].x0 = ox * p.stride - p.pad; }
            else             { rows[i].y0 = oy + p.pad;            rows[i].x0 = ox + p.pad; }
        } else {
            rows[i].base = nullptr; rows[i].y0 = 0; rows[i].x0 = 0;
        }
    }
    // parity precision (p.split): three passes over the k-loop -- A_hi*W_hi, A_lo*W_hi, A_hi*W_lo -- into the same accumulators
    const int nk1 = p.K / BK;
    const int nk = p.split ? 3 * nk1 : nk1;

    auto load_stage = [&](int stage, int kb) {
        bf16 *a_dst = sA + stage * BM * LDS;
        bf16 *b_dst = sB + stage * BN * LDS;
        const int part = kb / nk1;
        const int k0 = (kb - part * nk1) * BK;
        const long long a_off = part == 1 ? p.split : 0, w_off = part == 2 ? p.wsplit : 0;
        if (STEM) {
            // k-block == one kernel row kh; 8 taps kw (8th has zero weights) x 4 channels
            const int kh = kb, kw = a_chunk;
#pragma unroll
            for (int i = 0; i < A_ROWS; i++) {
                int r = a_row0 + i * A_ROW_STEP;
                int iy = rows[i].y0 + kh, ix = rows[i].x0 + kw;
                bool ok = rows[i].base != nullptr && iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
                const bf16 *src = ok ? rows[i].base + ((size_t)iy * p.IW + ix) * 4 : A;
                cp_async8(smem_u32(a_dst + r * LDS + kw * 4), src, ok ? 8 : 0);
            }
        } else {
            const int tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
#pragma unroll
            for (int i = 0; i < A_ROWS; i++) {
                int r = a_row0 + i * A_ROW_STEP;
                int iy, ix; bool ok = rows[i].base != nullptr;
                if (p.mode == 0) {
                    iy = rows[i].y0 + kh; ix = rows[i].x0 + kw;
                } else {
                    int ty = rows[i].y0 - kh, tx = rows[i].x0 - kw;
                    ok = ok && ty >= 0 && tx >= 0;
                    if (p.stride > 1) {
                        iy = ty / p.stride; ix = tx / p.stride;
                        ok = ok && (iy * p.stride == ty) && (ix * p.stride == tx);
                    } else { iy = ty; ix = tx; }
                }
                ok = ok && iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
                const bf16 *src = ok ? rows[i].base + a_off + ((size_t)iy * p.IW + ix) * p.lda + c0 + a_chunk * 8 : A;
                cp_async16(smem_u32(a_dst + r * LDS + a_chunk * 8), src, ok ? 16 : 0);
            }
        }
        {
            const int chunk = tid & 3;
#pragma unroll
            for (int i = 0; i < B_ROWS; i++) {
                int r = (tid >> 2) + i * 64;
                int n = n0 + r;
                bool ok = n < p.N;
                const bf16 *src = ok ? W + w_off + (size_t)n * p.ldw + k0 + chunk * 8 : W;
                cp_async16(smem_u32(b_dst + r * LDS + chunk * 8), src, ok ? 16 : 0);
            }
        }
    };

    float acc[MT][NT][4];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[i][j][k] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    for (int kb = 0; kb < nk; kb++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kb + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const bf16 *a_s = sA + (kb % STAGES) * BM * LDS;
        const bf16 *b_s = sB + (kb % STAGES) * BN * LDS;
#pragma unroll
        for (int kk = 0; kk < BK / 16; kk++) {
            uint32_t af[MT][4], bfr[NT][2];
#pragma unroll
            for (int i = 0; i < MT; i++) {
                int r = wm * WTM + i * 16 + (lane & 15);
                int c = kk * 16 + (lane >> 4) * 8;
                ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3], smem_u32(a_s + r * LDS + c));
            }
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
                int r = wn * WTN + j * 8 + (lane & 7) + ((lane >> 4) << 3);
                int c = kk * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(bfr[j][0], bfr[j][1], bfr[j + 1][0], bfr[j + 1][1], smem_u32(b_s + r * LDS + c));
            }
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) mma_bf16_16816(acc[i][j], af[i], bfr[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue
    const int g = lane >> 2, t = lane & 3;
    const uint32_t thresh = dropout_thresh16(p.drop_p);
    const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    bf16 *C = reinterpret_cast<bf16 *>(p.C);
    const bf16 *R = reinterpret_cast<const bf16 *>(p.residual);
    const bf16 *Mk = reinterpret_cast<const bf16 *>(p.mask);
#pragma unroll
    for (int i = 0; i < MT; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int m = m0 + wm * WTM + i * 16 + g + h * 8;
            if (m >= p.M) continue;
            size_t orow = m;
            if (p.out_stride > 1) {
                int b = m / ohw, rem = m - b * ohw;
                int oy = rem / p.OW, ox = rem - oy * p.OW;
                orow = ((size_t)b * p.SH + (size_t)oy * p.out_stride) * p.SW + (size_t)ox * p.out_stride;
            }
#pragma unroll
            for (int j = 0; j < NT; j++) {
                int n = n0 + wn * WTN + j * 8 + t * 2;
                if (n >= p.N) continue;
                float v0 = acc[i][j][h * 2 + 0], v1 = acc[i][j][h * 2 + 1];
                if (p.bias) { v0 += p.bias[n]; v1 += p.bias[n + 1]; }
                float2 res = make_float2(0.f, 0.f);
                if (R) res = sp_ld2(R + orow * p.ldr + n, p.split);
                // residual joins before the activation (conv blocks) unless dropout is active, in which case it
                // is the un-dropped skip path and joins last (x + dropout(sublayer(x)), transformer.py:169,176)
                if (!(p.drop_p > 0.f)) { v0 += res.x; v1 += res.y; }
                if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                if (Mk) {
                    float2 mk = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(Mk + orow * p.ldm + n));
                    v0 = mk.x > 0.f ? v0 * p.mask_scale : 0.f;
                    v1 = mk.y > 0.f ? v1 * p.mask_scale : 0.f;
                }
                if (p.sigmoid) { v0 = 1.f / (1.f + __expf(-v0)); v1 = 1.f / (1.f + __expf(-v1)); }
                if (p.drop_p > 0.f) {
                    bool k0, k1;
                    dropout_keep2(dropout_bits(seed, p.site, (uint32_t)m, (uint32_t)(n >> 1)), thresh, k0, k1);
                    v0 = (k0 ? v0 * drop_scale : 0.f) + res.x;
                    v1 = (k1 ? v1 * drop_scale : 0.f) + res.y;
                }
                if (C) {
                    bf16 *dst = C + orow * p.ldc + n;
                    if (p.accumulate) { float2 o = sp_ld2(dst, p.split); v0 += o.x; v1 += o.y; }
                    sp_st2(dst, p.split, v0, v1);
                }
                if (p.Cf) {
                    float2 *dst = reinterpret_cast<float2 *>(p.Cf + orow * p.ldcf + n);
                    *dst = make_float2(v0, v1);
                }
            }
        }
    }
}

template <int BN, bool STEM>
int launch(const detrb_igemm_t &p, cudaStream_t stream)
{
    constexpr int smem = STAGES * (BM + BN) * LDS * (int)sizeof(bf16);
    static detrb_per_device_flag configured_dev; bool &configured = configured_dev.slot();      // the opt-in is per device
    if (!configured) {
        DETRB_CUDA(cudaFuncSetAttribute(igemm_kernel<BN, STEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
    DETRB_LAUNCH((igemm_kernel<BN, STEM>), dim3(grid), dim3(NTHREADS), smem, stream, p);
    DETRB_CHECK_LAUNCH("igemm_kernel");
    return DETRB_OK;
}

}  // namespace

extern "C" int detrb_igemm(const detrb_igemm_t *pp, detrb_stream_t stream_)
{
    if (!pp) DETRB_FAIL(DETRB_E_BADARG, "detrb_igemm: null params");
    detrb_igemm_t p = *pp;
    cudaStream_t stream = (cudaStream_t)stream_;
    DETRB_REQUIRE(p.A && p.W && (p.C || p.Cf), "detrb_igemm: null A/W/C");
    DETRB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "detrb_igemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
    DETRB_REQUIRE(p.K % BK == 0, "detrb_igemm: K=%d must be a multiple of %d", p.K, BK);
    DETRB_REQUIRE(p.N % 2 == 0, "detrb_igemm: N=%d must be even", p.N);
    DETRB_REQUIRE(p.M == p.batch * p.OH * p.OW, "detrb_igemm: M=%d != batch*OH*OW=%d", p.M, p.batch * p.OH * p.OW);
    DETRB_REQUIRE(p.stride >= 1 && p.KH >= 1 && p.KW >= 1, "detrb_igemm: bad conv geometry");
    DETRB_REQUIRE(p.ldw >= p.K && p.ldw % 8 == 0, "detrb_igemm: ldw=%d", p.ldw);
    DETRB_REQUIRE(!p.C || p.ldc % 2 == 0, "detrb_igemm: ldc must be even");
    DETRB_REQUIRE(!p.Cf || p.ldcf % 2 == 0, "detrb_igemm: ldcf must be even");
    DETRB_REQUIRE(!p.residual || p.ldr % 2 == 0, "detrb_igemm: ldr must be even");
    DETRB_REQUIRE(!p.mask || p.ldm % 2 == 0, "detrb_igemm: ldm must be even");
    DETRB_REQUIRE(p.drop_p >= 0.f && p.drop_p < 1.f, "detrb_igemm: drop_p");
    DETRB_REQUIRE((p.split != 0) == (p.wsplit != 0) && p.split >= 0 && p.wsplit >= 0 && p.split % 8 == 0 && p.wsplit % 8 == 0,
                  "detrb_igemm: split=%lld wsplit=%lld (both zero, or both positive multiples of 8)", (long long)p.split, (long long)p.wsplit);
    if (p.out_stride < 1) p.out_stride = 1;
    if (p.a_kb_rows) {                                                // sliding-window A: a plain GEMM as far as the tcgen05 kernel is concerned
        DETRB_REQUIRE(p.a_kb_rows > 0 && detrb_gemm_tc_kind(p) == 1, "detrb_igemm: a_kb_rows needs plain geometry and the tcgen05 path");
        return detrb_gemm_tc(p, stream);
    }
    if (detrb_gemm_tc_enabled()) {                                    // tcgen05 / TMA / TMEM
        const int kind = detrb_gemm_tc_kind(p);
        if (kind == 1 || (kind >= 2 && detrb_gemm_tc_conv_enabled())) return detrb_gemm_tc(p, stream);
    }
    if (p.mask_bits || p.out_bits)
        DETRB_FAIL(DETRB_E_SHAPE, "detrb_igemm: 1-bit masks need the tcgen05 path (N %% 64 == 0, 8-byte aligned rows, not together with `mask`)");
    const bool stem = (p.Cin == 4);
    if (stem) {
        DETRB_REQUIRE(!p.split, "detrb_igemm: the Cin=4 stem path has no parity-precision variant");
        DETRB_REQUIRE(p.KW == 8 && p.K == p.KH * 32 && p.lda == 4 && p.mode == 0,
                      "detrb_igemm: stem path needs Cin=4, KW=8, K=KH*32, lda=4, mode=0");
        return p.N >= 128 ? launch<128, true>(p, stream) : launch<64, true>(p, stream);
    }
    DETRB_REQUIRE(p.Cin % BK == 0 && p.K == p.KH * p.KW * p.Cin, "detrb_igemm: Cin=%d K=%d KH=%d KW=%d", p.Cin, p.K, p.KH, p.KW);
    DETRB_REQUIRE(p.lda % 8 == 0 && p.lda >= p.Cin, "detrb_igemm: lda=%d", p.lda);
    return p.N >= 128 ? launch<128, false>(p, stream) : launch<64, false>(p, stream);
}
