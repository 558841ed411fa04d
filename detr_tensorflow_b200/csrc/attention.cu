// attention.cu -- fused multi-head attention core, forward and backward, head_dim = 32.
//   O = dropout(softmax(Q K^T)) V   per (batch, head);  no S x S tensor ever reaches HBM.
// Replaces transformer.py:308-345 (the reshape/transpose x4, matmul QK^T, softmax, dropout, matmul PV)
// and its gradient.  The 32^-0.5 query scaling (transformer.py:307) is applied to the scores (p.scale).
// Tensor path: mma.sync bf16 (flash-attention-2 style online softmax, probabilities stay in registers).
#include "common.cuh"

namespace {

constexpr int DH = 32;
constexpr int TQ = 64;             // queries per CTA (fwd / dQ), keys per CTA (dK/dV)
constexpr int TKV = 64;            // keys (or queries) per inner tile
constexpr int LDH = DH + 8;        // smem row stride (bf16): 80 B
constexpr float LOG2E = 1.4426950408889634f;

// 2^x, one MUFU (exp2f() adds a denormal-range rescale: two FMULs and a compare per element); ex2(-inf) = 0
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// cooperative 64 x 32 bf16 tile load (rows beyond `nrows` zero-filled); 128 threads, 2 chunks each
__device__ __forceinline__ void load_tile64(bf16 *dst, const bf16 *src, int ld, int row0, int nrows, int tid)
{
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int c = tid + i * 128;
        int r = c >> 2, ch = c & 3;
        bool ok = row0 + r < nrows;
        const bf16 *s = ok ? src + (size_t)(row0 + r) * ld + ch * 8 : src;
        cp_async16(smem_u32(dst + r * LDH + ch * 8), s, ok ? 16 : 0);
    }
}

// A-operand fragments (16 rows x 32 dh) of warp rows [r0, r0+16) from a [rows][LDH] tile
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], const bf16 *tile, int r0, int lane)
{
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        int r = r0 + (lane & 15), c = kk * 16 + (lane >> 4) * 8;
        ldmatrix_x4(a[kk][0], a[kk][1], a[kk][2], a[kk][3], smem_u32(tile + r * LDH + c));
    }
}

// S[16 x 64] = A(16 x 32) * T^T where T is a [64][LDH] tile (rows = n index, cols = dh)
__device__ __forceinline__ void mma_a_tileT(float (&s)[8][4], const uint32_t (&a)[2][4], const bf16 *tile, int lane)
{
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) s[j][k] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; kk++)
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            uint32_t b[2][2];
            int r = j * 8 + (lane & 7) + ((lane >> 4) << 3);
            int c = kk * 16 + ((lane >> 3) & 1) * 8;
            ldmatrix_x4(b[0][0], b[0][1], b[1][0], b[1][1], smem_u32(tile + r * LDH + c));
            mma_bf16_16816(s[j], a[kk], b[0]);
            mma_bf16_16816(s[j + 1], a[kk], b[1]);
        }
}

// P (16 x 64, fp32 registers in S-layout) -> bf16 A-operand fragments: pa[kk] = {(j=2kk, row g), (2kk, g+8), (2kk+1, g), (2kk+1, g+8)}
__device__ __forceinline__ void pack_p(uint32_t (&pa)[4][4], const float (&p)[8][4])
{
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        pa[kk][0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
        pa[kk][1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
        pa[kk][2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        pa[kk][3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    }
}
// O[16 x 32] += P(16 x 64, packed fragments) * T where T is a [64][LDH] tile (rows = k index, cols = dh)
__device__ __forceinline__ void mma_pa_tile(float (&o)[4][4], const uint32_t (&pa)[4][4], const bf16 *tile, int lane)
{
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
            uint32_t b[2][2];
            int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            int c = j * 8 + (lane >> 4) * 8;
            ldmatrix_x4_trans(b[0][0], b[0][1], b[1][0], b[1][1], smem_u32(tile + r * LDH + c));
            mma_bf16_16816(o[j], pa[kk], b[0]);
            mma_bf16_16816(o[j + 1], pa[kk], b[1]);
        }
    }
}
__device__ __forceinline__ void mma_p_tile(float (&o)[4][4], const float (&p)[8][4], const bf16 *tile, int lane)
{
    uint32_t pa[4][4];
    pack_p(pa, p);
    mma_pa_tile(o, pa, tile, lane);
}
// Dropout keep bits of one thread's scores in the forward / dQ layout (rows g and g+8 of a warp's 16 query rows; columns
// kbase + 8j + 2t + {0,1}, j = 0..7): kb[r][m][c] covers column 16m + 2t + c in its low field (j = 2m) and column
// 16m + 8 + 2t + c in its high field (j = 2m + 1) -- see attn_drop_word (common.cuh).
__device__ __forceinline__ void attn_keepbits_rowmajor(uint32_t (&kb)[2][4][2], const uint32_t (&rh)[2], int kbase, int t, uint32_t thresh2)
{
    const uint32_t p0 = (uint32_t)(kbase >> 4) * 8u + (uint32_t)t * 2u;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int c = 0; c < 2; c++) kb[r][m][c] = attn_drop_keepbits(attn_drop_word(rh[r], p0 + 8u * m + c), thresh2);
}

// --------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const detrb_attn_fwd_t p)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    __shared__ __align__(16) bf16 sQ[TQ * LDH];
    __shared__ __align__(16) bf16 sK[2][TKV * LDH];
    __shared__ __align__(16) bf16 sV[2][TKV * LDH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const bf16 *Q = reinterpret_cast<const bf16 *>(p.Q) + (size_t)b * p.Lq * p.ldq + h * DH;
    const bf16 *K = reinterpret_cast<const bf16 *>(p.K) + (size_t)b * p.Lk * p.ldk + h * DH;
    const bf16 *V = reinterpret_cast<const bf16 *>(p.V) + (size_t)b * p.Lk * p.ldv + h * DH;
    const int nkt = (p.Lk + TKV - 1) / TKV;

    load_tile64(sQ, Q, p.ldq, q0, p.Lq, tid);
    load_tile64(sK[0], K, p.ldk, 0, p.Lk, tid);
    load_tile64(sV[0], V, p.ldv, 0, p.Lk, tid);
    cp_async_commit();

    float o[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) o[j][k] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    uint32_t aq[2][4];

    const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
    const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint32_t rowbase = (uint32_t)((b * p.H + h) * p.Lq + q0 + warp * 16 + g);
    const float sl2 = p.scale * LOG2E;
    const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    const uint32_t rh[2] = {attn_drop_rowhash(seed, p.site, rowbase), attn_drop_rowhash(seed, p.site, rowbase + 8)};

    for (int kt = 0; kt < nkt; kt++) {
        if (kt + 1 < nkt) {
            load_tile64(sK[(kt + 1) & 1], K, p.ldk, (kt + 1) * TKV, p.Lk, tid);
            load_tile64(sV[(kt + 1) & 1], V, p.ldv, (kt + 1) * TKV, p.Lk, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) load_a_frags(aq, sQ, warp * 16, lane);

        float s[8][4];
        mma_a_tileT(s, aq, sK[kt & 1], lane);
        // row maxima on the raw scores (sl2 > 0: max commutes with the scaling); keys beyond Lk exist only in the last tile
        const int kbase = kt * TKV;
        if (kbase + TKV > p.Lk) {
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (kbase + j * 8 + t * 2 + (e & 1) >= p.Lk) s[j][e] = -INFINITY;
        }
        float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) tmax[e >> 1] = fmaxf(tmax[e >> 1], s[j][e]);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
        }
        float alpha[2], mneg[2], rsum[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const float mnew = fmaxf(mrow[r], tmax[r] * sl2);  // log2 domain
            alpha[r] = ex2(mrow[r] - mnew);                    // ex2(-inf) = 0 on the first tile
            mrow[r] = mnew;
            mneg[r] = -mnew;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float pv = ex2(fmaf(s[j][e], sl2, mneg[e >> 1]));
                rsum[e >> 1] += pv;
                s[j][e] = pv;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            rsum[r] += __shfl_xor_sync(0xffffffffu, rsum[r], 1);
            rsum[r] += __shfl_xor_sync(0xffffffffu, rsum[r], 2);
            lrow[r] = lrow[r] * alpha[r] + rsum[r];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
            o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
        }
        uint32_t pa[4][4];
        pack_p(pa, s);
        if (p.drop_p > 0.f) {                                  // the 1/(1-p) rescale is folded into the final normalisation
            uint32_t kb[2][4][2];
            attn_keepbits_rowmajor(kb, rh, kbase, t, thresh2);
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
#pragma unroll
                for (int r = 0; r < 2; r++) {                  // packed pair = columns (2t, 2t+1): words c = 0 / 1
                    pa[kk][r] &= attn_drop_mask2_lo(kb[r][kk][0], kb[r][kk][1]);          // j = 2kk
                    pa[kk][2 + r] &= attn_drop_mask2_hi(kb[r][kk][0], kb[r][kk][1]);      // j = 2kk + 1
                }
        }
        mma_pa_tile(o, pa, sV[kt & 1], lane);
        __syncthreads();
    }

    bf16 *O = reinterpret_cast<bf16 *>(p.O) + (size_t)b * p.Lq * p.ldo + h * DH;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int q = q0 + warp * 16 + g + r * 8;
        if (q >= p.Lq) continue;
        float inv = drop_scale / lrow[r];
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<uint32_t *>(O + (size_t)q * p.ldo + j * 8 + t * 2) =
                pack_bf16x2(o[j][r * 2] * inv, o[j][r * 2 + 1] * inv);
        if (t == 0 && p.lse)
            p.lse[((size_t)b * p.H + h) * p.Lq + q] = (mrow[r] + log2f(lrow[r])) * (1.f / LOG2E);
    }
}

// --------------------------------------------------------------------------------------- backward
// delta[b,h,q] = sum_d dO[b,q,h*32+d] * O[b,q,h*32+d]
__global__ void attn_delta_kernel(const bf16 *O, const bf16 *dO, int ldo, int lddo, float *delta,
                                  int B, int H, int Lq, long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H * Lq) return;
    int q = idx % Lq, h = (idx / Lq) % H, b = idx / (Lq * H);
    const bf16 *o = O + ((size_t)b * Lq + q) * ldo + h * DH;
    const bf16 *d = dO + ((size_t)b * Lq + q) * lddo + h * DH;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float ov[8], dv[8];
        sp_ld8(o + i * 8, split, ov);
        sp_ld8(d + i * 8, split, dv);
#pragma unroll
        for (int j = 0; j < 8; j += 2) acc += ov[j] * dv[j] + ov[j + 1] * dv[j + 1];
    }
    delta[idx] = acc;
}

// --------------------------------------------------------------------------------------- parity precision (p.split != 0)
// The reference computes attention in fp32 (transformer.py:308-345).  These kernels do the same arithmetic in fp32 on the bf16
// pairs: one warp per (batch, head, row), lane = head channel (head_dim == 32 == warp size), a warp reduction per score.  They
// exist for the fp32-tolerance parity tests (small batches); the throughput path is the tensor-core kernels above.  Dropout
// uses the same counter-based masks as the tensor-core kernels.
__device__ __forceinline__ bool attn_keep(uint32_t rowhash_m, uint32_t k, uint32_t thresh2) {
    const uint32_t kb = attn_drop_keepbits(attn_drop_word(rowhash_m, attn_drop_pair(k)), thresh2);
    return ((((k >> 3) & 1u) ? attn_drop_mask_hi(kb) : attn_drop_mask_lo(kb)) & 1u) != 0u;
}

__global__ void __launch_bounds__(256)
attn_fwd_sp_kernel(const detrb_attn_fwd_t p)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);          // (b*H + h)*Lq + q
    if (row >= (long long)p.B * p.H * p.Lq) return;
    const int q = (int)(row % p.Lq), h = (int)((row / p.Lq) % p.H), b = (int)(row / ((long long)p.Lq * p.H));
    const long long sp = p.split;
    const bf16 *K = reinterpret_cast<const bf16 *>(p.K) + (size_t)b * p.Lk * p.ldk + h * DH + lane;
    const bf16 *V = reinterpret_cast<const bf16 *>(p.V) + (size_t)b * p.Lk * p.ldv + h * DH + lane;
    const float qv = sp_ld1(reinterpret_cast<const bf16 *>(p.Q) + ((size_t)b * p.Lq + q) * p.ldq + h * DH + lane, sp) * p.scale;
    const bool drop = p.drop_p > 0.f;
    const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
    const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
    const uint32_t rh = attn_drop_rowhash(seed, p.site, (uint32_t)row);
    float m = -INFINITY, l = 0.f, acc = 0.f;
    for (int k = 0; k < p.Lk; k++) {
        const float s = warp_sum(qv * sp_ld1(K + (size_t)k * p.ldk, sp));
        const float mn = fmaxf(m, s);
        const float alpha = expf(m - mn), pr = expf(s - mn);
        l = l * alpha + pr;
        const float pd = (!drop || attn_keep(rh, (uint32_t)k, thresh2)) ? pr : 0.f;
        acc = acc * alpha + pd * sp_ld1(V + (size_t)k * p.ldv, sp);
        m = mn;
    }
    const float drop_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    sp_st1(reinterpret_cast<bf16 *>(p.O) + ((size_t)b * p.Lq + q) * p.ldo + h * DH + lane, sp, acc * drop_scale / l);
    if (lane == 0 && p.lse) p.lse[row] = m + logf(l);
}

// dQ[q] = scale * sum_k dS[q,k] K[k],  dS = P * (dP - delta),  dP = keep/(1-p) * (dO . V[k])
__global__ void __launch_bounds__(256)
attn_bwd_dq_sp_kernel(const detrb_attn_bwd_t p)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)p.B * p.H * p.Lq) return;
    const int q = (int)(row % p.Lq), h = (int)((row / p.Lq) % p.H), b = (int)(row / ((long long)p.Lq * p.H));
    const long long sp = p.split;
    const bf16 *K = reinterpret_cast<const bf16 *>(p.K) + (size_t)b * p.Lk * p.ldk + h * DH + lane;
    const bf16 *V = reinterpret_cast<const bf16 *>(p.V) + (size_t)b * p.Lk * p.ldv + h * DH + lane;
    const float qv = sp_ld1(reinterpret_cast<const bf16 *>(p.Q) + ((size_t)b * p.Lq + q) * p.ldq + h * DH + lane, sp) * p.scale;
    const float dov = sp_ld1(reinterpret_cast<const bf16 *>(p.dO) + ((size_t)b * p.Lq + q) * p.lddo + h * DH + lane, sp);
    const float lse = p.lse[row], delta = p.delta[row];
    const bool drop = p.drop_p > 0.f;
    const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
    const float drop_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
    const uint32_t rh = attn_drop_rowhash(seed, p.site, (uint32_t)row);
    float acc = 0.f;
    for (int k = 0; k < p.Lk; k++) {
        const float kv = sp_ld1(K + (size_t)k * p.ldk, sp);
        const float s = warp_sum(qv * kv);
        float dp = warp_sum(dov * sp_ld1(V + (size_t)k * p.ldv, sp));
        dp = (!drop || attn_keep(rh, (uint32_t)k, thresh2)) ? dp * drop_scale : 0.f;
        acc += expf(s - lse) * (dp - delta) * kv;
    }
    sp_st1(reinterpret_cast<bf16 *>(p.dQ) + ((size_t)b * p.Lq + q) * p.lddq + h * DH + lane, sp, acc * p.scale);
}

// dK[k] = scale * sum_q dS[q,k] Q[q],  dV[k] = sum_q keep/(1-p) P[q,k] dO[q]
__global__ void __launch_bounds__(256)
attn_bwd_dkv_sp_kernel(const detrb_attn_bwd_t p)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);          // (b*H + h)*Lk + key
    if (row >= (long long)p.B * p.H * p.Lk) return;
    const int key = (int)(row % p.Lk), h = (int)((row / p.Lk) % p.H), b = (int)(row / ((long long)p.Lk * p.H));
    const long long sp = p.split;
    const bf16 *Q = reinterpret_cast<const bf16 *>(p.Q) + (size_t)b * p.Lq * p.ldq + h * DH + lane;
    const bf16 *dO = reinterpret_cast<const bf16 *>(p.dO) + (size_t)b * p.Lq * p.lddo + h * DH + lane;
    const float kv = sp_ld1(reinterpret_cast<const bf16 *>(p.K) + ((size_t)b * p.Lk + key) * p.ldk + h * DH + lane, sp) * p.scale;
    const float vv = sp_ld1(reinterpret_cast<const bf16 *>(p.V) + ((size_t)b * p.Lk + key) * p.ldv + h * DH + lane, sp);
    const float *lse = p.lse + ((size_t)b * p.H + h) * p.Lq, *delta = p.delta + ((size_t)b * p.H + h) * p.Lq;
    const bool drop = p.drop_p > 0.f;
    const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
    const float drop_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint64_t seed = p.seed ^ ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
    float dk = 0.f, dv = 0.f;
    for (int q = 0; q < p.Lq; q++) {
        const float qv = sp_ld1(Q + (size_t)q * p.ldq, sp), dov = sp_ld1(dO + (size_t)q * p.lddo, sp);
        const float s = warp_sum(qv * kv);
        const float pr = expf(s - lse[q]);
        float dp = warp_sum(dov * vv);
        const bool keep = !drop || attn_keep(attn_drop_rowhash(seed, p.site, (uint32_t)(((size_t)b * p.H + h) * p.Lq + q)), (uint32_t)key, thresh2);
        dp = keep ? dp * drop_scale : 0.f;
        dv += (keep ? pr * drop_scale : 0.f) * dov;
        dk += pr * (dp - delta[q]) * qv;
    }
    sp_st1(reinterpret_cast<bf16 *>(p.dK) + ((size_t)b * p.Lk + key) * p.lddk + h * DH + lane, sp, dk * p.scale);
    sp_st1(reinterpret_cast<bf16 *>(p.dV) + ((size_t)b * p.Lk + key) * p.lddv + h * DH + lane, sp, dv);
}

// dK, dV: CTA owns 64 keys (warp: 16), loops over query tiles; works on S^T so that P^T / dS^T come out
// of the MMAs directly in A-operand layout.
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const detrb_attn_bwd_t p)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    __shared__ __align__(16) bf16 sK[TQ * LDH];
    __shared__ __align__(16) bf16 sV[TQ * LDH];
    __shared__ __align__(16) bf16 sQ[2][TKV * LDH];
    __shared__ __align__(16) bf16 sdO[2][TKV * LDH];
    __shared__ float sLse[2][TKV], sDelta[2][TKV];
    __shared__ uint32_t sRh[2][TKV];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const bf16 *Q = reinterpret_cast<const bf16 *>(p.Q) + (size_t)b * p.Lq * p.ldq + h * DH;
    const bf16 *K = reinterpret_cast<const bf16 *>(p.K) + (size_t)b * p.Lk * p.ldk + h * DH;
    const bf16 *V = reinterpret_cast<const bf16 *>(p.V) + (size_t)b * p.Lk * p.ldv + h * DH;
    const bf16 *dO = reinterpret_cast<const bf16 *>(p.dO) + (size_t)b * p.Lq * p.lddo + h * DH;
    const float *lse = p.lse + ((size_t)b * p.H + h) * p.Lq;
    const float *delta = p.delta + ((size_t)b * p.H + h) * p.Lq;
    const int nqt = (p.Lq + TKV - 1) / TKV;
    const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);

    auto load_q = [&](int st, int qt) {
        load_tile64(sQ[st], Q, p.ldq, qt * TKV, p.Lq, tid);
        load_tile64(sdO[st], dO, p.lddo, qt * TKV, p.Lq, tid);
        if (tid < TKV) {
            int q = qt * TKV + tid;
            sLse[st][tid] = q < p.Lq ? -lse[q] * LOG2E : -INFINITY;      // negated, log2 domain
            sDelta[st][tid] = q < p.Lq ? -delta[q] * p.scale : 0.f;       // -scale * delta
            sRh[st][tid] = attn_drop_rowhash(seed, p.site, (uint32_t)((b * p.H + h) * p.Lq + q));
        }
    };
    load_tile64(sK, K, p.ldk, k0, p.Lk, tid);
    load_tile64(sV, V, p.ldv, k0, p.Lk, tid);
    load_q(0, 0);
    cp_async_commit();

    float dk[4][4], dv[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) { dk[j][k] = 0.f; dv[j][k] = 0.f; }
    uint32_t ak[2][4], av[2][4];
    const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
    const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const int keyr[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
    // this thread's two key rows (g, g + 8 of a 16-key group) are the low / high field of ONE dropout word per query
    const uint32_t kpair = attn_drop_pair((uint32_t)keyr[0]);
    const float sl2 = p.scale * LOG2E, dss = p.scale * drop_scale;

    for (int qt = 0; qt < nqt; qt++) {
        if (qt + 1 < nqt) { load_q((qt + 1) & 1, qt + 1); cp_async_commit(); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        if (qt == 0) { load_a_frags(ak, sK, warp * 16, lane); load_a_frags(av, sV, warp * 16, lane); }
        const int st = qt & 1;
        float sT[8][4], dpT[8][4];
        mma_a_tileT(sT, ak, sQ[st], lane);        // S^T[key, q]
        mma_a_tileT(dpT, av, sdO[st], lane);      // dP^T[key, q]
        // rows of S^T are keys: rows beyond Lk only feed their own (never stored) dK / dV rows, so no key mask is needed;
        // queries beyond Lq carry lse = +inf -> p = 0.  The dropout rescale 1/(1-p) is folded into dss and the final dV.
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int ql = j * 8 + t * 2 + c;
                uint32_t mk[2] = {0xffffffffu, 0xffffffffu};
                if (p.drop_p > 0.f) {
                    const uint32_t kb = attn_drop_keepbits(attn_drop_word(sRh[st][ql], kpair), thresh2);
                    mk[0] = attn_drop_mask_lo(kb);                         // key row g
                    mk[1] = attn_drop_mask_hi(kb);                         // key row g + 8
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int e = r * 2 + c;
                    const float pv = ex2(fmaf(sT[j][e], sl2, sLse[st][ql]));
                    const float dpv = __uint_as_float(__float_as_uint(dpT[j][e]) & mk[r]);
                    sT[j][e] = __uint_as_float(__float_as_uint(pv) & mk[r]);   // dropped probabilities (for dV), unscaled
                    dpT[j][e] = pv * fmaf(dpv, dss, sDelta[st][ql]);           // scale * dS^T
                }
            }
        mma_p_tile(dv, sT, sdO[st], lane);        // dV += P_d^T dO
        mma_p_tile(dk, dpT, sQ[st], lane);        // dK += dS^T Q
        __syncthreads();
    }
    bf16 *dK = reinterpret_cast<bf16 *>(p.dK) + (size_t)b * p.Lk * p.lddk + h * DH;
    bf16 *dV = reinterpret_cast<bf16 *>(p.dV) + (size_t)b * p.Lk * p.lddv + h * DH;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int key = keyr[r];
        if (key >= p.Lk) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            *reinterpret_cast<uint32_t *>(dK + (size_t)key * p.lddk + j * 8 + t * 2) = pack_bf16x2(dk[j][r * 2], dk[j][r * 2 + 1]);
            *reinterpret_cast<uint32_t *>(dV + (size_t)key * p.lddv + j * 8 + t * 2) = pack_bf16x2(dv[j][r * 2] * drop_scale, dv[j][r * 2 + 1] * drop_scale);
        }
    }
}

// dQ: CTA owns 64 queries (warp: 16), loops over key tiles.
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const detrb_attn_bwd_t p)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    __shared__ __align__(16) bf16 sQ[TQ * LDH];
    __shared__ __align__(16) bf16 sdO[TQ * LDH];
    __shared__ __align__(16) bf16 sK[2][TKV * LDH];
    __shared__ __align__(16) bf16 sV[2][TKV * LDH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const bf16 *Q = reinterpret_cast<const bf16 *>(p.Q) + (size_t)b * p.Lq * p.ldq + h * DH;
    const bf16 *K = reinterpret_cast<const bf16 *>(p.K) + (size_t)b * p.Lk * p.ldk + h * DH;
    const bf16 *V = reinterpret_cast<const bf16 *>(p.V) + (size_t)b * p.Lk * p.ldv + h * DH;
    const bf16 *dO = reinterpret_cast<const bf16 *>(p.dO) + (size_t)b * p.Lq * p.lddo + h * DH;
    const int nkt = (p.Lk + TKV - 1) / TKV;

    load_tile64(sQ, Q, p.ldq, q0, p.Lq, tid);
    load_tile64(sdO, dO, p.lddo, q0, p.Lq, tid);
    load_tile64(sK[0], K, p.ldk, 0, p.Lk, tid);
    load_tile64(sV[0], V, p.ldv, 0, p.Lk, tid);
    cp_async_commit();

    float dq[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) dq[j][k] = 0.f;
    uint32_t aq[2][4], ado[2][4];
    float lrow[2], drow[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int q = q0 + warp * 16 + g + r * 8;
        size_t idx = ((size_t)b * p.H + h) * p.Lq + q;
        lrow[r] = q < p.Lq ? -p.lse[idx] * LOG2E : -INFINITY;         // negated, log2 domain
        drow[r] = q < p.Lq ? -p.delta[idx] * p.scale : 0.f;          // -scale * delta
    }
    const uint32_t thresh2 = attn_drop_thresh2(p.drop_p);
    const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint32_t rowbase = (uint32_t)((b * p.H + h) * p.Lq + q0 + warp * 16 + g);
    const float sl2 = p.scale * LOG2E, dss = p.scale * drop_scale;
    const uint64_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    const uint32_t rh[2] = {attn_drop_rowhash(seed, p.site, rowbase), attn_drop_rowhash(seed, p.site, rowbase + 8)};

    for (int kt = 0; kt < nkt; kt++) {
        if (kt + 1 < nkt) {
            load_tile64(sK[(kt + 1) & 1], K, p.ldk, (kt + 1) * TKV, p.Lk, tid);
            load_tile64(sV[(kt + 1) & 1], V, p.ldv, (kt + 1) * TKV, p.Lk, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else cp_async_wait<0>();
        __syncthreads();
        if (kt == 0) { load_a_frags(aq, sQ, warp * 16, lane); load_a_frags(ado, sdO, warp * 16, lane); }
        float s[8][4], dp[8][4];
        mma_a_tileT(s, aq, sK[kt & 1], lane);
        mma_a_tileT(dp, ado, sV[kt & 1], lane);
        const int kbase = kt * TKV;
        if (p.drop_p > 0.f) {                                   // dP of dropped probabilities is zero
            uint32_t kb[2][4][2];
            attn_keepbits_rowmajor(kb, rh, kbase, t, thresh2);
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const uint32_t w = kb[e >> 1][j >> 1][e & 1];
                    const uint32_t mk = (j & 1) ? attn_drop_mask_hi(w) : attn_drop_mask_lo(w);
                    dp[j][e] = __uint_as_float(__float_as_uint(dp[j][e]) & mk);
                }
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float pv = ex2(fmaf(s[j][e], sl2, lrow[e >> 1]));
                s[j][e] = pv * fmaf(dp[j][e], dss, drow[e >> 1]);    // scale * dS
            }
        // keys beyond Lk (last tile only): their zero-filled K rows give s = 0, p = exp(-lse), which overflows for rows whose
        // scores are all very negative -- and inf * 0 would poison dQ
        if (kbase + TKV > p.Lk) {
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (kbase + j * 8 + t * 2 + (e & 1) >= p.Lk) s[j][e] = 0.f;
        }
        mma_p_tile(dq, s, sK[kt & 1], lane);      // dQ += dS K
        __syncthreads();
    }
    bf16 *dQ = reinterpret_cast<bf16 *>(p.dQ) + (size_t)b * p.Lq * p.lddq + h * DH;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int q = q0 + warp * 16 + g + r * 8;
        if (q >= p.Lq) continue;
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<uint32_t *>(dQ + (size_t)q * p.lddq + j * 8 + t * 2) = pack_bf16x2(dq[j][r * 2], dq[j][r * 2 + 1]);
    }
}

}  // namespace

extern "C" int detrb_attn_fwd(const detrb_attn_fwd_t *pp, detrb_stream_t stream_)
{
    if (!pp) DETRB_FAIL(DETRB_E_BADARG, "detrb_attn_fwd: null params");
    const detrb_attn_fwd_t &p = *pp;
    DETRB_REQUIRE(p.Q && p.K && p.V && p.O, "detrb_attn_fwd: null pointer");
    DETRB_REQUIRE(p.B > 0 && p.H > 0 && p.Lq > 0 && p.Lk > 0, "detrb_attn_fwd: empty problem");
    DETRB_REQUIRE(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 2 == 0, "detrb_attn_fwd: strides must be multiples of 8");
    DETRB_REQUIRE(p.drop_p >= 0.f && p.drop_p < 1.f, "detrb_attn_fwd: drop_p");
    if (p.split) {
        const long long rows = (long long)p.B * p.H * p.Lq;
        DETRB_LAUNCH(attn_fwd_sp_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream_, p);
        DETRB_CHECK_LAUNCH("attn_fwd_sp_kernel");
        return DETRB_OK;
    }
    if (detrb_attn_tc_enabled() && detrb_attn_fwd_tc_supported(p)) return detrb_attn_fwd_tc(p, (cudaStream_t)stream_);   // tcgen05 / TMA / TMEM
    dim3 grid(ceil_div(p.Lq, TQ), p.H, p.B);
    DETRB_LAUNCH(attn_fwd_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream_, p);
    DETRB_CHECK_LAUNCH("attn_fwd_kernel");
    return DETRB_OK;
}

extern "C" int detrb_attn_bwd(const detrb_attn_bwd_t *pp, detrb_stream_t stream_)
{
    if (!pp) DETRB_FAIL(DETRB_E_BADARG, "detrb_attn_bwd: null params");
    const detrb_attn_bwd_t &p = *pp;
    cudaStream_t stream = (cudaStream_t)stream_;
    DETRB_REQUIRE(p.Q && p.K && p.V && p.O && p.dO && p.lse && p.delta && p.dQ && p.dK && p.dV, "detrb_attn_bwd: null pointer");
    DETRB_REQUIRE(p.B > 0 && p.H > 0 && p.Lq > 0 && p.Lk > 0, "detrb_attn_bwd: empty problem");
    DETRB_REQUIRE(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 8 == 0 && p.lddo % 8 == 0,
                  "detrb_attn_bwd: strides must be multiples of 8");
    DETRB_REQUIRE(p.parts >= 0 && p.parts <= 7, "detrb_attn_bwd: parts is a mask of 1 (delta) | 2 (dK/dV) | 4 (dQ)");
    const int parts = p.parts ? p.parts : 7;
    int n = p.B * p.H * p.Lq;
    if (parts & 1) {
        DETRB_LAUNCH(attn_delta_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, stream, reinterpret_cast<const bf16 *>(p.O), reinterpret_cast<const bf16 *>(p.dO),
                                                                p.ldo, p.lddo, p.delta, p.B, p.H, p.Lq, (long long)p.split);
        DETRB_CHECK_LAUNCH("attn_delta_kernel");
    }
    if (p.split) {
        const long long rq = (long long)p.B * p.H * p.Lq, rk = (long long)p.B * p.H * p.Lk;
        if (parts & 2) {
            DETRB_LAUNCH(attn_bwd_dkv_sp_kernel, dim3((unsigned)((rk + 7) / 8)), dim3(256), 0, stream, p);
            DETRB_CHECK_LAUNCH("attn_bwd_dkv_sp_kernel");
        }
        if (parts & 4) {
            DETRB_LAUNCH(attn_bwd_dq_sp_kernel, dim3((unsigned)((rq + 7) / 8)), dim3(256), 0, stream, p);
            DETRB_CHECK_LAUNCH("attn_bwd_dq_sp_kernel");
        }
        return DETRB_OK;
    }
    if (!(parts & 6)) return DETRB_OK;
    if (detrb_attn_tc_enabled() && detrb_attn_bwd_tc_supported(p)) return detrb_attn_bwd_tc(p, stream);            // tcgen05 / TMA / TMEM
    if (parts & 2) {
        DETRB_LAUNCH(attn_bwd_dkv_kernel, dim3(dim3(ceil_div(p.Lk, TQ), p.H, p.B)), dim3(128), 0, stream, p);
        DETRB_CHECK_LAUNCH("attn_bwd_dkv_kernel");
    }
    if (parts & 4) {
        DETRB_LAUNCH(attn_bwd_dq_kernel, dim3(dim3(ceil_div(p.Lq, TQ), p.H, p.B)), dim3(128), 0, stream, p);
        DETRB_CHECK_LAUNCH("attn_bwd_dq_kernel");
    }
    return DETRB_OK;
}
