// elementwise.cu -- bandwidth-bound pieces of the DETR step: LayerNorm fwd/bwd (d = 256), positional
// add, max-pool fwd/bwd, layout conversion, column sums (bias gradients).  All NHWC / row-major bf16,
// 16-byte vector accesses, one warp per row where a row reduction is needed.
#include "common.cuh"

namespace {

constexpr int D = 256;   // model_dim (transformer.py:8)

__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}

// ---------------------------------------------------------------- LayerNorm forward: warp per row
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const bf16 *x, const float *gamma, const float *beta, bf16 *y, bf16 *y2, const bf16 *pos, int S,
              float *mean, float *rstd, int M, long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= M) return;
    float v[8];
    sp_ld8(x + (size_t)row * D + lane * 8, split, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    float mu = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { float d = v[i] - mu; q += d * d; }
    float rs = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
    float4 g0 = *reinterpret_cast<const float4 *>(gamma + lane * 8), g1 = *reinterpret_cast<const float4 *>(gamma + lane * 8 + 4);
    float4 b0 = *reinterpret_cast<const float4 *>(beta + lane * 8), b1 = *reinterpret_cast<const float4 *>(beta + lane * 8 + 4);
    float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = (v[i] - mu) * rs * gg[i] + bb[i];
    sp_st8(y + (size_t)row * D + lane * 8, split, o);
    if (y2) {
        // y2 = stored(y) + pos : add to the ROUNDED y so that y2 == (stored y) + pos exactly like a separate add
        float pp[8];
        sp_round8(split, o);
        sp_ld8(pos + (size_t)(row % S) * D + lane * 8, split, pp);
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] += pp[i];
        sp_st8(y2 + (size_t)row * D + lane * 8, split, o);
    }
    if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
}

// ---------------------------------------------------------------- LayerNorm backward
// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma ; dgamma += sum dy*xhat ; dbeta += sum dy
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const bf16 *dy, const bf16 *dy2, const bf16 *x, const float *gamma, const float *mean, const float *rstd,
              bf16 *dx, bf16 *dx_drop, float drop_p, uint64_t seed_in, uint32_t site, const uint64_t *seed_ptr,
              float *dgamma, float *dbeta, int M, int rows_per_block, long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const uint64_t seed = seed_in ^ ((drop_p > 0.f && seed_ptr) ? *seed_ptr : 0ull);
    __shared__ float sg[8][D], sb[8][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 g0 = *reinterpret_cast<const float4 *>(gamma + lane * 8), g1 = *reinterpret_cast<const float4 *>(gamma + lane * 8 + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float acc_g[8], acc_b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc_g[i] = 0.f; acc_b[i] = 0.f; }
    const uint32_t thresh = dropout_thresh16(drop_p);
    const float drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(M, r_begin + rows_per_block);
    for (int row = r_begin + warp; row < r_end; row += 8) {
        float d[8], xv[8];
        sp_ld8(dy + (size_t)row * D + lane * 8, split, d);
        if (dy2) {
            float d2[8];
            sp_ld8(dy2 + (size_t)row * D + lane * 8, split, d2);
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] += d2[i];
        }
        sp_ld8(x + (size_t)row * D + lane * 8, split, xv);
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f, xh[8], gv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            xh[i] = (xv[i] - mu) * rs;
            gv[i] = d[i] * gg[i];
            s1 += gv[i]; s2 += gv[i] * xh[i];
            acc_g[i] += d[i] * xh[i]; acc_b[i] += d[i];
        }
        s1 = warp_sum(s1) * (1.f / D); s2 = warp_sum(s2) * (1.f / D);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = rs * (gv[i] - s1 - xh[i] * s2);
        sp_st8(dx + (size_t)row * D + lane * 8, split, o);
        if (dx_drop) {
            float od[8];
#pragma unroll
            for (int i = 0; i < 8; i++) od[i] = o[i];
            sp_round8(split, od);
            if (drop_p > 0.f) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    bool k0, k1;
                    dropout_keep2(dropout_bits(seed, site, (uint32_t)row, (uint32_t)((lane * 8 + i) >> 1)), thresh, k0, k1);
                    od[i] = k0 ? od[i] * drop_scale : 0.f;
                    od[i + 1] = k1 ? od[i + 1] * drop_scale : 0.f;
                }
            }
            sp_st8(dx_drop + (size_t)row * D + lane * 8, split, od);
        }
    }
    if (dgamma) {
#pragma unroll
        for (int i = 0; i < 8; i++) { sg[warp][lane * 8 + i] = acc_g[i]; sb[warp][lane * 8 + i] = acc_b[i]; }
        __syncthreads();
        int c = threadIdx.x;       // 256 threads == D columns
        float a = 0.f, bsum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) { a += sg[w][c]; bsum += sb[w][c]; }
        atomicAdd(dgamma + c, a);
        atomicAdd(dbeta + c, bsum);
    }
}

// ---------------------------------------------------------------- stride-2 scatter of densely written GEMM results
// dst[b, y, x, :] (op)= mask(src_class(y & 1, x & 1)[b, y >> 1, x >> 1, :]); the four parity classes are compact [B, A_c, B_c, C] tensors
// (A_c = (H - py + 1) / 2, B_c = (W - px + 1) / 2), a null class leaves its pixels untouched.  One thread per 8 channels of one pixel.
__global__ void __launch_bounds__(256)
scatter_s2_kernel(const bf16 *s0, const bf16 *s1, const bf16 *s2, const bf16 *s3, bf16 *dst, int ldc, const uint8_t *mask_bits, int ldmb,
                  int H, int W, int C8, int accumulate)
{
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * C8) return;
    const int x = t / C8, c8 = t - x * C8, y = blockIdx.y, b = blockIdx.z;
    const int py = y & 1, px = x & 1;
    const bf16 *src = py ? (px ? s3 : s2) : (px ? s1 : s0);
    if (!src) return;
    const int Ac = (H - py + 1) >> 1, Bc = (W - px + 1) >> 1;
    const size_t spix = ((size_t)b * Ac + (y >> 1)) * Bc + (x >> 1), dpix = ((size_t)b * H + y) * W + x;
    uint4 u = *reinterpret_cast<const uint4 *>(src + spix * (size_t)(C8 * 8) + c8 * 8);
    float v[8];
    sp_unpack8(u, v);
    bf16 *d = dst + dpix * ldc + c8 * 8;
    if (mask_bits) {
        const uint32_t m = mask_bits[dpix * ldmb + c8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = ((m >> i) & 1u) ? v[i] : 0.f;
    }
    if (accumulate) {
        float o[8];
        sp_unpack8(*reinterpret_cast<const uint4 *>(d), o);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] += o[i];
    }
    *reinterpret_cast<uint4 *>(d) = sp_pack8(v);
}

// ---------------------------------------------------------------- simple vector kernels
__global__ void add_rowbcast_kernel(const bf16 *x, const bf16 *pos, bf16 *out, int64_t nvec, int64_t svec, long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    float a[8], b[8];
    sp_ld8(x + i * 8, split, a); sp_ld8(pos + (i % svec) * 8, split, b);
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] += b[k];
    sp_st8(out + i * 8, split, a);
}
__global__ void add_kernel(const uint4 *x, const uint4 *y, uint4 *out, int64_t nvec)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    if (!y) { out[i] = x[i]; return; }
    float a[8], b[8];
    unpack8(x[i], a); unpack8(y[i], b);
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] += b[k];
    out[i] = pack8(a);
}
__global__ void image_to_nhwc4_kernel(const float *img, uint2 *out, int64_t npix)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float r = img[i * 3], g = img[i * 3 + 1], b = img[i * 3 + 2];
    out[i] = make_uint2(pack_bf16x2(r, g), pack_bf16x2(b, 0.f));
}
// space-to-depth(2) of the fp32 NHWC3 image: out[b, y, x, (ry*2+rx)*3 + c] = img[b, 2y+ry, 2x+rx, c] (0 outside, 4 zero pad channels)
// -> the 7x7 / stride-2 stem becomes a dense 4x4 / stride-1 convolution over 16-channel pixels (32 B: four consecutive pixels are
//    one 128-byte row of the sliding-window GEMM operand, see detrb_igemm_t.a_kb_rows)
__global__ void image_to_s2d16_kernel(const float *img, bf16 *out, int B, int H, int W, int H2, int W2, int pt, int pl, int HP, int WP,
                                      long long split)
{
    pdl_trigger();
    pdl_wait();
    // one thread per pixel of the (zero-padded) output [B, HP, WP, 16]; the frame's pixel (y2, x2) lands at (pt + y2, pl + x2)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HP * WP) return;
    int xp = i % WP; int yp = (i / WP) % HP; int b = i / ((int64_t)WP * HP);
    const int x2 = xp - pl, y2 = yp - pt;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.f;
    if (x2 >= 0 && x2 < W2 && y2 >= 0 && y2 < H2) {
#pragma unroll
        for (int ry = 0; ry < 2; ry++)
#pragma unroll
            for (int rx = 0; rx < 2; rx++) {
                int y = 2 * y2 + ry, x = 2 * x2 + rx;
                if (y < H && x < W) {
                    const float *px = img + (((size_t)b * H + y) * W + x) * 3;
                    v[(ry * 2 + rx) * 3 + 0] = px[0]; v[(ry * 2 + rx) * 3 + 1] = px[1]; v[(ry * 2 + rx) * 3 + 2] = px[2];
                }
            }
    }
    float lo8[8], hi8[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { lo8[k] = v[k]; hi8[k] = v[8 + k]; }
    sp_st8(out + i * 16, split, lo8);
    sp_st8(out + i * 16 + 8, split, hi8);
}
__global__ void f32_to_bf16_kernel(const float *x, bf16 *y, int64_t n)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16(x[i]);
}

// out[n] += scale[n] * sum_m x[m, n]; block = 32 column-pairs x 8 row lanes
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16 *x, int ldx, int M, int N, const float *scale, float *out, int rows_per_block)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    __shared__ float red[8][64];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = blockIdx.x * 64 + cx * 2;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float a0 = 0.f, a1 = 0.f;
    if (n < N) {
        for (int r = r0 + ry; r < r1; r += 8) {
            float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(x + (size_t)r * ldx + n));
            a0 += v.x; a1 += v.y;
        }
    }
    red[ry][cx * 2] = a0; red[ry][cx * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.x < 64) {
        int nn = blockIdx.x * 64 + threadIdx.x;
        if (nn < N) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) s += red[w][threadIdx.x];
            atomicAdd(out + nn, s * (scale ? scale[nn] : 1.f));
        }
    }
}

// ---------------------------------------------------------------- max pool 3x3 s2 pad 1 (zero pad == -inf pad: x >= 0)
__global__ void maxpool_fwd_kernel(const bf16 *x, bf16 *y, uint8_t *argmax, int B, int IH, int IW, int C, int OH, int OW, int XH, int XW,
                                   long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    // grid = (row segments, OH, B): no 64-bit divisions in the index math
    const int cv = C / 8;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ox = tx / cv, c8 = tx - ox * cv;
    if (ox >= OW) return;
    const int oy = blockIdx.y, b = blockIdx.z;
    const size_t pix = ((size_t)b * OH + oy) * OW + ox;
    float best[8]; int arg[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { best[i] = -INFINITY; arg[i] = 0; }
    // padded positions hold 0 (ZeroPadding2D): a window of all-negative values cannot occur post-ReLU,
    // so treating the pad as "absent" is identical; taps are scanned in row-major order, first max wins.
#pragma unroll
    for (int kh = 0; kh < 3; kh++)
#pragma unroll
        for (int kw = 0; kw < 3; kw++) {
            int iy = oy * 2 - 1 + kh, ix = ox * 2 - 1 + kw;
            if (iy < 0 || iy >= IH || ix < 0 || ix >= IW) continue;
            float v[8];
            sp_ld8(x + (((size_t)b * XH + iy) * XW + ix) * C + c8 * 8, split, v);   // x is [B, XH, XW, C]
#pragma unroll
            for (int i = 0; i < 8; i++) if (v[i] > best[i]) { best[i] = v[i]; arg[i] = kh * 3 + kw; }
        }
    size_t o = (size_t)pix * C + c8 * 8;
    sp_st8(y + o, split, best);
    // a window whose maximum is not positive passes no gradient through the stem's ReLU: tap 15 matches no input position,
    // so the backward pass needs neither x nor y
#pragma unroll
    for (int i = 0; i < 8; i++) if (!(best[i] > 0.f)) arg[i] = 15;
    uint2 a;
    a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2 *>(argmax + o) = a;
}

// The same pooling for plain bf16 storage (split == 0), on packed bf16x2 values: the fp32 form above is bound by instruction issue
// (ncu: 89 M warp instructions, issue slots 70 % busy, 121 us for 357 MB), this one needs three instructions per channel pair and tap
// (compare mask, two bit selects).  bf16 -> fp32 is exact and order preserving, so maxima, tie rule (first tap in row-major order
// wins: strict >) and the "not positive -> tap 15" rule are bit-identical to the fp32 form.
__device__ __forceinline__ uint32_t bf16x2_gt_mask(uint32_t a, uint32_t b) {       // 0xFFFF in each half where a > b (ordered)
    return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
}
__global__ void __launch_bounds__(256)
maxpool_fwd_bf16_kernel(const bf16 *x, bf16 *y, uint8_t *argmax, int IH, int IW, int C, int OH, int OW, int XH, int XW)
{
    pdl_trigger();
    pdl_wait();
    const int cv = C / 8;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ox = tx / cv, c8 = tx - ox * cv;
    if (ox >= OW) return;
    const int oy = blockIdx.y, b = blockIdx.z;
    uint32_t best[4], arg[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { best[i] = 0xFF80FF80u; arg[i] = 0u; }        // (-inf, -inf)
#pragma unroll
    for (int kh = 0; kh < 3; kh++)
#pragma unroll
        for (int kw = 0; kw < 3; kw++) {
            const int iy = oy * 2 - 1 + kh, ix = ox * 2 - 1 + kw;
            if (iy < 0 || iy >= IH || ix < 0 || ix >= IW) continue;
            const uint4 u = *reinterpret_cast<const uint4 *>(x + (((size_t)b * XH + iy) * XW + ix) * C + c8 * 8);
            const uint32_t v[4] = {u.x, u.y, u.z, u.w};
            const uint32_t t2 = (uint32_t)(kh * 3 + kw) * 0x00010001u;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t m = bf16x2_gt_mask(v[i], best[i]);
                best[i] = (v[i] & m) | (best[i] & ~m);
                arg[i] = (t2 & m) | (arg[i] & ~m);
            }
        }
    const size_t o = (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8;
    *reinterpret_cast<uint4 *>(y + o) = make_uint4(best[0], best[1], best[2], best[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t m = bf16x2_gt_mask(best[i], 0u);
        arg[i] = (arg[i] & m) | (0x000F000Fu & ~m);
    }
    uint2 a;
    a.x = prmt(arg[0], arg[1], 0x6420u);            // the four 16-bit taps of two words -> four bytes
    a.y = prmt(arg[2], arg[3], 0x6420u);
    *reinterpret_cast<uint2 *>(argmax + o) = a;
}

// dx[b,iy,ix,c] = sum over the <= 4 windows containing (iy,ix) whose argmax is this pixel (the stem's ReLU mask is already in the
// argmax: non-positive maxima were stored as tap 15).  dx is [B, XH, XW, C] (XH >= IH, XW >= IW): the positions outside
// IH x IW are written as zeros.
// One thread owns the 2x2 input block (2a..2a+1, 2b..2b+1) x 8 channels: the only windows that reach it are (a,b), (a,b+1),
// (a+1,b), (a+1,b+1) -- four (argmax, dy) loads for four stores (a thread per pixel loads 2.25 windows per store).
__global__ void maxpool_bwd_kernel(const bf16 *dy, const uint8_t *argmax, bf16 *dx,
                                   int B, int IH, int IW, int C, int OH, int OW, int XH, int XW, long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const int cv = C / 8;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int bq = tx / cv, c8 = tx - bq * cv;
    const int a = blockIdx.y, b = blockIdx.z;
    if (2 * bq >= XW) return;
    float acc[2][2][8];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[r][c][i] = 0.f;
#pragma unroll
    for (int wy = 0; wy < 2; wy++)
#pragma unroll
        for (int wx = 0; wx < 2; wx++) {
            const int oy = a + wy, ox = bq + wx;
            if (oy >= OH || ox >= OW) continue;
            const size_t o = (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8;
            const uint2 am = *reinterpret_cast<const uint2 *>(argmax + o);
            float d[8];
            sp_ld8(dy + o, split, d);
            // window (oy, ox) covers input (2oy-1+kh, 2ox-1+kw): inside this block rows r >= wy, columns c >= wx, reached by
            // kh = r + 1 - 2wy, kw = c + 1 - 2wx
#pragma unroll
            for (int r = wy; r < 2; r++)
#pragma unroll
                for (int c = wx; c < 2; c++) {
                    const int tap = (r + 1 - 2 * wy) * 3 + (c + 1 - 2 * wx);
                    // byte-wise compare of the 8 argmax bytes with `tap`: (a ^ tap x 0x01010101) has a zero byte where they match
                    const uint32_t t4 = (uint32_t)tap * 0x01010101u;
                    const uint32_t m0 = am.x ^ t4, m1 = am.y ^ t4;
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if ((((i < 4 ? m0 : m1) >> ((i & 3) * 8)) & 0xffu) == 0u) acc[r][c][i] += d[i];
                }
        }
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int iy = 2 * a + r, ix = 2 * bq + c;
            if (iy >= XH || ix >= XW) continue;
            bf16 *dst = dx + (((size_t)b * XH + iy) * XW + ix) * C + c8 * 8;
            const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (iy < IH && ix < IW) sp_st8(dst, split, acc[r][c]); else sp_st8(dst, split, zero8);
        }
}

// The same scatter for plain bf16 storage (split == 0) with the tap tests done on whole words: the form above spends a byte
// extraction, a compare and a predicated add per (window position, channel) and is bound by instruction issue (ncu: 95 M warp
// instructions, issue slots 77 % busy, 115 us for 324 MB).  Here one exact zero-byte test per argmax word marks the channels whose
// tap is this position (bit 7 of each byte), two PRMTs turn the marks into bf16x2 AND-masks, and the masked gradients are added
// in fp32 in the same window order: acc + (match ? d : +0) == the predicated add, bit for bit.
__global__ void __launch_bounds__(256)
maxpool_bwd_bf16_kernel(const bf16 *dy, const uint8_t *argmax, bf16 *dx, int IH, int IW, int C, int OH, int OW, int XH, int XW)
{
    pdl_trigger();
    pdl_wait();
    const int cv = C / 8;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int bq = tx / cv, c8 = tx - bq * cv;
    const int a = blockIdx.y, b = blockIdx.z;
    if (2 * bq >= XW) return;
    float acc[2][2][8];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[r][c][i] = 0.f;
#pragma unroll
    for (int wy = 0; wy < 2; wy++)
#pragma unroll
        for (int wx = 0; wx < 2; wx++) {
            const int oy = a + wy, ox = bq + wx;
            if (oy >= OH || ox >= OW) continue;
            const size_t o = (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8;
            const uint2 am = *reinterpret_cast<const uint2 *>(argmax + o);
            const uint4 dv = *reinterpret_cast<const uint4 *>(dy + o);
            const uint32_t d2[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int r = wy; r < 2; r++)
#pragma unroll
                for (int c = wx; c < 2; c++) {
                    const uint32_t t4 = (uint32_t)((r + 1 - 2 * wy) * 3 + (c + 1 - 2 * wx)) * 0x01010101u;
                    const uint32_t z0 = am.x ^ t4, z1 = am.y ^ t4;
                    // bit 7 of a byte <=> that byte of z is zero (exact)
                    const uint32_t e0 = ~(((z0 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z0) & 0x80808080u;
                    const uint32_t e1 = ~(((z1 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z1) & 0x80808080u;
                    // sign replication: channel pair (2p, 2p + 1) -> halves of all ones / all zeros
                    const uint32_t m[4] = {prmt(e0, 0u, 0x9988u), prmt(e0, 0u, 0xBBAAu), prmt(e1, 0u, 0x9988u), prmt(e1, 0u, 0xBBAAu)};
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const uint32_t v = d2[p] & m[p];
                        acc[r][c][2 * p] += __uint_as_float(v << 16);
                        acc[r][c][2 * p + 1] += __uint_as_float(v & 0xFFFF0000u);
                    }
                }
        }
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int iy = 2 * a + r, ix = 2 * bq + c;
            if (iy >= XH || ix >= XW) continue;
            bf16 *dst = dx + (((size_t)b * XH + iy) * XW + ix) * C + c8 * 8;
            const bool inside = iy < IH && ix < IW;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (inside) out = sp_pack8(acc[r][c]);
            *reinterpret_cast<uint4 *>(dst) = out;
        }
}

__global__ void dropout_mask_kernel(uint8_t *out, int M, int N, float drop_p, uint64_t seed_in, uint32_t site, const uint64_t *seed_ptr)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const uint64_t seed = seed_in ^ (seed_ptr ? *seed_ptr : 0ull);
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    int m = idx / N, n = idx % N;
    uint32_t bits = dropout_bits(seed, site, (uint32_t)m, (uint32_t)(n >> 1));
    uint32_t v = (n & 1) ? (bits >> 16) : (bits & 0xffffu);
    out[idx] = v >= dropout_thresh16(drop_p) ? 1 : 0;
}

// keep mask of an attention-probability dropout site: row = (b*H + h)*Lq + q, column = key (attn_drop_word, common.cuh)
__global__ void attn_dropout_mask_kernel(uint8_t *out, int M, int N, float drop_p, uint64_t seed_in, uint32_t site, const uint64_t *seed_ptr)
{
    pdl_trigger();
    pdl_wait();
    const uint64_t seed = seed_in ^ (seed_ptr ? *seed_ptr : 0ull);
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    const uint32_t m = (uint32_t)(idx / N), k = (uint32_t)(idx % N);
    const uint32_t kb = attn_drop_keepbits(attn_drop_word(attn_drop_rowhash(seed, site, m), attn_drop_pair(k)), attn_drop_thresh2(drop_p));
    out[idx] = (((k >> 3) & 1u) ? attn_drop_mask_hi(kb) : attn_drop_mask_lo(kb)) & 1u;
}

}  // namespace

extern "C" int detrb_layernorm_fwd(const detrb_bf16 *x, const float *gamma, const float *beta, detrb_bf16 *y, detrb_bf16 *y2,
                                   const detrb_bf16 *pos, int S, float *mean, float *rstd, int M, int64_t split, detrb_stream_t stream)
{
    DETRB_REQUIRE(x && gamma && beta && y && M > 0, "detrb_layernorm_fwd: bad args");
    DETRB_REQUIRE(!y2 || (pos && S > 0), "detrb_layernorm_fwd: y2 needs pos and S");
    DETRB_LAUNCH(ln_fwd_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)x, gamma, beta, (bf16 *)y, (bf16 *)y2,
                                                                   (const bf16 *)pos, S, mean, rstd, M, (long long)split);
    DETRB_CHECK_LAUNCH("ln_fwd_kernel");
    return DETRB_OK;
}

extern "C" int detrb_layernorm_bwd(const detrb_bf16 *dy, const detrb_bf16 *dy2, const detrb_bf16 *x, const float *gamma,
                                   const float *mean, const float *rstd, detrb_bf16 *dx, detrb_bf16 *dx_drop,
                                   float drop_p, uint64_t seed, uint32_t site, const uint64_t *seed_ptr,
                                   float *dgamma, float *dbeta, int M, int64_t split, detrb_stream_t stream)
{
    DETRB_REQUIRE(dy && x && gamma && mean && rstd && dx && M > 0, "detrb_layernorm_bwd: bad args");
    DETRB_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "detrb_layernorm_bwd: dgamma/dbeta must both be set or both NULL");
    int blocks = ceil_div(M, 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    int rpb = ceil_div(ceil_div(M, blocks), 8) * 8;
    blocks = ceil_div(M, rpb);
    DETRB_LAUNCH(ln_bwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)dy, (const bf16 *)dy2, (const bf16 *)x, gamma, mean, rstd,
                                                            (bf16 *)dx, (bf16 *)dx_drop, drop_p, seed, site, seed_ptr, dgamma, dbeta, M, rpb, (long long)split);
    DETRB_CHECK_LAUNCH("ln_bwd_kernel");
    return DETRB_OK;
}

extern "C" int detrb_add_rowbcast(const detrb_bf16 *x, const detrb_bf16 *pos, detrb_bf16 *out, int M, int S, int d, int64_t split,
                                  detrb_stream_t stream)
{
    DETRB_REQUIRE(x && pos && out && M > 0 && S > 0 && d % 8 == 0, "detrb_add_rowbcast: bad args");
    int64_t nvec = (int64_t)M * d / 8, svec = (int64_t)S * d / 8;
    DETRB_LAUNCH(add_rowbcast_kernel, dim3((unsigned)((nvec + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)x, (const bf16 *)pos, (bf16 *)out, nvec, svec, (long long)split);
    DETRB_CHECK_LAUNCH("add_rowbcast_kernel");
    return DETRB_OK;
}

extern "C" int detrb_add(const detrb_bf16 *a, const detrb_bf16 *b, detrb_bf16 *out, int64_t n, detrb_stream_t stream)
{
    DETRB_REQUIRE(a && out && n > 0 && n % 8 == 0, "detrb_add: bad args");
    int64_t nvec = n / 8;
    DETRB_LAUNCH(add_kernel, dim3((unsigned)((nvec + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const uint4 *)a, (const uint4 *)b, (uint4 *)out, nvec);
    DETRB_CHECK_LAUNCH("add_kernel");
    return DETRB_OK;
}

extern "C" int detrb_image_to_nhwc4(const float *img, detrb_bf16 *out, int64_t npix, detrb_stream_t stream)
{
    DETRB_REQUIRE(img && out && npix > 0, "detrb_image_to_nhwc4: bad args");
    DETRB_LAUNCH(image_to_nhwc4_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, img, (uint2 *)out, npix);
    DETRB_CHECK_LAUNCH("image_to_nhwc4_kernel");
    return DETRB_OK;
}

extern "C" int detrb_image_to_s2d16(const float *img, detrb_bf16 *out, int B, int H, int W, int pad_top, int pad_left, int HP, int WP,
                                    int64_t split, detrb_stream_t stream)
{
    DETRB_REQUIRE(img && out && B > 0 && H > 0 && W > 0, "detrb_image_to_s2d16: bad args");
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    DETRB_REQUIRE(pad_top >= 0 && pad_left >= 0 && HP >= H2 + pad_top && WP >= W2 + pad_left, "detrb_image_to_s2d16: padded size too small");
    int64_t n = (int64_t)B * HP * WP;
    DETRB_LAUNCH(image_to_s2d16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, img, (bf16 *)out, B, H, W, H2, W2,
                 pad_top, pad_left, HP, WP, (long long)split);
    DETRB_CHECK_LAUNCH("image_to_s2d16_kernel");
    return DETRB_OK;
}

extern "C" int detrb_f32_to_bf16(const float *x, detrb_bf16 *y, int64_t n, detrb_stream_t stream)
{
    DETRB_REQUIRE(x && y && n > 0, "detrb_f32_to_bf16: bad args");
    DETRB_LAUNCH(f32_to_bf16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, (bf16 *)y, n);
    DETRB_CHECK_LAUNCH("f32_to_bf16_kernel");
    return DETRB_OK;
}

extern "C" int detrb_colsum(const detrb_bf16 *x, int ldx, int M, int N, const float *scale, float *out, detrb_stream_t stream)
{
    DETRB_REQUIRE(x && out && M > 0 && N > 0 && ldx % 2 == 0 && N % 2 == 0, "detrb_colsum: bad args");
    int gy = ceil_div(M, 256);
    if (gy > 296) gy = 296;
    int rpb = ceil_div(M, gy);
    gy = ceil_div(M, rpb);
    DETRB_LAUNCH(colsum_kernel, dim3(dim3(ceil_div(N, 64), gy)), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)x, ldx, M, N, scale, out, rpb);
    DETRB_CHECK_LAUNCH("colsum_kernel");
    return DETRB_OK;
}

// developer switch (env DETRB_POOL_FP32=1): the fp32 pooling kernels also for plain bf16 storage (tests run both forms)
static bool pool_fp32_forced()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DETRB_POOL_FP32"); v = e ? atoi(e) : 0; }
    return v != 0;
}

extern "C" int detrb_maxpool_fwd(const detrb_bf16 *x, detrb_bf16 *y, uint8_t *argmax, int B, int IH, int IW, int C, int OH, int OW,
                                 int XH, int XW, int64_t split, detrb_stream_t stream)
{
    DETRB_REQUIRE(x && y && argmax && C % 8 == 0, "detrb_maxpool_fwd: bad args");
    DETRB_REQUIRE(OH == (IH + 2 - 3) / 2 + 1 && OW == (IW + 2 - 3) / 2 + 1, "detrb_maxpool_fwd: bad output size");
    DETRB_REQUIRE(OH <= 65535 && B <= 65535, "detrb_maxpool_fwd: grid too large");
    DETRB_REQUIRE(XH >= IH && XW >= IW, "detrb_maxpool_fwd: allocated extent smaller than the image");
    const dim3 grid((unsigned)ceil_div(OW * (C / 8), 256), (unsigned)OH, (unsigned)B);
    if (split == 0 && !pool_fp32_forced() && !(((uintptr_t)x | (uintptr_t)y) & 15) && !((uintptr_t)argmax & 7))
        DETRB_LAUNCH(maxpool_fwd_bf16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)x, (bf16 *)y, argmax, IH, IW, C, OH, OW, XH, XW);
    else
        DETRB_LAUNCH(maxpool_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)x, (bf16 *)y, argmax, B, IH, IW, C, OH, OW, XH, XW, (long long)split);
    DETRB_CHECK_LAUNCH("maxpool_fwd_kernel");
    return DETRB_OK;
}

extern "C" int detrb_maxpool_bwd(const detrb_bf16 *dy, const uint8_t *argmax, detrb_bf16 *dx,
                                 int B, int IH, int IW, int C, int OH, int OW, int XH, int XW, int64_t split, detrb_stream_t stream)
{
    DETRB_REQUIRE(dy && argmax && dx && C % 8 == 0, "detrb_maxpool_bwd: bad args");
    DETRB_REQUIRE(XH <= 65535 && B <= 65535, "detrb_maxpool_bwd: grid too large");
    DETRB_REQUIRE(XH >= IH && XW >= IW, "detrb_maxpool_bwd: allocated extent smaller than the image");
    const dim3 grid((unsigned)ceil_div(((XW + 1) / 2) * (C / 8), 256), (unsigned)((XH + 1) / 2), (unsigned)B);
    if (split == 0 && !pool_fp32_forced() && !(((uintptr_t)dy | (uintptr_t)dx) & 15) && !((uintptr_t)argmax & 7))
        DETRB_LAUNCH(maxpool_bwd_bf16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)dy, argmax, (bf16 *)dx, IH, IW, C, OH, OW, XH, XW);
    else
        DETRB_LAUNCH(maxpool_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16 *)dy, argmax, (bf16 *)dx,
                     B, IH, IW, C, OH, OW, XH, XW, (long long)split);
    DETRB_CHECK_LAUNCH("maxpool_bwd_kernel");
    return DETRB_OK;
}

extern "C" int detrb_attn_dropout_mask(uint8_t *out, int M, int N, float drop_p, uint64_t seed, uint32_t site,
                                       const uint64_t *seed_ptr, detrb_stream_t stream)
{
    DETRB_REQUIRE(out && M > 0 && N > 0, "detrb_attn_dropout_mask: bad args");
    int64_t total = (int64_t)M * N;
    DETRB_LAUNCH(attn_dropout_mask_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, out, M, N, drop_p, seed, site, seed_ptr);
    DETRB_CHECK_LAUNCH("attn_dropout_mask_kernel");
    return DETRB_OK;
}

extern "C" int detrb_dropout_mask(uint8_t *out, int M, int N, float drop_p, uint64_t seed, uint32_t site,
                                  const uint64_t *seed_ptr, detrb_stream_t stream)
{
    DETRB_REQUIRE(out && M > 0 && N > 0, "detrb_dropout_mask: bad args");
    int64_t total = (int64_t)M * N;
    DETRB_LAUNCH(dropout_mask_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, out, M, N, drop_p, seed, site, seed_ptr);
    DETRB_CHECK_LAUNCH("dropout_mask_kernel");
    return DETRB_OK;
}

int detrb_scatter_s2(const bf16 *const src[4], bf16 *dst, int ldc, const uint8_t *mask_bits, int ldmb, int B, int H, int W, int C,
                     int accumulate, cudaStream_t stream)
{
    DETRB_REQUIRE(dst && C % 8 == 0 && ldc % 8 == 0 && B > 0 && H > 0 && W > 0 && H <= 65535 && B <= 65535, "detrb_scatter_s2: bad args");
    const int C8 = C / 8;
    DETRB_LAUNCH(scatter_s2_kernel, dim3((unsigned)ceil_div(W * C8, 256), (unsigned)H, (unsigned)B), dim3(256), 0, stream, src[0], src[1], src[2],
                 src[3], dst, ldc, mask_bits, ldmb, H, W, C8, accumulate);
    DETRB_CHECK_LAUNCH("scatter_s2_kernel");
    return DETRB_OK;
}
