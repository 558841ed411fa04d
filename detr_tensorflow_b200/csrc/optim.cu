// optim.cu -- optimizer step of the reference (optimizers.py:86-88,137-163): three Keras Adam optimizers
// (beta1 .9, beta2 .999, eps 1e-7, no weight decay) with PER-VARIABLE clipnorm, as one multi-tensor pass over a
// flat fp32 arena, plus the fp32-master -> bf16 kernel-layout weight refresh.
#include "common.cuh"

namespace {

constexpr int SLICES = 16;

__global__ void adam_prologue_kernel(int32_t *steps, const uint8_t *enabled, int G, float *norms, int T)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G && enabled[i]) steps[i] += 1;
    if (i < T) norms[i] = 0.f;
}

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float *grads, const int64_t *table, float *norms)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const int t = blockIdx.x, s = blockIdx.y;
    const int64_t off = table[2 * t], n = table[2 * t + 1];
    const int64_t per = ((n + SLICES - 1) / SLICES + 3) & ~(int64_t)3;
    const int64_t b = s * per, e = min(n, b + per);
    float acc = 0.f;
    for (int64_t i = b + threadIdx.x; i < e; i += 256) { float g = grads[off + i]; acc += g * g; }
    acc = warp_sum(acc);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) tot += red[w];
        if (b < e) atomicAdd(norms + t, tot);
    }
}

__global__ void __launch_bounds__(256)
adam_update_kernel(float *params, const float *grads, float *m, float *v, const int64_t *table,
                   const int32_t *lr_group, const float *lrs, const uint8_t *enabled, const int32_t *steps,
                   const float *norms, float clipnorm, float beta1, float beta2, float eps)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const int t = blockIdx.x, s = blockIdx.y;
    const int g = lr_group[t];
    if (!enabled[g]) return;
    const int64_t off = table[2 * t], n = table[2 * t + 1];
    const int64_t per = ((n + SLICES - 1) / SLICES + 3) & ~(int64_t)3;
    const int64_t b = s * per, e = min(n, b + per);
    const float norm = sqrtf(norms[t]);
    const float coef = (clipnorm > 0.f && norm > clipnorm) ? clipnorm / norm : 1.f;   // tf.clip_by_norm
    const float step = (float)steps[g];
    const float lr_t = lrs[g] * sqrtf(1.f - powf(beta2, step)) / (1.f - powf(beta1, step));
    for (int64_t i = b + threadIdx.x; i < e; i += 256) {
        float gr = grads[off + i] * coef;
        float mi = beta1 * m[off + i] + (1.f - beta1) * gr;
        float vi = beta2 * v[off + i] + (1.f - beta2) * gr * gr;
        m[off + i] = mi; v[off + i] = vi;
        params[off + i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

__global__ void prep_weight_kernel(const float *master, const float *fold, int N, int taps, int Cin,
                                   bf16 *Wf, int ldf, bf16 *Wd, int ldd)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)N * taps * Cin;
    if (idx >= total) return;
    int c = idx % Cin; int t = (idx / Cin) % taps; int n = idx / ((int64_t)Cin * taps);
    float w = master[idx] * (fold ? fold[n] : 1.f);
    bf16 wb = __float2bfloat16(w);
    if (Wf) Wf[(size_t)n * ldf + t * Cin + c] = wb;
    if (Wd) Wd[((size_t)c * taps + t) * ldd + n] = wb;
}

// ---- chunked variants: uniform work items {tensor, start, len} (len <= 8192, start 16-byte aligned) and float4 accesses
__global__ void __launch_bounds__(256)
chunk_sumsq_kernel(const float *grads, const int32_t *chunks, float *norms)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const int t = chunks[3 * blockIdx.x], start = chunks[3 * blockIdx.x + 1], len = chunks[3 * blockIdx.x + 2];
    const float4 *g4 = reinterpret_cast<const float4 *>(grads + start);
    float acc = 0.f;
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) { float4 g = g4[i]; acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w; }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += 256) { float g = grads[start + i]; acc += g * g; }
    acc = warp_sum(acc);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) tot += red[w];
        atomicAdd(norms + t, tot);
    }
}

__global__ void __launch_bounds__(256)
chunk_adam_kernel(float *params, const float *grads, float *m, float *v, const int32_t *chunks, const int32_t *lr_group,
                  const float *lrs, const uint8_t *enabled, const int32_t *steps, const float *norms,
                  float clipnorm, float beta1, float beta2, float eps)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const int t = chunks[3 * blockIdx.x], start = chunks[3 * blockIdx.x + 1], len = chunks[3 * blockIdx.x + 2];
    const int g = lr_group[t];
    if (!enabled[g]) return;
    const float norm = sqrtf(norms[t]);
    const float coef = (clipnorm > 0.f && norm > clipnorm) ? clipnorm / norm : 1.f;
    const float step = (float)steps[g];
    const float lr_t = lrs[g] * sqrtf(1.f - powf(beta2, step)) / (1.f - powf(beta1, step));
    const float ob1 = 1.f - beta1, ob2 = 1.f - beta2;
    float4 *p4 = reinterpret_cast<float4 *>(params + start), *m4 = reinterpret_cast<float4 *>(m + start), *v4 = reinterpret_cast<float4 *>(v + start);
    const float4 *g4 = reinterpret_cast<const float4 *>(grads + start);
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
        float4 gr = g4[i], mi = m4[i], vi = v4[i], pi = p4[i];
        gr.x *= coef; gr.y *= coef; gr.z *= coef; gr.w *= coef;
        mi.x = beta1 * mi.x + ob1 * gr.x; mi.y = beta1 * mi.y + ob1 * gr.y; mi.z = beta1 * mi.z + ob1 * gr.z; mi.w = beta1 * mi.w + ob1 * gr.w;
        vi.x = beta2 * vi.x + ob2 * gr.x * gr.x; vi.y = beta2 * vi.y + ob2 * gr.y * gr.y;
        vi.z = beta2 * vi.z + ob2 * gr.z * gr.z; vi.w = beta2 * vi.w + ob2 * gr.w * gr.w;
        pi.x -= lr_t * mi.x / (sqrtf(vi.x) + eps); pi.y -= lr_t * mi.y / (sqrtf(vi.y) + eps);
        pi.z -= lr_t * mi.z / (sqrtf(vi.z) + eps); pi.w -= lr_t * mi.w / (sqrtf(vi.w) + eps);
        m4[i] = mi; v4[i] = vi; p4[i] = pi;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += 256) {
        float gr = grads[start + i] * coef;
        float mi = beta1 * m[start + i] + ob1 * gr, vi = beta2 * v[start + i] + ob2 * gr * gr;
        m[start + i] = mi; v[start + i] = vi;
        params[start + i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// multi-tensor weight refresh: one launch for all layers.  Work item = 64x64 (n, c) tile of one filter tap of one weight;
// the tile is read coalesced along c (256-byte rows), written to Wf along c and, through a shared-memory transpose, to Wd
// along n -- both as packed bf16 pairs in full 128-byte rows.  Cin and the row strides are even (checked on the host).
constexpr int PT = 64;
__global__ void __launch_bounds__(256)
prep_weights_multi_kernel(const detrb_prep_desc_t *descs, int nslots, long long wsplit)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    __shared__ float tile[PT][PT + 1];
    const int t = blockIdx.x;
    // this tile's slot = the last one whose tile_begin <= t (tile_begin ascends from 0): counted by the whole block with one load per
    // thread -- a binary search is seven DEPENDENT global loads at the start of a block that lives for ~3 us
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < nslots; i += 256) cnt += descs[i].tile_begin <= t ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    const int lo = s_cnt - 1;
    const detrb_prep_desc_t d = descs[lo];
    const int local = t - d.tile_begin;
    const int ct = (d.Cin + PT - 1) / PT, nt = (d.N + PT - 1) / PT;
    const int ci = local % ct, ni = (local / ct) % nt, tap = local / (ct * nt);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = ci * PT + 2 * tx;                    // this thread's column pair
    for (int r = ty; r < PT; r += 8) {
        const int n = ni * PT + r;
        float2 w = make_float2(0.f, 0.f);
        if (n < d.N && c < d.Cin) {
            w = *reinterpret_cast<const float2 *>(d.master + ((size_t)n * d.taps + tap) * d.Cin + c);
            const float f = d.fold ? d.fold[n] : 1.f;
            w.x *= f; w.y *= f;
            if (d.Wf) sp_st2(reinterpret_cast<bf16 *>(d.Wf) + (size_t)n * d.ldf + tap * d.Cin + c, wsplit, w.x, w.y);
        }
        tile[r][2 * tx] = w.x; tile[r][2 * tx + 1] = w.y;
    }
    if (!d.Wd) return;
    __syncthreads();
    const int n = ni * PT + 2 * tx;                    // this thread's output-channel pair
    for (int r = ty; r < PT; r += 8) {
        const int cc = ci * PT + r;
        if (cc >= d.Cin || n >= d.N) continue;
        bf16 *dst = reinterpret_cast<bf16 *>(d.Wd) + ((size_t)cc * d.taps + tap) * d.ldd + n;
        if (n + 1 < d.N) sp_st2(dst, wsplit, tile[2 * tx][r], tile[2 * tx + 1][r]);
        else sp_st1(dst, wsplit, tile[2 * tx][r]);
    }
}

// gradient accumulation over micro-batches (optimizers.py:150-157): acc = (zero_first ? 0 : acc) + g, float4 grid-stride
__global__ void __launch_bounds__(256)
accumulate_kernel(float4 *acc, const float4 *g, long long n4, int zero_first)
{
    pdl_trigger();
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 b = g[i];
        float4 a = zero_first ? make_float4(0.f, 0.f, 0.f, 0.f) : acc[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        acc[i] = a;
    }
}

}  // namespace

extern "C" int detrb_accumulate(float *acc, const float *g, int64_t n, int zero_first, detrb_stream_t stream)
{
    DETRB_REQUIRE(acc && g && n > 0 && n % 4 == 0 && (((uintptr_t)acc | (uintptr_t)g) & 15) == 0, "detrb_accumulate: bad args");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    DETRB_LAUNCH(accumulate_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<float4 *>(acc),
                 reinterpret_cast<const float4 *>(g), n4, zero_first);
    DETRB_CHECK_LAUNCH("accumulate_kernel");
    return DETRB_OK;
}

extern "C" int detrb_prep_weights_multi(const detrb_prep_desc_t *descs, int nslots, int total_tiles, int64_t wsplit, detrb_stream_t stream)
{
    DETRB_REQUIRE(descs && nslots > 0 && total_tiles > 0, "detrb_prep_weights_multi: bad args");
    // descs live in device memory (the table is built once by the caller): tile_begin counts 64x64 tiles, taps * ceil(N/64) *
    // ceil(Cin/64) per slot; Cin, ldf, ldd even and master / Wf / Wd 8 / 4 / 4-byte aligned
    DETRB_LAUNCH(prep_weights_multi_kernel, dim3(total_tiles), dim3(256), 0, (cudaStream_t)stream, descs, nslots, (long long)wsplit);
    DETRB_CHECK_LAUNCH("prep_weights_multi_kernel");
    return DETRB_OK;
}

extern "C" int detrb_adam_clipnorm(float *params, const float *grads, float *m, float *v, const int64_t *table,
                                   const int32_t *lr_group, const float *lrs, const uint8_t *group_enabled, int T,
                                   int64_t total, float clipnorm, float beta1, float beta2, float eps,
                                   int32_t *steps, float *norms, detrb_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    DETRB_REQUIRE(params && grads && m && v && table && lr_group && lrs && group_enabled && steps && norms, "detrb_adam_clipnorm: null pointer");
    DETRB_REQUIRE(T > 0 && T <= 65535 && total > 0, "detrb_adam_clipnorm: bad sizes");
    // groups: at most 8 (the reference has 3: backbone, transformers, nlayers)
    DETRB_LAUNCH(adam_prologue_kernel, dim3(ceil_div((T > 8 ? T : 8), 256)), dim3(256), 0, stream, steps, group_enabled, 8, norms, T);
    DETRB_CHECK_LAUNCH("adam_prologue_kernel");
    DETRB_LAUNCH(grad_sumsq_kernel, dim3(dim3(T, SLICES)), dim3(256), 0, stream, grads, table, norms);
    DETRB_CHECK_LAUNCH("grad_sumsq_kernel");
    DETRB_LAUNCH(adam_update_kernel, dim3(dim3(T, SLICES)), dim3(256), 0, stream, params, grads, m, v, table, lr_group, lrs, group_enabled, steps, norms,
                                                           clipnorm, beta1, beta2, eps);
    DETRB_CHECK_LAUNCH("adam_update_kernel");
    return DETRB_OK;
}

extern "C" int detrb_adam_clipnorm_chunked(float *params, const float *grads, float *m, float *v, const int32_t *chunks, int nchunks,
                                           const int32_t *lr_group, const float *lrs, const uint8_t *group_enabled, int T,
                                           float clipnorm, float beta1, float beta2, float eps, int32_t *steps, float *norms,
                                           int prologue, detrb_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    DETRB_REQUIRE(params && grads && m && v && chunks && lr_group && lrs && group_enabled && steps && norms, "detrb_adam_clipnorm_chunked: null pointer");
    DETRB_REQUIRE(T > 0 && nchunks > 0, "detrb_adam_clipnorm_chunked: bad sizes");
    if (prologue) {                // first call of an optimizer step: bump the enabled groups' iteration counters, zero every norm
        DETRB_LAUNCH(adam_prologue_kernel, dim3(ceil_div((T > 8 ? T : 8), 256)), dim3(256), 0, stream, steps, group_enabled, 8, norms, T);
        DETRB_CHECK_LAUNCH("adam_prologue_kernel");
    }
    DETRB_LAUNCH(chunk_sumsq_kernel, dim3(nchunks), dim3(256), 0, stream, grads, chunks, norms);
    DETRB_CHECK_LAUNCH("chunk_sumsq_kernel");
    DETRB_LAUNCH(chunk_adam_kernel, dim3(nchunks), dim3(256), 0, stream, params, grads, m, v, chunks, lr_group, lrs, group_enabled, steps, norms,
                                                   clipnorm, beta1, beta2, eps);
    DETRB_CHECK_LAUNCH("chunk_adam_kernel");
    return DETRB_OK;
}

extern "C" int detrb_prep_weight(const float *master, const float *fold, int N, int taps, int Cin,
                                 detrb_bf16 *Wf, int ldf, detrb_bf16 *Wd, int ldd, detrb_stream_t stream)
{
    DETRB_REQUIRE(master && (Wf || Wd) && N > 0 && taps > 0 && Cin > 0, "detrb_prep_weight: bad args");
    DETRB_REQUIRE(!Wf || ldf >= taps * Cin, "detrb_prep_weight: ldf");
    DETRB_REQUIRE(!Wd || ldd >= N, "detrb_prep_weight: ldd");
    int64_t total = (int64_t)N * taps * Cin;
    DETRB_LAUNCH(prep_weight_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, master, fold, N, taps, Cin, (bf16 *)Wf, ldf, (bf16 *)Wd, ldd);
    DETRB_CHECK_LAUNCH("prep_weight_kernel");
    return DETRB_OK;
}
