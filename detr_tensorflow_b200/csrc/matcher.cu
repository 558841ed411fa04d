// matcher.cu -- Hungarian matcher and set criterion, entirely on device ("the loss never leaves the GPU").
//
//  detrb_matcher : per problem (decoder layer x image) one CTA: softmax statistics, the fp32 cost matrix
//                  (loss/hungarian_matching.py:165-195), then an exact shortest-augmenting-path assignment
//                  run by one warp that mirrors scipy.optimize.linear_sum_assignment (the routine the
//                  reference calls at hungarian_matching.py:29) lane for lane, fp64 duals, tie rules included
//                  -> indices are bit-exact vs scipy on the same cost matrix.
//  detrb_set_loss: weighted CE + L1 + GIoU over the matched pairs with the reference's batch-level
//                  normalisers (loss/loss.py:37-96) and the analytic gradient wrt logits / pre-sigmoid boxes.
#include "common.cuh"
#include "box_math.h"

namespace {

constexpr int MAXQ = 128;     // queries per image supported (reference: 100)
constexpr int MAXT = 100;     // wire format holds at most 99 targets (data/processing.py:49)
constexpr int MAXC = 128;     // classes staged in shared memory (more: read from global memory)

struct Cand { double v; int it; int un; };

__device__ __forceinline__ Cand cand_shfl_xor(const Cand &c, int mask)
{
    Cand o;
    o.v = __shfl_xor_sync(0xffffffffu, c.v, mask);
    o.it = __shfl_xor_sync(0xffffffffu, c.it, mask);
    o.un = __shfl_xor_sync(0xffffffffu, c.un, mask);
    return o;
}

__global__ void __launch_bounds__(128)
matcher_kernel(const float *logits, int ldl, const float *boxes, const float *t_bbox, const int64_t *t_class,
               int B, int Q, int C, float fc, float fb, float fg,
               int64_t *p_indices, int64_t *t_indices, uint8_t *p_selector, int32_t *match, float *cost_out,
               int32_t *status)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: costT[n][Q] f32 | v[Q] f64 | spc[Q] f64 | u[MAXT] f64 | ints...
    double *v = reinterpret_cast<double *>(smem_raw);
    double *spc = v + MAXQ;
    double *u = spc + MAXQ;
    float *costT = reinterpret_cast<float *>(u + MAXT);          // [MAXT][Q]
    float *sP = costT + MAXT * MAXQ;                             // [MAXQ][4] cxcywh
    float *sPxy = sP + MAXQ * 4;
    float *sT = sPxy + MAXQ * 4;                                 // [MAXT][4]
    float *sTxy = sT + MAXT * 4;
    float *smax = sTxy + MAXT * 4;                               // [MAXQ]
    float *ssum = smax + MAXQ;
    int *sTc = reinterpret_cast<int *>(ssum + MAXQ);             // [MAXT]
    int *path = sTc + MAXT;                                      // [MAXQ]
    int *row4col = path + MAXQ;
    int *remaining = row4col + MAXQ;
    int *col4row = remaining + MAXQ;                             // [MAXT]
    int *SC = col4row + MAXT;                                    // [MAXQ]
    int *SR = SC + MAXQ;                                         // [MAXT]
    float *sL = reinterpret_cast<float *>(SR + MAXT);            // [Q][C] logits of this problem (C <= MAXC)
    __shared__ int s_bad;

    const int p = blockIdx.x, b = p % B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *lg = logits + (size_t)p * Q * ldl;
    const float *bx = boxes + (size_t)p * Q * 4;
    const float *tb = t_bbox + (size_t)b * 100 * 4;
    const int64_t *tc = t_class + (size_t)b * 100;
    int n = (int)tb[0];                                   // header row: number of boxes (processing.py:39-43)
    if (n < 0) n = 0;
    if (n > MAXT - 1) n = MAXT - 1;
    if (n > Q) n = Q;
    if (tid == 0) s_bad = 0;
    // ---- the problem's logits -> shared memory in one coalesced sweep (every later read -- softmax statistics, the class
    // probabilities of the cost matrix -- would otherwise be a dependent global round trip on a kernel that is pure latency)
    const bool staged = C <= MAXC;
    if (staged) {
        if (ldl == C && ((reinterpret_cast<uintptr_t>(lg) & 15) == 0) && ((Q * C) & 3) == 0) {
            const float4 *src = reinterpret_cast<const float4 *>(lg);
            float4 *dst = reinterpret_cast<float4 *>(sL);
            for (int i = tid; i < (Q * C) >> 2; i += 128) dst[i] = src[i];
        } else {
            for (int i = tid; i < Q * C; i += 128) { const int q = i / C; sL[i] = lg[(size_t)q * ldl + (i - q * C)]; }
        }
        __syncthreads();
    }
    const float *lgs = staged ? sL : lg;
    const int lds = staged ? C : ldl;

    // ---- stage boxes / targets
    for (int i = tid; i < Q; i += 128) {
        float pb[4] = {bx[i * 4], bx[i * 4 + 1], bx[i * 4 + 2], bx[i * 4 + 3]}, xy[4];
        detrb_to_xyxy(pb, xy);
#pragma unroll
        for (int k = 0; k < 4; k++) { sP[i * 4 + k] = pb[k]; sPxy[i * 4 + k] = xy[k]; }
    }
    for (int i = tid; i < n; i += 128) {
        float t4[4] = {tb[(1 + i) * 4], tb[(1 + i) * 4 + 1], tb[(1 + i) * 4 + 2], tb[(1 + i) * 4 + 3]}, xy[4];
        detrb_to_xyxy(t4, xy);
#pragma unroll
        for (int k = 0; k < 4; k++) { sT[i * 4 + k] = t4[k]; sTxy[i * 4 + k] = xy[k]; }
        sTc[i] = (int)tc[1 + i];
    }
    // ---- softmax statistics per query (tf.nn.softmax, hungarian_matching.py:176)
    for (int q = warp; q < Q; q += 4) {
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lgs[(size_t)q * lds + c]);
        mx = warp_max(mx);
        float sm = 0.f;
        for (int c = lane; c < C; c += 32) sm += expf(lgs[(size_t)q * lds + c] - mx);
        sm = warp_sum(sm);
        if (lane == 0) { smax[q] = mx; ssum[q] = sm; }
    }
    __syncthreads();
    // ---- cost matrix, stored transposed [target][query] (the LSAP below augments targets)
    for (int e = tid; e < n * Q; e += 128) {
        int t = e / Q, q = e - t * Q;
        int cls = sTc[t];
        float prob = 0.f;
        if (cls >= 0 && cls < C) prob = __fdiv_rn(expf(lgs[(size_t)q * lds + cls] - smax[q]), ssum[q]);
        float c = detrb_match_cost(sP + q * 4, sPxy + q * 4, sT + t * 4, sTxy + t * 4, prob, fc, fb, fg);
        costT[t * Q + q] = c;
        if (cost_out) cost_out[((size_t)p * Q + q) * 100 + t] = c;
        if (c != c || c == -INFINITY) s_bad = 1;         // scipy: "matrix contains invalid numeric entries"
    }
    __syncthreads();
    if (warp != 0) return;

    // =========================== LSAP (warp 0) ===========================
    // transposed problem as scipy solves a tall matrix: rows = targets (nr = n), columns = queries (nc = Q)
    const int nr = n, nc = Q;
    int bad = s_bad;
    for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
    __syncwarp();

    for (int cur = 0; cur < nr && !bad; cur++) {
        for (int j = lane; j < nc; j += 32) { remaining[j] = nc - j - 1; SC[j] = 0; spc[j] = INFINITY; }
        for (int i = lane; i < nr; i += 32) SR[i] = 0;
        __syncwarp();
        int num_remaining = nc, sink = -1, i = cur;
        double minVal = 0.0;
        while (sink == -1) {
            if (lane == 0) SR[i] = 1;
            const double ui = u[i];
            const float *crow = costT + i * nc;
            Cand best; best.v = INFINITY; best.it = -1; best.un = 0;
            // (fixed trip count, guarded: the four column visits of a lane are independent until the final compare, and unrolled
            //  their shared-memory loads and fp64 chains overlap -- this loop is the latency of the whole kernel)
#pragma unroll
            for (int k = 0; k < MAXQ / 32; k++) {
                const int it = lane + 32 * k;
                const bool ok = it < num_remaining;
                const int j = ok ? remaining[it] : 0;
                const double r = minVal + (double)crow[j] - ui - v[j];
                double sj = spc[j];
                if (ok && r < sj) { path[j] = i; spc[j] = r; sj = r; }
                const int un = (row4col[j] == -1);
                if (ok && detrb_lsap_better(sj, it, un, best.v, best.it, best.un)) { best.v = sj; best.it = it; best.un = un; }
            }
            // the winner under detrb_lsap_better, without a shuffle tree over (double, int, int) triples: two 32-bit warp minima over the
            // order-preserving integer image of the value, then one warp maximum over a key that encodes the tie rule (among equal
            // values the LAST unassigned position, else the FIRST)
            {
                const double bv = best.v + 0.0;                                        // (-0.0 -> +0.0: equal values, equal images)
                const unsigned long long bits = (unsigned long long)__double_as_longlong(bv);
                const unsigned long long img = bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
                const unsigned hi = best.it >= 0 ? (unsigned)(img >> 32) : 0xFFFFFFFFu, lo = (unsigned)img;
                const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
                const bool in_hi = best.it >= 0 && hi == mhi;
                const unsigned mlo = __reduce_min_sync(0xffffffffu, in_hi ? lo : 0xFFFFFFFFu);
                const bool mine = in_hi && lo == mlo;
                const unsigned key = mine ? (best.un ? (0x20000u + (unsigned)best.it) : (0x10000u + (0xFFFFu - (unsigned)best.it))) : 0u;
                const unsigned mk = __reduce_max_sync(0xffffffffu, key);
                Cand w;
                if (mk == 0u) { w.v = INFINITY; w.it = -1; w.un = 0; }
                else {
                    w.un = mk >= 0x20000u;
                    w.it = w.un ? (int)(mk - 0x20000u) : (int)(0xFFFFu - (mk - 0x10000u));
                    const unsigned long long wimg = ((unsigned long long)mhi << 32) | mlo;
                    const unsigned long long wbits = (wimg >> 63) ? (wimg ^ 0x8000000000000000ull) : ~wimg;
                    w.v = __longlong_as_double((long long)wbits);
                }
                best = w;
            }
            minVal = best.v;
            if (best.it < 0 || minVal == INFINITY) { bad = 2; break; }     // infeasible
            const int index = best.it;
            const int j = remaining[index];
            const int r4c = row4col[j];
            __syncwarp();
            if (r4c == -1) sink = j; else i = r4c;
            if (lane == 0) { SC[j] = 1; remaining[index] = remaining[num_remaining - 1]; }
            num_remaining--;
            __syncwarp();
        }
        if (bad) break;
        // dual updates (before the augmentation, like scipy)
        if (lane == 0) u[cur] += minVal;
        for (int i2 = lane; i2 < nr; i2 += 32)
            if (SR[i2] && i2 != cur) u[i2] += minVal - spc[col4row[i2]];
        for (int j = lane; j < nc; j += 32)
            if (SC[j]) v[j] -= minVal - spc[j];
        __syncwarp();
        if (lane == 0) {
            int j = sink;
            for (;;) {
                int i2 = path[j];
                row4col[j] = i2;
                int tmp = col4row[i2]; col4row[i2] = j; j = tmp;
                if (i2 == cur) break;
            }
        }
        __syncwarp();
    }

    // ---- outputs: match[q], selector, and (query, target) pairs sorted by query (scipy's row order)
    int base = 0;
    for (int q0 = 0; q0 < nc; q0 += 32) {
        int q = q0 + lane;
        int m = (q < nc && !bad) ? row4col[q] : -1;
        if (q < nc) {
            match[(size_t)p * Q + q] = m;
            p_selector[(size_t)p * Q + q] = m >= 0 ? 1 : 0;
        }
        unsigned ball = __ballot_sync(0xffffffffu, m >= 0);
        if (m >= 0) {
            int pos = base + __popc(ball & ((1u << lane) - 1u));
            p_indices[(size_t)p * Q + pos] = q;
            t_indices[(size_t)p * Q + pos] = m;
        }
        base += __popc(ball);
    }
    for (int k = base + lane; k < nc; k += 32) { p_indices[(size_t)p * Q + k] = -1; t_indices[(size_t)p * Q + k] = -1; }
    if (lane == 0) status[p] = bad;
}

// ------------------------------------------------------------------------------------ set criterion
// sums[l][8]: 0 sum w*ce | 1 #neg with argmax==bg | 2 #neg | 3 #pos with argmax!=bg | 4 #pos with argmax==cls
//             5 #pos | 6 sum (1-giou) | 7 sum l1
__global__ void __launch_bounds__(256)
set_loss_kernel(const float *logits, int ldl, const float *boxes, const float *t_bbox, const int64_t *t_class,
                const int32_t *match, int L, int B, int Q, int C, int bg,
                const float *normalisers, float loss_scale, float *sums,
                bf16 *d_logits, int ld_dl, bf16 *d_boxpre, int ld_db, long long split)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= L * B * Q) return;
    const int b = (row / Q) % B, l = row / (Q * B);
    // batch-level normalisers (loss.py:66-67,82,94): functions of the target counts only
    float n_matched, sum_w;
    if (normalisers) { n_matched = normalisers[0]; sum_w = normalisers[1]; }
    else {
        float nb = 0.f;
        for (int i = lane; i < B; i += 32) {
            float h = t_bbox[(size_t)i * 400];
            nb += fminf(fmaxf(h, 0.f), (float)(Q < 99 ? Q : 99));
        }
        nb = warp_sum(nb);
        n_matched = nb;
        sum_w = 0.1f * ((float)(B * Q) - nb) + nb;
    }
    const int m = match[row];
    const int cls = m >= 0 ? (int)t_class[(size_t)b * 100 + 1 + m] : bg;
    const float w = m >= 0 ? 1.0f : 0.1f;
    const float *lg = logits + (size_t)row * ldl;
    // softmax / CE / argmax (first max index, like tf.argmax)
    float x[4]; float mx = -INFINITY; int amax = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int c = lane + 32 * k;
        x[k] = c < C ? lg[c] : -INFINITY;
        if (x[k] > mx) { mx = x[k]; amax = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        int oam = __shfl_xor_sync(0xffffffffu, amax, o);
        if (omx > mx || (omx == mx && oam < amax)) { mx = omx; amax = oam; }
    }
    float e[4], se = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) { e[k] = (lane + 32 * k) < C ? expf(x[k] - mx) : 0.f; se += e[k]; }
    se = warp_sum(se);
    const float lse = mx + logf(se);
    float xc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) if (lane + 32 * k == cls) xc = x[k];
    xc = warp_sum(xc);
    const float ce = lse - xc;
    if (d_logits) {
        const float gs = loss_scale * w / sum_w;          // label_cost weight 1 (loss.py:10)
        bf16 *dl = d_logits + (size_t)row * ld_dl;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int c = lane + 32 * k;
            if (c < ld_dl) {
                float gval = c < C ? gs * (e[k] / se - (c == cls ? 1.f : 0.f)) : 0.f;
                sp_st1(dl + c, split, gval);
            }
        }
    }
    if (lane == 0) {
        float *s = sums + l * 8;
        atomicAdd(s + 0, w * ce);
        if (m < 0) { atomicAdd(s + 2, 1.f); if (amax == bg) atomicAdd(s + 1, 1.f); }
        else {
            atomicAdd(s + 5, 1.f);
            if (amax != bg) atomicAdd(s + 3, 1.f);
            if (amax == cls) atomicAdd(s + 4, 1.f);
        }
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        if (m >= 0) {
            const float *pb = boxes + (size_t)row * 4;
            const float *tb = t_bbox + ((size_t)b * 100 + 1 + m) * 4;
            float pbox[4] = {pb[0], pb[1], pb[2], pb[3]}, tbox[4] = {tb[0], tb[1], tb[2], tb[3]};
            float l1, gl;
            // total = 5*l1_loss + 2*giou_loss, both normalised by the number of matched boxes
            detrb_box_loss_grad(pbox, tbox, 5.f * loss_scale / n_matched, 2.f * loss_scale / n_matched, &l1, &gl, g4);
            atomicAdd(s + 6, gl);
            atomicAdd(s + 7, l1);
#pragma unroll
            for (int k = 0; k < 4; k++) g4[k] *= pbox[k] * (1.f - pbox[k]);      // sigmoid backward (detr.py:188)
        }
        if (d_boxpre) {
            bf16 *db = d_boxpre + (size_t)row * ld_db;
            for (int k = 0; k < ld_db; k++) sp_st1(db + k, split, k < 4 ? g4[k] : 0.f);
        }
    }
}

// clears the loss accumulators.  A kernel, not cudaMemsetAsync: a memset node between two kernels breaks the programmatic-dependent-
// launch chain of the stream (the kernel behind it starts with the full launch latency: 13 us between matcher and loss in the
// step's timeline, tests/trace_step.py)
__global__ void zero_sums_kernel(float *sums, int n)
{
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < n; i += blockDim.x) sums[i] = 0.f;
}

__global__ void set_loss_finalize_kernel(const float *sums, const float *t_bbox, int L, int B, int Q,
                                         const float *normalisers, float loss_scale,
                                         float *losses, float *total, const int32_t *status, int nstatus)
{
    pdl_trigger();      // let the next kernel of the stream become resident
    pdl_wait();         // predecessor complete, its writes visible
    // a matcher problem with NaN / -inf costs (status != 0): the reference raises through scipy; here the result is poisoned
    int bad = 0;
    if (status) for (int i = threadIdx.x; i < nstatus; i += 32) bad |= status[i];
    bad = __any_sync(0xffffffffu, bad != 0);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float n_matched, sum_w;
    if (normalisers) { n_matched = normalisers[0]; sum_w = normalisers[1]; }
    else {
        float nb = 0.f;
        for (int i = 0; i < B; i++) nb += fminf(fmaxf(t_bbox[(size_t)i * 400], 0.f), (float)(Q < 99 ? Q : 99));
        n_matched = nb;
        sum_w = 0.1f * ((float)(B * Q) - nb) + nb;
    }
    float tot = 0.f;
    for (int l = 0; l < L; l++) {
        const float *s = sums + l * 8;
        float *o = losses + l * 6;
        o[0] = s[0] / sum_w;             // label_cost
        o[1] = s[1] / s[2];              // true_neg
        o[2] = s[3] / s[5];              // true_pos
        o[3] = s[4] / s[5];              // pos_accuracy
        o[4] = s[6] / n_matched;         // giou_loss
        o[5] = s[7] / n_matched;         // l1_loss
        tot += 1.f * o[0] + 2.f * o[4] + 5.f * o[5];
        if (bad) for (int k = 0; k < 6; k++) o[k] = __int_as_float(0x7fc00000);
    }
    *total = bad ? __int_as_float(0x7fc00000) : tot * loss_scale;
}

static_assert(MAXQ % 32 == 0, "the LSAP scan visits MAXQ / 32 columns per lane");
constexpr size_t matcher_smem_bytes()
{
    return sizeof(double) * (MAXQ + MAXQ + MAXT) + sizeof(float) * (MAXT * MAXQ + MAXQ * 8 + MAXT * 8 + MAXQ * 2) +
           sizeof(int) * (MAXT + MAXQ * 3 + MAXT + MAXQ + MAXT);
}

}  // namespace

extern "C" int detrb_matcher(const float *logits, int ldl, const float *boxes, const float *t_bbox, const int64_t *t_class,
                             int P, int B, int Q, int C, float fcost_class, float fcost_bbox, float fcost_giou,
                             int64_t *p_indices, int64_t *t_indices, uint8_t *p_selector, int32_t *match,
                             float *cost, int32_t *status, detrb_stream_t stream)
{
    DETRB_REQUIRE(logits && boxes && t_bbox && t_class && p_indices && t_indices && p_selector && match && status,
                  "detrb_matcher: null pointer");
    DETRB_REQUIRE(P > 0 && B > 0 && P % B == 0, "detrb_matcher: P=%d must be a positive multiple of B=%d", P, B);
    DETRB_REQUIRE(Q > 0 && Q <= MAXQ, "detrb_matcher: Q=%d out of range (1..%d)", Q, MAXQ);
    DETRB_REQUIRE(C > 0 && ldl >= C, "detrb_matcher: C=%d ldl=%d", C, ldl);
    // + the staged logits [Q][C] (C <= MAXC; 36.8 KB at 100 x 92: two CTAs per SM, 256 problems in one wave)
    const size_t smem = matcher_smem_bytes() + (C <= MAXC ? sizeof(float) * (size_t)Q * C : 0);
    static bool configured[64] = {false};          // the opt-in is per device
    int dev = 0;
    DETRB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        DETRB_CUDA(cudaFuncSetAttribute(matcher_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(matcher_smem_bytes() + sizeof(float) * MAXQ * MAXC)));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    DETRB_LAUNCH(matcher_kernel, dim3(P), dim3(128), smem, (cudaStream_t)stream, logits, ldl, boxes, t_bbox, t_class, B, Q, C,
                                                          fcost_class, fcost_bbox, fcost_giou,
                                                          p_indices, t_indices, p_selector, match, cost, status);
    DETRB_CHECK_LAUNCH("matcher_kernel");
    return DETRB_OK;
}

extern "C" int detrb_set_loss(const float *logits, int ldl, const float *boxes, const float *t_bbox, const int64_t *t_class,
                              const int32_t *match, int L, int B, int Q, int C, int background_class,
                              const float *normalisers, float loss_scale, float *sums, float *losses, float *total,
                              detrb_bf16 *d_logits, int ld_dl, detrb_bf16 *d_boxpre, int ld_db, const int32_t *status, int64_t split,
                              detrb_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    DETRB_REQUIRE(logits && boxes && t_bbox && t_class && match && sums && losses && total, "detrb_set_loss: null pointer");
    DETRB_REQUIRE(L > 0 && B > 0 && Q > 0 && C > 0 && C <= 128 && ldl >= C, "detrb_set_loss: bad sizes");
    DETRB_REQUIRE(!d_logits || (ld_dl >= C && ld_dl <= 128), "detrb_set_loss: ld_dl=%d", ld_dl);
    DETRB_REQUIRE(!d_boxpre || (ld_db >= 4 && ld_db <= 64), "detrb_set_loss: ld_db=%d", ld_db);
    DETRB_LAUNCH(zero_sums_kernel, dim3(1), dim3(64), 0, stream, sums, 8 * L);
    DETRB_CHECK_LAUNCH("zero_sums_kernel");
    int rows = L * B * Q;
    DETRB_LAUNCH(set_loss_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, stream, logits, ldl, boxes, t_bbox, t_class, match, L, B, Q, C, background_class,
                                                           normalisers, loss_scale, sums, (bf16 *)d_logits, ld_dl, (bf16 *)d_boxpre, ld_db, (long long)split);
    DETRB_CHECK_LAUNCH("set_loss_kernel");
    DETRB_LAUNCH(set_loss_finalize_kernel, dim3(1), dim3(32), 0, stream, sums, t_bbox, L, B, Q, normalisers, loss_scale, losses, total, status, L * B);
    DETRB_CHECK_LAUNCH("set_loss_finalize_kernel");
    return DETRB_OK;
}
