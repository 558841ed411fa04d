// pipeline.cu -- the rows either side of the train step, on device:
//   * input normalisation of uint8 frames (data/processing.py:6-23) as a 3x256-entry table lookup, either to the fp32 NHWC
//     tensor the model call takes or fused straight into the bf16 space-to-depth tensor the stem convolution gathers from;
//   * inference post-process (inference.py:68-95): softmax -> max score / argmax label -> background filter with an
//     order-preserving compaction -> box format conversion, for a whole batch in one launch.
// Both are HBM / latency bound byte work: coalesced vector accesses, no tensor cores.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ normalisation
// lut[c*256 + v] = normalised value of byte v in OUTPUT channel c (built on the host with the reference's float64 arithmetic,
// so the result is bit-identical to normalized_images()); swap != 0 reads input channel 2-c (tf_resnet: RGB -> BGR).
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t *__restrict__ img, const float *__restrict__ lut, int swap, float *__restrict__ out, int64_t npix4)
{
    pdl_trigger();
    pdl_wait();
    __shared__ float s_lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = 4 pixels = 12 bytes in, 48 bytes out
    if (i >= npix4) return;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(img) + i * 3;
    uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
    uint8_t b[12];
#pragma unroll
    for (int k = 0; k < 4; k++) { b[k] = (w0 >> (8 * k)) & 0xff; b[4 + k] = (w1 >> (8 * k)) & 0xff; b[8 + k] = (w2 >> (8 * k)) & 0xff; }
    float v[12];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int c = 0; c < 3; c++) v[p * 3 + c] = s_lut[c * 256 + b[p * 3 + (swap ? 2 - c : c)]];
    float4 *dst = reinterpret_cast<float4 *>(out) + i * 3;
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    dst[2] = make_float4(v[8], v[9], v[10], v[11]);
}
// tail pixels (npix % 4) and unaligned bases
__global__ void normalize_u8_tail_kernel(const uint8_t *img, const float *lut, int swap, float *out, int64_t first, int64_t npix)
{
    pdl_trigger();
    pdl_wait();
    int64_t p = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    for (int c = 0; c < 3; c++) out[p * 3 + c] = lut[c * 256 + img[p * 3 + (swap ? 2 - c : c)]];
}

// uint8 NHWC3 frame -> normalised bf16 space-to-depth(2) tensor [B, ceil(H/2), ceil(W/2), 16]: the fp32 image never exists
// (same channel order as image_to_s2d16_kernel: (ry*2+rx)*3 + c, 4 zero channels; pixels outside the frame are 0)
__global__ void __launch_bounds__(256)
image_u8_to_s2d16_kernel(const uint8_t *__restrict__ img, const float *__restrict__ lut, int swap, bf16 *__restrict__ out,
                         int B, int H, int W, int H2, int W2, int pt, int pl, int HP, int WP, long long split)
{
    pdl_trigger();
    pdl_wait();
    __shared__ float s_lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    // one thread per pixel of the zero-padded output [B, HP, WP, 16]; frame pixel (y2, x2) lands at (pt + y2, pl + x2)
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;
    const int yp = blockIdx.y, b = blockIdx.z;
    if (xp >= WP) return;
    const int x2 = xp - pl, y2 = yp - pt;
    const bool inside = x2 >= 0 && x2 < W2 && y2 >= 0 && y2 < H2;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.f;
#pragma unroll
    for (int ry = 0; ry < 2; ry++) {
        const int y = 2 * y2 + ry;
        if (!inside || y >= H) continue;
        const uint8_t *row = img + ((size_t)b * H + y) * (size_t)W * 3;
#pragma unroll
        for (int rx = 0; rx < 2; rx++) {
            const int x = 2 * x2 + rx;
            if (x >= W) continue;
            const uint8_t *px = row + (size_t)x * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) v[(ry * 2 + rx) * 3 + c] = s_lut[c * 256 + px[swap ? 2 - c : c]];
        }
    }
    float lo8[8], hi8[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { lo8[k] = v[k]; hi8[k] = v[8 + k]; }
    const size_t i = ((size_t)b * HP + yp) * WP + xp;
    sp_st8(out + i * 16, split, lo8);
    sp_st8(out + i * 16 + 8, split, hi8);
}

// ------------------------------------------------------------------------------------------ inference post-process
// One CTA per image, one warp per query row at a time.  The reference takes the argmax of the SOFTMAX output (inference.py:
// 73-75), first index on ties: e_i = exp(x_i - max) is computed in fp32 and the (value desc, index asc) order is reduced over
// the warp; score = e_max / sum = 1 / sum.  Kept queries are compacted in ascending query order (tf.where + tf.gather).
constexpr int PP_MAXQ = 1024;

__global__ void __launch_bounds__(128)
postprocess_kernel(const float *__restrict__ logits, int ldl, const float *__restrict__ boxes, int Q, int C, int background_class,
                   int bbox_format, float *__restrict__ out_boxes, int64_t *__restrict__ out_labels, float *__restrict__ out_scores,
                   int32_t *__restrict__ out_query, int32_t *__restrict__ out_count)
{
    pdl_trigger();
    pdl_wait();
    __shared__ int s_label[PP_MAXQ];
    __shared__ float s_score[PP_MAXQ];
    __shared__ int s_warp_total[4];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *lg = logits + (size_t)b * Q * ldl;
    for (int q = warp; q < Q; q += 4) {
        const float *row = lg + (size_t)q * ldl;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
        m = warp_max(m);
        float sum = 0.f, best = -1.f;
        int besti = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            float e = expf(row[c] - m);
            sum += e;
            if (e > best) { best = e; besti = c; }                 // strict '>': keeps the lowest index within a lane
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) { s_label[q] = besti; s_score[q] = __fdiv_rn(best, sum); }
    }
    __syncthreads();
    // order-preserving compaction: warp w owns queries [w*chunk, (w+1)*chunk), ballot scan in 32-query steps
    const int chunk = ((Q + 3) / 4 + 31) / 32 * 32;
    const int q0 = warp * chunk, q1 = min(Q, q0 + chunk);
    int cnt = 0;
    for (int q = q0 + lane; q < q0 + chunk; q += 32) {
        bool keep = q < q1 && s_label[q] != background_class;
        cnt += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) s_warp_total[warp] = cnt;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; w++) base += s_warp_total[w];
    for (int q = q0 + lane; q < q0 + chunk; q += 32) {
        bool keep = q < q1 && s_label[q] != background_class;
        unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int k = base + __popc(mask & ((1u << lane) - 1u));
            const size_t o = (size_t)b * Q + k;
            const float4 bx = reinterpret_cast<const float4 *>(boxes)[(size_t)b * Q + q];
            float4 r = bx;                                             // xy_center: as predicted
            if (bbox_format != 0) {                                    // bbox.py:171-183: corners, clipped to [0,1]
                const float hw = __fmul_rn(bx.z, 0.5f), hh = __fmul_rn(bx.w, 0.5f);
                const float x0 = fminf(fmaxf(__fsub_rn(bx.x, hw), 0.f), 1.f), y0 = fminf(fmaxf(__fsub_rn(bx.y, hh), 0.f), 1.f);
                const float x1 = fminf(fmaxf(__fadd_rn(bx.x, hw), 0.f), 1.f), y1 = fminf(fmaxf(__fadd_rn(bx.y, hh), 0.f), 1.f);
                r = (bbox_format == 1) ? make_float4(x0, y0, x1, y1) : make_float4(y0, x0, y1, x1);
            }
            reinterpret_cast<float4 *>(out_boxes)[o] = r;
            out_labels[o] = s_label[q];
            out_scores[o] = s_score[q];
            if (out_query) out_query[o] = q;
        }
        base += __popc(mask);
    }
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 4; w++) tot += s_warp_total[w];
        out_count[b] = tot;
    }
}


// ------------------------------------------------------------------------------------------ mAP matching
// loss/compute_map.py:183-272 (cal_map, 'box' entries) for one image per warp.  Detections are visited in descending score order
// (stable); each takes the unused ground-truth box of its class with the highest IoU above the threshold (strictly greater, first
// maximum wins: :231-241); lane t runs the greedy pass of IoU threshold t.  IoU in the reference's fp32 operation order
// (compute_iou :104-122: inter / ((area_gt + area_pred) - inter), no FMA contraction), compared as doubles like the reference's
// python floats.
__device__ __forceinline__ float map_iou(const float4 g, float ga, const float4 q, float qa)
{
    const float y1 = fmaxf(g.x, q.x), y2 = fminf(g.z, q.z), x1 = fmaxf(g.y, q.y), x2 = fminf(g.w, q.w);
    const float inter = __fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.f), fmaxf(__fsub_rn(y2, y1), 0.f));
    const float uni = __fsub_rn(__fadd_rn(ga, qa), inter);
    return __fdiv_rn(inter, uni);
}

constexpr int MAPQ = 256, MAPT = 100;
__global__ void __launch_bounds__(32)
map_match_kernel(const float *__restrict__ pred_boxes, const int64_t *__restrict__ pred_labels, const float *__restrict__ pred_scores,
                 const int32_t *__restrict__ pred_count, int Q, const float *__restrict__ t_boxes, const int64_t *__restrict__ t_labels,
                 const int32_t *__restrict__ t_count, int NT, int t_wire, const double *__restrict__ thresholds, int T, int num_classes,
                 int32_t *__restrict__ rank, uint8_t *__restrict__ tp, int32_t *__restrict__ gt_count)
{
    pdl_trigger();
    pdl_wait();
    __shared__ float4 s_pb[MAPQ], s_tb[MAPT];
    __shared__ float s_pa[MAPQ], s_ps[MAPQ], s_ta[MAPT];
    __shared__ int s_pc[MAPQ], s_tc[MAPT], s_order[MAPQ];
    const int b = blockIdx.x, lane = threadIdx.x;
    int cnt = pred_count[b];
    cnt = cnt < 0 ? 0 : (cnt > Q ? Q : cnt);
    int n;
    if (t_wire) {                                   // data/processing.py:35-55: row 0 = [n, 0, 0, 0], rows 1..n = (cx, cy, w, h)
        const float hdr = t_boxes[(size_t)b * NT * 4];
        n = (int)fminf(fmaxf(hdr, 0.f), (float)(NT - 1));
    } else {
        n = t_count[b];
        n = n < 0 ? 0 : (n > NT ? NT : n);
    }
    if (n > MAPT) n = MAPT;
    for (int i = lane; i < cnt; i += 32) {
        const float4 q = *reinterpret_cast<const float4 *>(pred_boxes + ((size_t)b * Q + i) * 4);
        s_pb[i] = q;
        s_pa[i] = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
        s_ps[i] = pred_scores[(size_t)b * Q + i];
        s_pc[i] = (int)pred_labels[(size_t)b * Q + i];
    }
    for (int j = lane; j < n; j += 32) {
        float4 g;
        if (t_wire) {                               // bbox.py:171-183 + :125-138: clipped corners, yx order
            const float4 c = *reinterpret_cast<const float4 *>(t_boxes + ((size_t)b * NT + 1 + j) * 4);
            const float hw = __fmul_rn(c.z, 0.5f), hh = __fmul_rn(c.w, 0.5f);
            const float xmin = fminf(fmaxf(__fsub_rn(c.x, hw), 0.f), 1.f), ymin = fminf(fmaxf(__fsub_rn(c.y, hh), 0.f), 1.f);
            const float xmax = fminf(fmaxf(__fadd_rn(c.x, hw), 0.f), 1.f), ymax = fminf(fmaxf(__fadd_rn(c.y, hh), 0.f), 1.f);
            g = make_float4(ymin, xmin, ymax, xmax);
            s_tc[j] = (int)t_labels[(size_t)b * NT + 1 + j];
        } else {
            g = *reinterpret_cast<const float4 *>(t_boxes + ((size_t)b * NT + j) * 4);
            s_tc[j] = (int)t_labels[(size_t)b * NT + j];
        }
        s_tb[j] = g;
        s_ta[j] = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
        if (gt_count && s_tc[j] >= 0 && s_tc[j] < num_classes) atomicAdd(gt_count + s_tc[j], 1);
    }
    __syncwarp();
    // descending-score order, stable (sorted(range(num_pred), key=lambda i: -score[i]), :211)
    for (int i = lane; i < cnt; i += 32) {
        const float si = s_ps[i];
        int r = 0;
        for (int k = 0; k < cnt; k++) r += (s_ps[k] > si) || (s_ps[k] == si && k < i);
        s_order[r] = i;
        rank[(size_t)b * Q + i] = r;
    }
    for (int i = cnt + lane; i < Q; i += 32) rank[(size_t)b * Q + i] = -1;
    __syncwarp();
    if (lane < T) {
        const double thr = thresholds[lane];
        uint32_t used[4] = {0u, 0u, 0u, 0u};
        uint8_t *out = tp + ((size_t)b * T + lane) * Q;
        for (int pos = 0; pos < cnt; pos++) {
            const int i = s_order[pos], cls = s_pc[i];
            double best = thr;
            int bj = -1;
            for (int j = 0; j < n; j++) {
                if (s_tc[j] != cls || ((used[j >> 5] >> (j & 31)) & 1u)) continue;
                const double v = (double)map_iou(s_tb[j], s_ta[j], s_pb[i], s_pa[i]);
                if (v > best) { best = v; bj = j; }
            }
            if (bj >= 0) used[bj >> 5] |= 1u << (bj & 31);
            out[i] = bj >= 0 ? 1 : 0;
        }
        for (int i = cnt; i < Q; i++) out[i] = 0;
    }
}


// ------------------------------------------------------------------------------------------ resize / augmentation geometry
// data/transformation.py:54-114 (detr_aug_seq: Fliplr, Resize / CropToFixedSize / Affine scale, final Resize to
// config.image_size) collapses into ONE axis-aligned affine map per image; the host samples it and transforms the boxes, this
// kernel resamples the pixels: B ragged uint8 source frames -> one [B, H, W, 3] uint8 batch, bilinear, one thread per output
// pixel.  inv[b] = {ax, bx, ay, by}: source coordinates of the output pixel centre (x + .5, y + .5) are
// xs = ax * (x + .5) + bx - .5, ys = ay * (y + .5) + by - .5 (pixel-centre convention of cv2.resize / imgaug).  Samples outside
// the source frame read 0 (imgaug's constant fill) when zero_border, the nearest edge pixel otherwise (plain resize).
// Explicit round-to-nearest arithmetic (no FMA contraction): the numpy restatement in oracle/ reproduces it bit for bit.
__global__ void __launch_bounds__(256)
resize_affine_u8_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ src_off, const int32_t *__restrict__ src_hw,
                        const float *__restrict__ inv, const uint8_t *__restrict__ zero_border, uint8_t *__restrict__ out, int H, int W)
{
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const int sh = src_hw[2 * b], sw = src_hw[2 * b + 1];
    const uint8_t *img = src + src_off[b];
    const float ax = inv[4 * b], bx = inv[4 * b + 1], ay = inv[4 * b + 2], by = inv[4 * b + 3];
    const float xs = __fsub_rn(__fadd_rn(__fmul_rn(ax, (float)x + 0.5f), bx), 0.5f);
    const float ys = __fsub_rn(__fadd_rn(__fmul_rn(ay, (float)y + 0.5f), by), 0.5f);
    const float x0f = floorf(xs), y0f = floorf(ys);
    const float fx = __fsub_rn(xs, x0f), fy = __fsub_rn(ys, y0f);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const bool zb = zero_border[b] != 0;
    float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            int xx = x0 + dx, yy = y0 + dy;
            const float w = __fmul_rn(dx ? fx : __fsub_rn(1.f, fx), dy ? fy : __fsub_rn(1.f, fy));
            bool inside = xx >= 0 && xx < sw && yy >= 0 && yy < sh;
            if (!inside && zb) continue;
            xx = min(max(xx, 0), sw - 1); yy = min(max(yy, 0), sh - 1);
            const uint8_t *px = img + ((size_t)yy * sw + xx) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) v[c] = __fadd_rn(v[c], __fmul_rn(w, (float)px[c]));
        }
    uint8_t *o = out + (((size_t)b * H + y) * W + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) o[c] = (uint8_t)fminf(fmaxf(rintf(v[c]), 0.f), 255.f);
}

}  // namespace

extern "C" int detrb_normalize_u8(const uint8_t *img, const float *lut, int swap_rb, float *out, int64_t npix, detrb_stream_t stream)
{
    DETRB_REQUIRE(img && lut && out && npix > 0, "detrb_normalize_u8: bad args");
    int64_t n4 = 0;
    if ((((uintptr_t)img) & 3) == 0 && (((uintptr_t)out) & 15) == 0) n4 = npix / 4;
    if (n4 > 0) {
        DETRB_LAUNCH(normalize_u8_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, img, lut, swap_rb, out, n4);
        DETRB_CHECK_LAUNCH("normalize_u8_kernel");
    }
    if (n4 * 4 < npix) {
        int64_t rest = npix - n4 * 4;
        DETRB_LAUNCH(normalize_u8_tail_kernel, dim3((unsigned)((rest + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, img, lut, swap_rb, out, n4 * 4, npix);
        DETRB_CHECK_LAUNCH("normalize_u8_tail_kernel");
    }
    return DETRB_OK;
}

extern "C" int detrb_image_u8_to_s2d16(const uint8_t *img, const float *lut, int swap_rb, detrb_bf16 *out, int B, int H, int W,
                                       int pad_top, int pad_left, int HP, int WP, int64_t split, detrb_stream_t stream)
{
    DETRB_REQUIRE(img && lut && out && B > 0 && H > 0 && W > 0, "detrb_image_u8_to_s2d16: bad args");
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    DETRB_REQUIRE(pad_top >= 0 && pad_left >= 0 && HP >= H2 + pad_top && WP >= W2 + pad_left, "detrb_image_u8_to_s2d16: padded size too small");
    DETRB_REQUIRE(HP <= 65535 && B <= 65535, "detrb_image_u8_to_s2d16: grid too large");
    DETRB_LAUNCH(image_u8_to_s2d16_kernel, dim3((unsigned)ceil_div(WP, 256), (unsigned)HP, (unsigned)B), dim3(256), 0, (cudaStream_t)stream,
                 img, lut, swap_rb, (bf16 *)out, B, H, W, H2, W2, pad_top, pad_left, HP, WP, (long long)split);
    DETRB_CHECK_LAUNCH("image_u8_to_s2d16_kernel");
    return DETRB_OK;
}

extern "C" int detrb_postprocess(const float *logits, int ldl, const float *boxes, int B, int Q, int C, int background_class,
                                 int bbox_format, float *out_boxes, int64_t *out_labels, float *out_scores, int32_t *out_query,
                                 int32_t *out_count, detrb_stream_t stream)
{
    DETRB_REQUIRE(logits && boxes && out_boxes && out_labels && out_scores && out_count, "detrb_postprocess: null pointer");
    DETRB_REQUIRE(B > 0 && Q > 0 && Q <= PP_MAXQ && C > 0 && ldl >= C, "detrb_postprocess: B=%d Q=%d C=%d ldl=%d", B, Q, C, ldl);
    DETRB_REQUIRE(bbox_format >= 0 && bbox_format <= 2, "detrb_postprocess: bbox_format %d (0 xy_center, 1 xyxy, 2 yxyx)", bbox_format);
    DETRB_REQUIRE((((uintptr_t)boxes) & 15) == 0 && (((uintptr_t)out_boxes) & 15) == 0, "detrb_postprocess: boxes must be 16-byte aligned");
    DETRB_LAUNCH(postprocess_kernel, dim3((unsigned)B), dim3(128), 0, (cudaStream_t)stream, logits, ldl, boxes, Q, C, background_class,
                 bbox_format, out_boxes, out_labels, out_scores, out_query, out_count);
    DETRB_CHECK_LAUNCH("postprocess_kernel");
    return DETRB_OK;
}

extern "C" int detrb_map_match(const float *pred_boxes, const int64_t *pred_labels, const float *pred_scores, const int32_t *pred_count,
                               int B, int Q, const float *t_boxes, const int64_t *t_labels, const int32_t *t_count, int NT, int t_wire,
                               const double *thresholds, int T, int num_classes, int32_t *rank, uint8_t *tp, int32_t *gt_count,
                               detrb_stream_t stream)
{
    DETRB_REQUIRE(pred_boxes && pred_labels && pred_scores && pred_count && t_boxes && t_labels && thresholds && rank && tp,
                  "detrb_map_match: null pointer");
    DETRB_REQUIRE(t_wire || t_count, "detrb_map_match: t_count is required unless the targets are in the padded wire format");
    DETRB_REQUIRE(B > 0 && Q > 0 && Q <= MAPQ && NT > 0 && NT <= MAPT + 1 && T > 0 && T <= 32, "detrb_map_match: B=%d Q=%d (<= %d) NT=%d T=%d (<= 32)",
                  B, Q, MAPQ, NT, T);
    DETRB_REQUIRE((((uintptr_t)pred_boxes | (uintptr_t)t_boxes) & 15) == 0, "detrb_map_match: boxes must be 16-byte aligned");
    DETRB_LAUNCH(map_match_kernel, dim3(B), dim3(32), 0, (cudaStream_t)stream, pred_boxes, pred_labels, pred_scores, pred_count, Q, t_boxes,
                 t_labels, t_count, NT, t_wire, thresholds, T, num_classes, rank, tp, gt_count);
    DETRB_CHECK_LAUNCH("map_match_kernel");
    return DETRB_OK;
}

extern "C" int detrb_resize_affine_u8(const uint8_t *src, const int64_t *src_off, const int32_t *src_hw, const float *inv,
                                      const uint8_t *zero_border, uint8_t *out, int B, int H, int W, detrb_stream_t stream)
{
    DETRB_REQUIRE(src && src_off && src_hw && inv && zero_border && out && B > 0 && H > 0 && W > 0, "detrb_resize_affine_u8: bad args");
    DETRB_REQUIRE(H <= 65535 && B <= 65535, "detrb_resize_affine_u8: grid too large");
    DETRB_LAUNCH(resize_affine_u8_kernel, dim3((unsigned)ceil_div(W, 256), (unsigned)H, (unsigned)B), dim3(256), 0, (cudaStream_t)stream,
                 src, src_off, src_hw, inv, zero_border, out, H, W);
    DETRB_CHECK_LAUNCH("resize_affine_u8_kernel");
    return DETRB_OK;
}
