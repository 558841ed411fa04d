/* box_math.h -- scalar box / matching arithmetic shared by the CUDA kernels (matcher.cu) and by a host
 * test harness (tests/test_box_math_cpu.py compiles it with gcc, -ffp-contract=off) so the exact formulas
 * the device runs can be checked against the oracle without a GPU.
 *
 * Every product/sum is written through DETRB_MUL/ADD/SUB so that on the device no FMA contraction
 * happens: the fp32 cost matrix then follows the same operation order and rounding as the reference's TF
 * elementwise graph (loss/hungarian_matching.py:172-195, bbox.py:29-105,171-183).
 */
#ifndef DETRB_BOX_MATH_H
#define DETRB_BOX_MATH_H

#ifdef __CUDACC__
#define DETRB_HD __host__ __device__ __forceinline__
#else
#define DETRB_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define DETRB_MUL(a, b) __fmul_rn((a), (b))
#define DETRB_ADD(a, b) __fadd_rn((a), (b))
#define DETRB_SUB(a, b) __fsub_rn((a), (b))
#define DETRB_DIV(a, b) __fdiv_rn((a), (b))
#else
#define DETRB_MUL(a, b) ((a) * (b))
#define DETRB_ADD(a, b) ((a) + (b))
#define DETRB_SUB(a, b) ((a) - (b))
#define DETRB_DIV(a, b) ((a) / (b))
#endif

DETRB_HD float detrb_clip01(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
DETRB_HD float detrb_relu(float x) { return x > 0.f ? x : 0.f; }
DETRB_HD float detrb_min(float a, float b) { return a < b ? a : b; }
DETRB_HD float detrb_max(float a, float b) { return a > b ? a : b; }
DETRB_HD float detrb_abs(float a) { return a < 0.f ? -a : a; }

/* bbox.py:171-183: (cx,cy,w,h) -> clipped (x0,y0,x1,y1) */
DETRB_HD void detrb_to_xyxy(const float *b, float *o)
{
    float hw = DETRB_DIV(b[2], 2.f), hh = DETRB_DIV(b[3], 2.f);
    o[0] = detrb_clip01(DETRB_SUB(b[0], hw));
    o[1] = detrb_clip01(DETRB_SUB(b[1], hh));
    o[2] = detrb_clip01(DETRB_ADD(b[0], hw));
    o[3] = detrb_clip01(DETRB_ADD(b[1], hh));
}

/* GIoU of two clipped xyxy boxes (bbox.py:29-105 + hungarian_matching.py:186-192 / loss.py:84-91) */
DETRB_HD float detrb_giou(const float *a, const float *b)
{
    float iw = detrb_relu(DETRB_SUB(detrb_min(a[2], b[2]), detrb_max(a[0], b[0])));
    float ih = detrb_relu(DETRB_SUB(detrb_min(a[3], b[3]), detrb_max(a[1], b[1])));
    float inter = DETRB_MUL(iw, ih);
    float area_a = DETRB_MUL(DETRB_SUB(a[2], a[0]), DETRB_SUB(a[3], a[1]));
    float area_b = DETRB_MUL(DETRB_SUB(b[2], b[0]), DETRB_SUB(b[3], b[1]));
    float uni = DETRB_SUB(DETRB_ADD(area_a, area_b), inter);
    float iou = DETRB_DIV(inter, uni);
    float cw = detrb_relu(DETRB_SUB(detrb_max(a[2], b[2]), detrb_min(a[0], b[0])));
    float ch = detrb_relu(DETRB_SUB(detrb_max(a[3], b[3]), detrb_min(a[1], b[1])));
    float area = DETRB_MUL(cw, ch);
    return DETRB_SUB(iou, DETRB_DIV(DETRB_SUB(area, uni), area));
}

/* one entry of the Hungarian cost matrix, hungarian_matching.py:178-195:
 *   C = fb * sum|p - t| + fc * (-prob) + fg * (-giou)     evaluated left to right */
DETRB_HD float detrb_match_cost(const float *p_cxcywh, const float *p_xyxy, const float *t_cxcywh,
                                const float *t_xyxy, float prob, float fc, float fb, float fg)
{
    float l1 = DETRB_ADD(DETRB_ADD(DETRB_ADD(detrb_abs(DETRB_SUB(p_cxcywh[0], t_cxcywh[0])),
                                             detrb_abs(DETRB_SUB(p_cxcywh[1], t_cxcywh[1]))),
                                   detrb_abs(DETRB_SUB(p_cxcywh[2], t_cxcywh[2]))),
                         detrb_abs(DETRB_SUB(p_cxcywh[3], t_cxcywh[3])));
    float cg = -detrb_giou(p_xyxy, t_xyxy);
    return DETRB_ADD(DETRB_ADD(DETRB_MUL(fb, l1), DETRB_MUL(fc, -prob)), DETRB_MUL(fg, cg));
}

/* Matched-pair box losses and their gradient wrt the predicted (cx,cy,w,h)  (loss.py:72-96):
 *   l1 = sum|p - t| ; gl = 1 - giou(clip(p), clip(t)) ; grad = w_l1 * d l1 + w_giou * d gl.
 * Clip (tf.clip_by_value) passes gradient where 0 <= x <= 1; relu passes where x > 0; min/max pass to the
 * selected argument (ties are measure-zero and resolved towards the prediction). */
DETRB_HD void detrb_box_loss_grad(const float *p, const float *t, float w_l1, float w_giou,
                                  float *l1_out, float *gl_out, float *grad)
{
    float a[4], b[4];
    detrb_to_xyxy(p, a);
    detrb_to_xyxy(t, b);
    float l1 = 0.f;
    for (int i = 0; i < 4; i++) {
        float d = p[i] - t[i];
        l1 += detrb_abs(d);
        grad[i] = w_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    *l1_out = l1;
    *gl_out = 1.f - detrb_giou(a, b);

    float iw = detrb_min(a[2], b[2]) - detrb_max(a[0], b[0]);
    float ih = detrb_min(a[3], b[3]) - detrb_max(a[1], b[1]);
    float riw = detrb_relu(iw), rih = detrb_relu(ih);
    float I = riw * rih;
    float wa = a[2] - a[0], ha = a[3] - a[1];
    float A = wa * ha + (b[2] - b[0]) * (b[3] - b[1]);
    float U = A - I;
    float cw = detrb_max(a[2], b[2]) - detrb_min(a[0], b[0]);
    float ch = detrb_max(a[3], b[3]) - detrb_min(a[1], b[1]);
    float rcw = detrb_relu(cw), rch = detrb_relu(ch);
    float C = rcw * rch;
    /* giou = I/U - 1 + U/C, U = A - I */
    float dG_dI = A / (U * U) - 1.f / C;
    float dG_dA = -I / (U * U) + 1.f / C;
    float dG_dC = -U / (C * C);
    /* d/d(x0,y0,x1,y1) of the prediction */
    float gx[4];
    float dI_x0 = (iw > 0.f && a[0] >= b[0]) ? -rih : 0.f;
    float dI_x1 = (iw > 0.f && a[2] <= b[2]) ? rih : 0.f;
    float dI_y0 = (ih > 0.f && a[1] >= b[1]) ? -riw : 0.f;
    float dI_y1 = (ih > 0.f && a[3] <= b[3]) ? riw : 0.f;
    float dC_x0 = (cw > 0.f && a[0] <= b[0]) ? -rch : 0.f;
    float dC_x1 = (cw > 0.f && a[2] >= b[2]) ? rch : 0.f;
    float dC_y0 = (ch > 0.f && a[1] <= b[1]) ? -rcw : 0.f;
    float dC_y1 = (ch > 0.f && a[3] >= b[3]) ? rcw : 0.f;
    gx[0] = dG_dI * dI_x0 + dG_dA * (-ha) + dG_dC * dC_x0;
    gx[1] = dG_dI * dI_y0 + dG_dA * (-wa) + dG_dC * dC_y0;
    gx[2] = dG_dI * dI_x1 + dG_dA * (ha) + dG_dC * dC_x1;
    gx[3] = dG_dI * dI_y1 + dG_dA * (wa) + dG_dC * dC_y1;
    /* clip pass-through masks on the raw corners */
    float r0 = p[0] - p[2] / 2.f, r1 = p[1] - p[3] / 2.f, r2 = p[0] + p[2] / 2.f, r3 = p[1] + p[3] / 2.f;
    if (!(r0 >= 0.f && r0 <= 1.f)) gx[0] = 0.f;
    if (!(r1 >= 0.f && r1 <= 1.f)) gx[1] = 0.f;
    if (!(r2 >= 0.f && r2 <= 1.f)) gx[2] = 0.f;
    if (!(r3 >= 0.f && r3 <= 1.f)) gx[3] = 0.f;
    /* loss = 1 - giou */
    grad[0] += -w_giou * (gx[0] + gx[2]);
    grad[1] += -w_giou * (gx[1] + gx[3]);
    grad[2] += -w_giou * 0.5f * (gx[2] - gx[0]);
    grad[3] += -w_giou * 0.5f * (gx[3] - gx[1]);
}

/* ---- LSAP tie rule (scipy rectangular_lsap augmenting_path): candidate = (value, position `it` in the
 * `remaining` scan list, unassigned?).  The sequential scan keeps the first minimum but lets an unassigned
 * column replace an equal incumbent, i.e. among equal values: the LAST unassigned one, else the FIRST.
 * This total order lets a warp pick the same winner with an unordered tree reduction. */
DETRB_HD int detrb_lsap_better(double av, int ait, int aun, double bv, int bit, int bun)
{   /* returns 1 if candidate a beats candidate b; it < 0 means "no candidate" */
    if (bit < 0) return ait >= 0;
    if (ait < 0) return 0;
    if (av != bv) return av < bv;
    if (aun != bun) return aun;
    return aun ? (ait > bit) : (ait < bit);
}

#endif
