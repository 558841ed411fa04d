/* Host harness: compiles the product's shared scalar header (csrc/box_math.h) with gcc so that the exact
 * formulas the CUDA kernels run (cost entries, box-loss gradient, LSAP tie rule) are testable without a GPU.
 * lsap_warp_model() re-enacts matcher.cu's warp algorithm (32 strided lanes + xor-butterfly argmin) serially. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../detr_tensorflow_b200/csrc/box_math.h"

void h_cost_matrix(const float *p_bbox, const float *probs, int Q, int C, const float *t_bbox, const int64_t *t_cls,
                   int n, float fc, float fb, float fg, float *out /* [Q][n] */)
{
    for (int q = 0; q < Q; q++) {
        float pxy[4]; detrb_to_xyxy(p_bbox + q * 4, pxy);
        for (int t = 0; t < n; t++) {
            float txy[4]; detrb_to_xyxy(t_bbox + t * 4, txy);
            out[q * n + t] = detrb_match_cost(p_bbox + q * 4, pxy, t_bbox + t * 4, txy, probs[q * C + t_cls[t]], fc, fb, fg);
        }
    }
}

void h_box_loss_grad(const float *p, const float *t, float w_l1, float w_giou, float *l1, float *gl, float *grad)
{
    detrb_box_loss_grad(p, t, w_l1, w_giou, l1, gl, grad);
}

typedef struct { double v; int it; int un; } Cand;

/* costT [nr][nc] (targets x queries), nr <= nc.  row4col[nc] out.  returns 0 / 2 (infeasible) */
int lsap_warp_model(const float *costT, int nr, int nc, int *row4col)
{
    double *u = calloc(nr > 0 ? nr : 1, sizeof(double)), *v = calloc(nc, sizeof(double)), *spc = malloc(sizeof(double) * nc);
    int *path = malloc(sizeof(int) * nc), *remaining = malloc(sizeof(int) * nc), *col4row = malloc(sizeof(int) * (nr > 0 ? nr : 1));
    int *SC = malloc(sizeof(int) * nc), *SR = malloc(sizeof(int) * (nr > 0 ? nr : 1));
    for (int j = 0; j < nc; j++) { row4col[j] = -1; path[j] = -1; }
    for (int i = 0; i < nr; i++) col4row[i] = -1;
    int bad = 0;
    for (int cur = 0; cur < nr && !bad; cur++) {
        for (int j = 0; j < nc; j++) { remaining[j] = nc - j - 1; SC[j] = 0; spc[j] = INFINITY; }
        for (int i = 0; i < nr; i++) SR[i] = 0;
        int num_remaining = nc, sink = -1, i = cur;
        double minVal = 0.0;
        while (sink == -1) {
            SR[i] = 1;
            Cand best[32];
            for (int lane = 0; lane < 32; lane++) {
                Cand b = {INFINITY, -1, 0};
                for (int it = lane; it < num_remaining; it += 32) {
                    int j = remaining[it];
                    double r = minVal + (double)costT[i * nc + j] - u[i] - v[j];
                    if (r < spc[j]) { path[j] = i; spc[j] = r; }
                    Cand c = {spc[j], it, row4col[j] == -1};
                    if (detrb_lsap_better(c.v, c.it, c.un, b.v, b.it, b.un)) b = c;
                }
                best[lane] = b;
            }
            for (int o = 16; o > 0; o >>= 1) {
                Cand nb[32];
                for (int lane = 0; lane < 32; lane++) {
                    Cand other = best[lane ^ o];
                    nb[lane] = detrb_lsap_better(other.v, other.it, other.un, best[lane].v, best[lane].it, best[lane].un) ? other : best[lane];
                }
                memcpy(best, nb, sizeof(nb));
            }
            for (int lane = 1; lane < 32; lane++) if (best[lane].it != best[0].it) { bad = 3; }
            minVal = best[0].v;
            if (best[0].it < 0 || minVal == INFINITY) { bad = 2; break; }
            int index = best[0].it, j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1; remaining[index] = remaining[num_remaining - 1]; num_remaining--;
        }
        if (bad) break;
        u[cur] += minVal;
        for (int i2 = 0; i2 < nr; i2++) if (SR[i2] && i2 != cur) u[i2] += minVal - spc[col4row[i2]];
        for (int j = 0; j < nc; j++) if (SC[j]) v[j] -= minVal - spc[j];
        int j = sink;
        for (;;) { int i2 = path[j]; row4col[j] = i2; int t = col4row[i2]; col4row[i2] = j; j = t; if (i2 == cur) break; }
    }
    free(u); free(v); free(spc); free(path); free(remaining); free(col4row); free(SC); free(SR);
    return bad;
}
