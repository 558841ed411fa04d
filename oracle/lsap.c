/* CPU oracle: rectangular linear-sum-assignment  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Restates the published algorithm of the third-party routine the reference calls at
 * detr_tf/loss/hungarian_matching.py:7,29 -- scipy.optimize.linear_sum_assignment
 * (scipy is unpinned by the reference; the container has 1.18.1): the modified
 * Jonker-Volgenant shortest-augmenting-path method of D. F. Crouse, "On implementing 2D
 * rectangular assignment algorithms", IEEE T-AES 52(4), 2016, as implemented in
 * scipy/optimize/rectangular_lsap.  This C restatement is pinned against the installed
 * scipy binary in tests/test_oracle_cpu.py (random float costs, tie-heavy integer costs and
 * the known-answer table of SURVEY.md Appendix C), and it is the scalar model that the CUDA
 * kernel (csrc/matcher.cu) mirrors lane-for-lane, tie rules included:
 *   - a tall matrix (rows > cols: always the case for DETR's [100 queries, n targets]) is
 *     solved transposed, the n targets being augmented one by one in index order;
 *   - candidate columns are scanned through a `remaining` list initialised in REVERSE order,
 *     a removed entry being replaced by the list's last entry;
 *   - strict '<' keeps the first minimum in scan order, except that on an exact tie an
 *     UNASSIGNED column replaces the incumbent;
 *   - duals / path costs are fp64 on the fp32 costs.
 *
 * int lsap_f32(const float *cost, int nr, int nc, int64_t *rows, int64_t *cols)
 *   cost row-major [nr, nc].  Writes min(nr,nc) pairs with rows ascending (scipy's output
 *   order).  Returns 0, -1 (NaN / -inf entry: scipy raises ValueError) or -2 (infeasible).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int lsap_f32(const float *cost_in, int nr_in, int nc_in, int64_t *rows, int64_t *cols)
{
    if (nr_in == 0 || nc_in == 0) return 0;
    int transpose = nc_in < nr_in;
    int nr = transpose ? nc_in : nr_in;
    int nc = transpose ? nr_in : nc_in;
    double *cost = (double *)malloc(sizeof(double) * (size_t)nr * nc);
    for (int i = 0; i < nr_in; i++)
        for (int j = 0; j < nc_in; j++) {
            double c = (double)cost_in[(size_t)i * nc_in + j];
            if (transpose) cost[(size_t)j * nr_in + i] = c; else cost[(size_t)i * nc_in + j] = c;
        }
    for (size_t k = 0; k < (size_t)nr * nc; k++)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) { free(cost); return -1; }

    double *u = (double *)calloc(nr, sizeof(double));
    double *v = (double *)calloc(nc, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * nc);
    int *path = (int *)malloc(sizeof(int) * nc);
    int *col4row = (int *)malloc(sizeof(int) * nr);
    int *row4col = (int *)malloc(sizeof(int) * nc);
    char *SR = (char *)malloc(nr), *SC = (char *)malloc(nc);
    int *remaining = (int *)malloc(sizeof(int) * nc);
    for (int i = 0; i < nr; i++) col4row[i] = -1;
    for (int j = 0; j < nc; j++) { row4col[j] = -1; path[j] = -1; }
    int rc = 0;

    for (int cur = 0; cur < nr && rc == 0; cur++) {
        double minVal = 0;
        int num_remaining = nc;
        for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;
        for (int i = 0; i < nr; i++) SR[i] = 0;
        for (int j = 0; j < nc; j++) { SC[j] = 0; spc[j] = INFINITY; }
        int sink = -1, i = cur;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; it++) {
                int j = remaining[it];
                double r = minVal + cost[(size_t)i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                    lowest = spc[j];
                    index = it;
                }
            }
            minVal = lowest;
            if (minVal == INFINITY) { rc = -2; break; }
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[cur] += minVal;
        for (int i2 = 0; i2 < nr; i2++)
            if (SR[i2] && i2 != cur) u[i2] += minVal - spc[col4row[i2]];
        for (int j = 0; j < nc; j++)
            if (SC[j]) v[j] -= minVal - spc[j];
        int j = sink;
        for (;;) {
            int i2 = path[j];
            row4col[j] = i2;
            int t = col4row[i2]; col4row[i2] = j; j = t;
            if (i2 == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            /* pairs (col4row[v], v) sorted by col4row[v]: i.e. walk the original rows ascending */
            int k = 0;
            for (int j = 0; j < nc; j++)
                if (row4col[j] != -1) { rows[k] = j; cols[k] = row4col[j]; k++; }
        } else {
            for (int i = 0; i < nr; i++) { rows[i] = i; cols[i] = col4row[i]; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col);
    free(SR); free(SC); free(remaining);
    return rc;
}

/* Batched driver used as the CPU baseline for the matcher microbench: B problems of
 * [Q, n_b] stored with row stride ld (floats); outputs padded to Q entries per problem. */
int lsap_f32_batch(const float *cost, int B, int Q, const int *n, int ld, int64_t *rows, int64_t *cols)
{
    float *tmp = (float *)malloc(sizeof(float) * (size_t)Q * Q);
    for (int b = 0; b < B; b++) {
        const float *c = cost + (size_t)b * Q * ld;
        for (int q = 0; q < Q; q++) for (int t = 0; t < n[b]; t++) tmp[(size_t)q * n[b] + t] = c[(size_t)q * ld + t];
        int rc = lsap_f32(tmp, Q, n[b], rows + (size_t)b * Q, cols + (size_t)b * Q);
        if (rc) { free(tmp); return rc; }
    }
    free(tmp);
    return 0;
}
