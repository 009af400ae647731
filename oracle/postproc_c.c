/* Oracle (TEST INFRASTRUCTURE, never linked by balf_b200/): plain-C restatement of the
 * BALF score-map post-processing, fast enough to check the CUDA path at full sizes.
 * Parity: PINNED -- tests/test_oracle_postproc.py checks it against oracle/postproc.py and
 * against tests/golden/postproc_*.npz (written by the reference's own test_utils.py).
 *
 * Reference followed (/root/reference/balf/utils/test_utils.py):
 *   remove_borders :34-47, apply_nms :50-54, find_index_higher_scores :74-95,
 *   get_points_direct_from_score_map :97-112 (threshold), nms_fast :130-168.
 * Tie rule: score descending, raster index ascending (see oracle/postproc.py).
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/liboracle_postproc.so oracle/postproc_c.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int32_t idx; } cand_t;

static int cand_cmp(const void *a, const void *b) {
    const cand_t *p = (const cand_t *)a, *q = (const cand_t *)b;
    if (p->s > q->s) return -1;
    if (p->s < q->s) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

/* test_utils.py:34-47 */
void oracle_remove_borders(const float *in, float *out, int h, int w, int b) {
    memset(out, 0, sizeof(float) * (size_t)h * w);
    for (int y = b; y < h - b; ++y)
        for (int x = b; x < w - b; ++x) out[(size_t)y * w + x] = in[(size_t)y * w + x];
}

/* test_utils.py:50-54; window [i - size/2, i + (size-1)/2] clipped to the map */
void oracle_apply_nms(const float *in, float *out, int h, int w, int size) {
    int lo = size / 2, hi = (size - 1) / 2;
    float *rowmax = (float *)malloc(sizeof(float) * (size_t)h * w);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int a = x - lo < 0 ? 0 : x - lo, e = x + hi >= w ? w - 1 : x + hi;
            float m = in[(size_t)y * w + a];
            for (int k = a + 1; k <= e; ++k) if (in[(size_t)y * w + k] > m) m = in[(size_t)y * w + k];
            rowmax[(size_t)y * w + x] = m;
        }
    for (int y = 0; y < h; ++y) {
        int a = y - lo < 0 ? 0 : y - lo, e = y + hi >= h ? h - 1 : y + hi;
        for (int x = 0; x < w; ++x) {
            float m = rowmax[(size_t)a * w + x];
            for (int k = a + 1; k <= e; ++k) if (rowmax[(size_t)k * w + x] > m) m = rowmax[(size_t)k * w + x];
            float v = in[(size_t)y * w + x];
            out[(size_t)y * w + x] = (v == m) ? v : v * 0.0f;
        }
    }
    free(rowmax);
}

/* test_utils.py:74-95: k-th largest value with the <=0 fallbacks, then the first k raster
 * pixels >= t.  Writes raster indices to out_idx, returns the count, or -1 if h*w < k. */
int oracle_kth_value_topk(const float *map, int h, int w, int k, int32_t *out_idx, float *out_thr) {
    int n = h * w;
    if (k > n || k < 1) return -1;
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (size_t)n);
    for (int i = 0; i < n; ++i) { c[i].s = map[i]; c[i].idx = i; }
    qsort(c, (size_t)n, sizeof(cand_t), cand_cmp);
    float t = c[k - 1].s;
    if (t <= 0.0f) {
        int npos = 0;
        while (npos < n && c[npos].s > 0.0f) ++npos;
        t = npos ? c[npos - 1].s : 0.0f;
    }
    free(c);
    int m = 0;
    for (int i = 0; i < n && m < k; ++i) if (map[i] >= t) out_idx[m++] = i;
    *out_thr = t;
    return m;
}

/* test_utils.py:103 + :130-168: threshold (fp32 compare) then greedy NMS with a
 * (2r+1)^2 exclusion box.  Writes surviving raster indices in (score desc, raster asc)
 * order; returns the count. */
int oracle_greedy_nms(const float *map, int h, int w, float thr, int r, int32_t *out_idx, int max_out) {
    int n = h * w, nc = 0;
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (size_t)n);
    for (int i = 0; i < n; ++i) if (map[i] >= thr) { c[nc].s = map[i]; c[nc].idx = i; ++nc; }
    qsort(c, (size_t)nc, sizeof(cand_t), cand_cmp);
    uint8_t *alive = (uint8_t *)calloc((size_t)n, 1);
    for (int i = 0; i < nc; ++i) alive[c[i].idx] = 1;
    int m = 0;
    for (int i = 0; i < nc; ++i) {
        int idx = c[i].idx;
        if (!alive[idx]) continue;
        int y = idx / w, x = idx % w;
        int y0 = y - r < 0 ? 0 : y - r, y1 = y + r >= h ? h - 1 : y + r;
        int x0 = x - r < 0 ? 0 : x - r, x1 = x + r >= w ? w - 1 : x + r;
        for (int yy = y0; yy <= y1; ++yy) memset(alive + (size_t)yy * w + x0, 0, (size_t)(x1 - x0 + 1));
        if (m < max_out) out_idx[m] = idx;
        ++m;
    }
    free(alive);
    free(c);
    return m;
}
