/* CPU oracle for the k-NN step of prod_knn_sample (reference Model.py:82-86,
 * which calls scikit-learn 1.9.0 NearestNeighbors(metric='euclidean').kneighbors).
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_build/.
 *
 * scikit-learn is an un-vendored dependency of the reference (no pinned version;
 * 1.9.0 in this image).  Its published algorithm, restated:
 *   - route 'brute' (width > 15 or k >= N/2; sklearn/neighbors/_base.py:615-648):
 *     float32 rows are upcast to float64, squared distance is computed in the
 *     GEMM form  ||q||^2 + (-2 q.z) + ||z||^2  in float64, clamped at 0
 *     (_argkmin.pyx.tp:492-502), and the k smallest are kept by a bounded
 *     max-heap that rejects on >= (utils/_heap.pyx:46-47) while scanning keys in
 *     ascending index order: among exactly-equal distances the LOWEST indices win.
 *   - route 'kd_tree' (width <= 15): true squared differences in float64.
 * Result order: ascending distance; this oracle breaks exact ties by ascending
 * index (sklearn's own order inside an exact tie is unspecified).
 *
 * excluded[n] != 0 removes key n (the rows drawn as queries, Model.py:83-84).
 * out_idx holds ORIGINAL (uncompacted) key indices, [m][k].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void insert_sorted(double *bd, long *bi, int k, int *cnt, double dist, long idx)
{
    /* keep (bd,bi) sorted ascending by (dist, idx); capacity k */
    int n = *cnt;
    if (n == k) {
        if (dist > bd[k - 1] || (dist == bd[k - 1] && idx > bi[k - 1])) return;
        n = k - 1;
    }
    int p = n;
    while (p > 0 && (bd[p - 1] > dist || (bd[p - 1] == dist && bi[p - 1] > idx))) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
    }
    bd[p] = dist;
    bi[p] = idx;
    *cnt = n + 1;
}

int knn_oracle_f64(const float *keys, long n_keys, int width, const float *queries, long n_queries,
                   const unsigned char *excluded, int k, int gemm_form, long *out_idx, double *out_dist)
{
    double *knorm = (double *)malloc(sizeof(double) * (size_t)n_keys);
    if (!knorm) return 1;
    for (long n = 0; n < n_keys; ++n) {
        double s = 0.0;
        for (int c = 0; c < width; ++c) { double v = keys[n * width + c]; s += v * v; }
        knorm[n] = s;
    }
    int status = 0;
    for (long q = 0; q < n_queries; ++q) {
        double *bd = (double *)malloc(sizeof(double) * (size_t)k);
        long *bi = (long *)malloc(sizeof(long) * (size_t)k);
        int cnt = 0;
        const float *qr = queries + q * width;
        double qn = 0.0;
        for (int c = 0; c < width; ++c) { double v = qr[c]; qn += v * v; }
        for (long n = 0; n < n_keys; ++n) {
            if (excluded && excluded[n]) continue;
            const float *kr = keys + n * width;
            double dist;
            if (gemm_form) {
                double dot = 0.0;
                for (int c = 0; c < width; ++c) dot += (double)qr[c] * (double)kr[c];
                dist = (qn + (-2.0 * dot)) + knorm[n];
                if (dist < 0.0) dist = 0.0;
            } else {
                dist = 0.0;
                for (int c = 0; c < width; ++c) { double t = (double)qr[c] - (double)kr[c]; dist += t * t; }
            }
            insert_sorted(bd, bi, k, &cnt, dist, n);
        }
        if (cnt < k) status = 2;
        for (int j = 0; j < k; ++j) {
            out_idx[q * k + j] = j < cnt ? bi[j] : -1;
            if (out_dist) out_dist[q * k + j] = j < cnt ? bd[j] : INFINITY;
        }
        free(bd);
        free(bi);
    }
    free(knorm);
    return status;
}
