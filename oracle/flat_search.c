/* TEST INFRASTRUCTURE ONLY - scalar C restatement of the FLAT top-k search.
 *
 * Second, independent restatement of the arithmetic that the reference reaches
 * through pymilvus -> Milvus Lite (un-vendored third party; call sites
 * /root/reference/milvus/search_embeddings.py:15-22, /root/reference/milvus/RAG.py:383-390).
 * It exists to cross-check oracle/flat_search.py (numpy) and is never linked
 * into or called from the product library.
 *
 *   COSINE: s = <x,q> / (||x|| ||q||)     IP: s = <x,q>       (double accumulate)
 *   order:  (s desc, id asc), size-k insertion list (knowhere keeps a size-k heap)
 *   pad:    id -1, dist -inf when k > n
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int better(double s, int64_t id, double s2, int64_t id2) {
    return s > s2 || (s == s2 && id < id2);
}

int oracle_flat_search(const float* X, const int64_t* ids, int64_t n, int d,
                       const float* Q, int nq, int k, int metric /*0=COSINE,1=IP*/,
                       int64_t* out_ids, float* out_dist) {
    if (d <= 0 || k < 0 || nq < 0 || n < 0) return -1;
    double* best_s = (double*)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
    int64_t* best_i = (int64_t*)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
    if (!best_s || !best_i) return -2;
    for (int qi = 0; qi < nq; ++qi) {
        const float* q = Q + (size_t)qi * d;
        double qq = 0.0;
        for (int j = 0; j < d; ++j) qq += (double)q[j] * (double)q[j];
        double qn = sqrt(qq);
        int cnt = 0;
        for (int64_t r = 0; r < n; ++r) {
            const float* x = X + (size_t)r * d;
            double dot = 0.0, xx = 0.0;
            for (int j = 0; j < d; ++j) {
                dot += (double)x[j] * (double)q[j];
                xx += (double)x[j] * (double)x[j];
            }
            double s = dot;
            if (metric == 0) {
                double den = sqrt(xx) * qn;
                s = den > 0.0 ? dot / den : 0.0;
            }
            if (cnt == k && !(k > 0 && better(s, ids[r], best_s[k - 1], best_i[k - 1]))) continue;
            int pos = cnt < k ? cnt++ : k - 1;
            while (pos > 0 && better(s, ids[r], best_s[pos - 1], best_i[pos - 1])) {
                best_s[pos] = best_s[pos - 1];
                best_i[pos] = best_i[pos - 1];
                --pos;
            }
            best_s[pos] = s;
            best_i[pos] = ids[r];
        }
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)qi * k + j] = j < cnt ? best_i[j] : -1;
            out_dist[(size_t)qi * k + j] = j < cnt ? (float)best_s[j] : -INFINITY;
        }
    }
    free(best_s);
    free(best_i);
    return 0;
}
