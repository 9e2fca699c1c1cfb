/*
 * TEST INFRASTRUCTURE ONLY -- the oracle.  Nothing under geometric-aware-dense-matching_b200/
 * may import, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Plain-C restatement of the reference's exact 3-D k-nearest-neighbour search
 *   DataProcessing.knn_search            /root/reference/models/RandLA/helper_tool.py:161-170
 *   -> knn_batch(..., omp=True)          .../nearest_neighbors/knn.pyx:71-109
 *   -> cpp_knn_batch_omp                 .../nearest_neighbors/knn_.cxx:104-135
 *   -> nanoflann 1.2.3 KDTree findNeighbors, KNNResultSet::addPoint   nanoflann.hpp:115-139
 *      metric L2_Adaptor::evalMetric     nanoflann.hpp:323-348
 *
 * What is restated exactly:
 *   - the metric: for dim == 3 the 4-wide loop never runs (nanoflann.hpp:331-341) and the tail loop
 *     (:343-346) computes  d2 = ((dx*dx) + (dy*dy)) + (dz*dz), dx = q.x - p.x, sequentially in fp32,
 *     no FMA contraction (the reference builds with -std=c++11 -fopenmp on x86-64 baseline).
 *   - the result: the k smallest d2 in ascending order (KNNResultSet keeps a sorted array).
 * What is canonicalised: nanoflann orders EQUAL distances by KD-tree traversal order and rejects a
 *   candidate equal to the current k-th distance (strict '<', nanoflann.hpp:1361).  That order depends
 *   on the tree, not on the data alone.  This oracle (and the CUDA kernel) order ties by ascending
 *   index: the result is the k lexicographically smallest (d2, index) pairs.  tests/ pin this port
 *   against the compiled reference (oracle/_ref/libref_knn.so): identical rows wherever a row has no
 *   exact fp32 distance tie among its first k+1 candidates, bit-identical sorted d2 vectors everywhere.
 *
 * Also here: the same search in D dims with the dgcnn distance form (models/dgcnn.py:21-27) is NOT
 * restated in C -- its arithmetic is an ATen GEMM; see oracle/dgcnn_oracle.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <float.h>

#if defined(__FMA__)
#error "build the oracle without -mfma / -march=native: the reference does not contract to FMA"
#endif

static inline float d2_ref(const float *q, const float *p) {
  /* nanoflann.hpp:343-346, three iterations of  result += diff0 * diff0  */
  volatile float r = 0.0f; /* volatile: forbid re-association / contraction by the optimiser */
  float d0 = q[0] - p[0];
  r = r + d0 * d0;
  float d1 = q[1] - p[1];
  r = r + d1 * d1;
  float d2 = q[2] - p[2];
  r = r + d2 * d2;
  return r;
}

/* sorted insertion of (d, id) under lexicographic (d, id) order, capacity k (cf. addPoint :115-139) */
static inline void insert_lex(float *bd, int64_t *bi, size_t k, size_t *count, float d, int64_t id) {
  size_t i = *count;
  if (i == k) {
    if (!(d < bd[k - 1] || (d == bd[k - 1] && id < bi[k - 1]))) return;
    i = k - 1;
  } else {
    (*count)++;
  }
  while (i > 0 && (bd[i - 1] > d || (bd[i - 1] == d && bi[i - 1] > id))) {
    bd[i] = bd[i - 1];
    bi[i] = bi[i - 1];
    --i;
  }
  bd[i] = d;
  bi[i] = id;
}

/*
 * support [B, ns, 3], query [B, nq, 3] fp32 row-major  ->  idx [B, nq, k] int64 (as knn.pyx:93),
 * dist2 [B, nq, k] fp32 (may be NULL).  Requires k <= ns (the reference leaves stale ids otherwise,
 * knn_.cxx:120-121).  Returns 0, or -1 on bad arguments.
 */
int oracle_knn_batch(const float *support, size_t B, size_t ns, const float *query, size_t nq,
                     size_t k, int64_t *idx, float *dist2) {
  if (k == 0 || k > ns || k > 64) return -1;
  for (size_t b = 0; b < B; ++b) {
    const float *S = support + b * ns * 3;
    const float *Q = query + b * nq * 3;
#pragma omp parallel for schedule(static)
    for (long qi = 0; qi < (long)nq; ++qi) {
      float bd[64];
      int64_t bi[64];
      size_t count = 0;
      const float *q = Q + (size_t)qi * 3;
      for (size_t j = 0; j < ns; ++j) {
        float d = d2_ref(q, S + j * 3);
        insert_lex(bd, bi, k, &count, d, (int64_t)j);
      }
      int64_t *o = idx + (b * nq + (size_t)qi) * k;
      for (size_t t = 0; t < k; ++t) o[t] = bi[t];
      if (dist2) {
        float *od = dist2 + (b * nq + (size_t)qi) * k;
        for (size_t t = 0; t < k; ++t) od[t] = bd[t];
      }
    }
  }
  return 0;
}

/* d2 of given (query, support-index) pairs with the reference metric: lets tests compare the
 * sorted-distance vectors of the compiled reference's index output.  idx [B, nq, k]. */
int oracle_knn_dist2_of(const float *support, size_t B, size_t ns, const float *query, size_t nq,
                        size_t k, const int64_t *idx, float *dist2) {
  for (size_t b = 0; b < B; ++b)
    for (size_t qi = 0; qi < nq; ++qi)
      for (size_t t = 0; t < k; ++t) {
        int64_t j = idx[(b * nq + qi) * k + t];
        if (j < 0 || (size_t)j >= ns) return -1;
        dist2[(b * nq + qi) * k + t] = d2_ref(query + (b * nq + qi) * 3, support + (b * ns + (size_t)j) * 3);
      }
  return 0;
}

int oracle_knn_abi_version(void) { return 1; }
