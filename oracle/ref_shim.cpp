// TEST INFRASTRUCTURE ONLY -- not product code.
//
// Two-symbol extern "C" shim over the reference's own native kNN
// (/root/reference/models/RandLA/utils/nearest_neighbors/knn_.cxx:71-135,
//  cpp_knn_batch / cpp_knn_batch_omp).  The reference's Cython layer
// (knn.pyx:71-109) only marshals numpy buffers into these calls, so a ctypes
// call through this shim is the reference's knn_batch().  The reference sources
// are compiled where they lie (see oracle/Makefile); nothing is copied.
#include <cstddef>
#include "knn_.h"

extern "C" {

// mirrors knn.pyx:101-107 (omp != 0 -> cpp_knn_batch_omp, else cpp_knn_batch)
void gadm_ref_knn_batch(const float* pts, size_t batch, size_t npts, size_t dim,
                        const float* queries, size_t nqueries, size_t K,
                        long* out, int omp) {
  if (omp) cpp_knn_batch_omp(pts, batch, npts, dim, queries, nqueries, K, out);
  else     cpp_knn_batch(pts, batch, npts, dim, queries, nqueries, K, out);
}

int gadm_ref_abi_version(void) { return 1; }

}
