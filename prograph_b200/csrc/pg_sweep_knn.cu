// Instantiations of the kNN-specialised sweep variants (see pg_sweep_knn.cuh).
#include "pg_sweep_knn.cuh"

namespace pg {

int knn_variant_rows_per_cta(int variant) {
  switch (variant) {
    case 1: case 2: return kConsumers * 2;
    default: return kConsumers;
  }
}

int sweep_knn_variant(int variant, int planes, int words, const SweepParams& prm, size_t list_bytes, cudaStream_t s) {
  if (planes == 5 && words == 8) {
    switch (variant) {
      case 1: return launch_knn_variant<5, 8, 2, 1>(prm, list_bytes, s);
      case 2: return launch_knn_variant<5, 8, 2, 2>(prm, list_bytes, s);
      case 3: return launch_knn_variant<5, 8, 1, 2>(prm, list_bytes, s);
      case 4: return launch_knn_variant<5, 8, 1, 3>(prm, list_bytes, s);
    }
  }
  set_error("no kNN sweep variant %d for planes=%d words=%d", variant, planes, words);
  return PG_ERR_UNSUPPORTED;
}

}  // namespace pg
