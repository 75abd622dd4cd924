// One instantiation unit of the fused sweep per (bit planes, words); built with
// -DPG_P=<planes> -DPG_W=<words> so the variants compile in parallel.
#include "pg_sweep.cuh"
#include "pg_sweep_sym.cuh"

#ifndef PG_P
#error "compile with -DPG_P=<planes> -DPG_W=<words>"
#endif

#define PG_CAT2(a, b, c, d) a##b##c##d
#define PG_NAME(P, W) PG_CAT2(sweep_p, P, _w, W)

namespace pg {
int PG_NAME(PG_P, PG_W)(const SweepParams& prm, const SweepLaunch& l) { return launch_sweep<PG_P, PG_W>(prm, l); }
#define PG_SYM_NAME(P, W) PG_CAT2(sweep_sym_p, P, _w, W)
int PG_SYM_NAME(PG_P, PG_W)(const SymParams& prm, const SymLaunch& l, int* resident) {
  return launch_sweep_sym<PG_P, PG_W>(prm, l, resident);
}
}  // namespace pg
