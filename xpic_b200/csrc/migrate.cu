// migrate.cu -- multi-rank (z-slab) particle migration and ghost-cell moments.
// Filled in with the multi-GPU milestone; single-rank runs never reach this file.
#include "comm.cuh"
#include "common.cuh"

namespace xb {

int migrate_and_sort(xb_ctx* c, Species& s, double dt_move)
{
  (void)c; (void)s; (void)dt_move;
  XB_FAIL("multi-rank particle migration is not available in this build");
}

int deposit_ghost_cells(xb_ctx* c, Species& s, double* stage)
{
  (void)c; (void)s; (void)stage;
  XB_FAIL("multi-rank ghost-cell moments are not available in this build");
}

}  // namespace xb
