// tcgen05 reverse-step path (S == 256) — placeholder until the tensor kernel lands; AUTO falls to SIMT.
#include "ctdd_common.cuh"
namespace ctdd {
bool tc_supports(const ctdd_step_params*) { return false; }
long long tc_workspace_bytes(long long, int) { return 0; }
int launch_step_tc(const ctdd_step_params*, cudaStream_t) { set_error("tcgen05 path not built"); return 3; }
}  // namespace ctdd
extern "C" int64_t ctdd_tc_tables_bytes(int) { return 0; }
extern "C" int ctdd_prep_tc_tables(const float*, const float*, const float*, int, int, float, int, void*, void*) {
  ctdd::set_error("ctdd_prep_tc_tables: tcgen05 path not built");
  return 3;
}
