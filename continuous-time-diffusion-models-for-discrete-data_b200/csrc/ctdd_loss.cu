// loss kernels — placeholder
#include "ctdd_common.cuh"
extern "C" int ctdd_loss_forward(const ctdd_loss_params*, void*) { ctdd::set_error("ctdd_loss_forward: not built"); return 3; }
extern "C" int ctdd_loss_backward(const ctdd_loss_params*, void*) { ctdd::set_error("ctdd_loss_backward: not built"); return 3; }
