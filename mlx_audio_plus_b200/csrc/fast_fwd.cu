// placeholder until the specialised kernels land
#include "common.cuh"
namespace b2a {
bool fast_frontend_supported(const b2a_plan*) { return false; }
int fast_frontend_init(b2a_plan*) { return B2A_ERR_UNSUPPORTED; }
void fast_frontend_destroy(b2a_plan*) {}
int fast_frontend_partial(b2a_plan*, const b2a_forward_args*, float*, float*, double*, cudaStream_t) { return B2A_ERR_UNSUPPORTED; }
}
