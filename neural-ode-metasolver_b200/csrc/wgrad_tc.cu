// placeholder until the tcgen05 wgrad lands
#include "msb_internal.h"
namespace msb {
bool wgrad_tc_supported(ConvShape) { return false; }
int wgrad_tc_nparts(ConvShape) { return 1; }
int launch_wgrad3x3_tc(const __nv_bfloat16*, const __nv_bfloat16*, float*, int*, ConvShape, cudaStream_t) {
    set_error("tcgen05 wgrad not built"); return -1;
}
}
