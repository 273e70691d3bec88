// placeholder until the tcgen05 engine lands (next commit)
#include "msb_internal.h"
namespace msb {
bool tc_shape_supported(int, int, int) { return false; }
size_t tc_packed_weight_bytes(int C) { return (size_t)((C == 64) ? 9 : 9 * (C / 64) * 2) * 128 * 64 * 2; }
int launch_conv3x3_tc(const __nv_bfloat16*, const __nv_bfloat16*, const EpiParams&, ConvShape, cudaStream_t) {
    set_error("tcgen05 engine not built"); return -1;
}
}
