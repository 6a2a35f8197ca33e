// tcgen05 implicit-GEMM path for the dense W->W convolutions (sm_100a).  Under construction: until the
// kernel lands, asking for tensor cores is an error (never a silent fallback).
#include "bsw_common.cuh"
#include "nets.cuh"

extern "C" int bsw_has_tensor_cores(void) { return 0; }

int bsw_model_tc_prepare(bsw_model *m) {
    (void)m;
    bsw_set_error("use_tensor_cores=1: the tcgen05 conv path is not available in this build");
    return BSW_E_INVALID;
}
void bsw_model_tc_release(bsw_model *m) { (void)m; }
int bsw_conv_tc(bsw_model *m, const ConvSlot &c, const ConvArgs &a, int64_t n, cudaStream_t st) {
    (void)m; (void)c; (void)a; (void)n; (void)st;
    bsw_set_error("tcgen05 conv path not available");
    return BSW_E_INVALID;
}
