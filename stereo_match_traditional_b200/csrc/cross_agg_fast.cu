// placeholder, replaced below
#include "common.cuh"
smt_status_t k_aggregate_cross_fast(smt_ctx* ctx, float*, float*, const uint8_t*, const uint16_t*, const uint16_t*, int, int, int, int,
                                    float**) {
    return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "SMT_SUM_FAST not built yet");
}
