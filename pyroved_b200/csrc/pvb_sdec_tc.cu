// placeholder until the tcgen05 kernel lands (keeps every pvb.h symbol defined)
#include "pvb_common.cuh"
extern "C" int pvb_has_tcgen05(void) { return 0; }
extern "C" int pvb_sdec_tc_sizes(int64_t, int, pvb_tc_sizes*) { pvb::set_error("tcgen05 path not built"); return -1; }
extern "C" int pvb_sdec_tc_step(const float*, const float*, const float*, const float*, const float*,
                                const float*, const float*, const float*, const float*, float*,
                                float*, float*, float*, int64_t, int64_t, int, int, int, int, int,
                                float, int, void*) { pvb::set_error("tcgen05 path not built"); return -1; }
extern "C" int pvb_sdec_tc_gather_gUv(const float*, float*, int64_t, int, void*) { pvb::set_error("tcgen05 path not built"); return -1; }
