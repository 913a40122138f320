// Autoregressive teacher (fastgen) — placeholder entry points while the persistent
// kernel is being brought up; every call fails loudly.
#include "nsw_gemm.cuh"

using namespace nsw;

struct nsw_fastgen {
  int dummy;
};

#define NSW_FG_NOT_YET()                                                          \
  do {                                                                            \
    set_error("fastgen persistent kernel is not part of this build yet");        \
    return NSW_EINVAL;                                                            \
  } while (0)

extern "C" {
int nsw_fastgen_create(const nsw_wavenet_config*, const nsw_tensor*, int32_t, int32_t,
                       nsw_fastgen**) { NSW_FG_NOT_YET(); }
void nsw_fastgen_destroy(nsw_fastgen* h) { delete h; }
int nsw_fastgen_encode_device(nsw_fastgen*, const float*, int32_t, int32_t, float*, void*) {
  NSW_FG_NOT_YET();
}
int nsw_fastgen_encode_host(nsw_fastgen*, const float*, int32_t, int32_t, float*) {
  NSW_FG_NOT_YET();
}
int nsw_fastgen_run_device(nsw_fastgen*, const float*, int32_t, int32_t, const float*, uint64_t,
                           float*, float*, void*) { NSW_FG_NOT_YET(); }
int nsw_fastgen_run_host(nsw_fastgen*, const float*, int32_t, int32_t, const float*, uint64_t,
                         float*, float*) { NSW_FG_NOT_YET(); }
int nsw_fastgen_last_timing(nsw_fastgen*, float*) { NSW_FG_NOT_YET(); }
}
