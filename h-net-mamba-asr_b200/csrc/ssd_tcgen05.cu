// tcgen05 SSD kernels (impl 1).  Placeholder entry points until the tensor-core path lands: they fail
// loudly rather than fall back.
#include "common.cuh"
using namespace hnb;

int hnb_ssd_fwd_tc(const void*, const float*, const float*, const float*, int, int, int, int, int, int, void*, void*,
                   void*) {
  set_error("ssd_fwd: tcgen05 implementation not built in this revision");
  return HNB_ERR_UNSUPPORTED;
}
int hnb_ssd_bwd_tc(const void*, const void*, const void*, const float*, const float*, const float*, const void*, int,
                   int, int, int, int, int, void*, float*, float*, float*, float*, void*, void*) {
  set_error("ssd_bwd: tcgen05 implementation not built in this revision");
  return HNB_ERR_UNSUPPORTED;
}
