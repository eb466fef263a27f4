// placeholder until the tcgen05 attention kernels land (next commit)
#include "common.cuh"
#include "../../include/ctrlv_b200.h"
extern "C" int ctrlv_attn_spatial(const void*, int32_t, int32_t, int32_t, float, void*, void*) {
  ctrlv::set_last_error("attn_spatial: not built yet");
  return CTRLV_ERR_UNSUPPORTED;
}
extern "C" int ctrlv_attn_temporal(const void*, int32_t, int32_t, int32_t, int32_t, float, void*, void*) {
  ctrlv::set_last_error("attn_temporal: not built yet");
  return CTRLV_ERR_UNSUPPORTED;
}
