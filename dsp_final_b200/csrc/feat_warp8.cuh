// feat_warp8.cuh -- placeholder until the warp-autonomous radix-8 kernel lands.
#pragma once
#include "dspx_internal.cuh"
namespace dspx {
inline bool warp8_supported(const dspx_plan *) { return false; }
inline int warp8_prepare(dspx_plan *) { return DSPX_EUNSUPPORTED; }
inline int launch_warp8(const dspx_plan *, const float *, int64_t, int64_t, int64_t, int64_t, float *, float *, cudaStream_t)
{
    set_error("warp8 kernel not built");
    return DSPX_EUNSUPPORTED;
}
}  // namespace dspx
