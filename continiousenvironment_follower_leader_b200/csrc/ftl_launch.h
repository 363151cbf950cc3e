// ftl_launch.h -- launch wrappers of the per-NB translation units (ftl_step_nb.cu).
#pragma once
#include <cuda_runtime.h>

#include "ftl_device.cuh"

#define FTL_DECLARE_NB(NB)                                                                                            \
    void ftl_launch_kin_nb##NB(const ftl::DevCfg&, const ftl::DevState&, const ftl::DevPool&, const ftl::DevState&,  \
                               const ftl::DevOutputs&, const void*, const ftl::DevOutputs&, double*, int, int, int,  \
                               cudaStream_t);                                                                         \
    void ftl_launch_reset_nb##NB(const ftl::DevCfg&, const ftl::DevState&, const ftl::DevPool&, const uint8_t*,      \
                                 const int*, const ftl::DevOutputs&, int, cudaStream_t);
FTL_DECLARE_NB(0)
FTL_DECLARE_NB(1)
FTL_DECLARE_NB(2)
FTL_DECLARE_NB(3)
FTL_DECLARE_NB(4)
#undef FTL_DECLARE_NB
