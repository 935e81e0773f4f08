// TEST INFRASTRUCTURE: NVTX ranges are no-ops in the emulated build (see ../cuda_runtime.h).
#pragma once
static inline int nvtxRangePushA(const char *) { return 0; }
static inline int nvtxRangePop() { return 0; }
