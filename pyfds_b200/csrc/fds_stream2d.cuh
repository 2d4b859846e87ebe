// Streaming multi-step 2-D kernel (temporal blocking). Filled in after the one-step kernels are
// parity-green; see DESIGN.md.
#pragma once

#include "fds_common.cuh"
