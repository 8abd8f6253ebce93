// Tensor-core projection path (tb_linear precision 1). Placeholder until the tcgen05 kernel lands: reports
// TB_ERR_UNSUPPORTED so callers fail loudly instead of silently falling back.
#include "common.cuh"

int tb_linear_tc(const float*, int, const float*, const float*, float*, int, int, int, int, int, const uint8_t*,
                 const float*, int, const uint8_t*, cudaStream_t) {
  return TB_ERR_UNSUPPORTED;
}
