// discretize_tc.cu -- stage 1 on tcgen05 tensor cores (placeholder until the TMA/tcgen05 kernel lands).
#include "discretize.cuh"

namespace sh {

bool discretize_tc_supported(int64_t R, int d, int M)
{
    (void)R; (void)d; (void)M;
    return false;
}

int launch_discretize_tc(const float *, const float *, int64_t, int, int, int64_t *, int64_t, int64_t, int64_t,
                         const DiscWorkspace &, cudaStream_t)
{
    set_error("discretize: tensor-core path not built");
    return 1;
}

}  // namespace sh
