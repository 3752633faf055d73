// tcgen05 tensor-core contraction of the geometric structure embedding (placeholder until the kernel lands).
#include "common.cuh"
#include "geob200.h"

int geob200_gse_embed_tc(const float*, const float*, long long, int, const float*, const float*, const float*,
                         const float*, const float*, float*, int, void*, size_t, cudaStream_t) {
    return 1;  // not handled -> caller reports the unsupported mode
}
