// Batch descriptor shared by the attention kernels (attention.cu, attention_tma.cu).
#pragma once
#include <cuda_runtime.h>

namespace geob200 {

// One launch covers a batch of independent attention problems (the clouds / pairs of a batched forward): the work units of all
// items form one sequence that the persistent CTAs split evenly (gprefix = running number of (query, 4-key group) units of the
// lanes<->channels kernel, uprefix = running number of query rows of the TMA kernel).  The descriptor travels by value in the
// kernel parameter space (no device allocation, no H2D copy).
constexpr int ATT_MAX_ITEMS = 32;
struct AttItem { const float *q, *k, *v, *qp, *qb, *E; float *out, *S; int N, M; };
struct AttBatch {
    int n_items;
    int reserved;
    long long gprefix[ATT_MAX_ITEMS + 1];
    long long uprefix[ATT_MAX_ITEMS + 1];
    AttItem it[ATT_MAX_ITEMS];
};

int attention_tma_batch(const AttBatch& b, int ldq, int ldk, int ldv, int ldo, int channels, int heads, float div, cudaStream_t st);

}  // namespace geob200
