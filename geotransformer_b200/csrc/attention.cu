// Superpoint transformer kernels: fused multi-head attention (with the geometric structure term), residual
// LayerNorm, row L2-normalisation.
//
// Reference: geotransformer/modules/transformer/rpe_transformer.py:36-103 (self attention with relative positional
// embedding), vanilla_transformer.py:50-101 (cross attention), output_layer.py:15-21 (FFN).
//
// The reference projects the (N,M,C) embedding with proj_p ((N*M, C) x (C, C) GEMM, 9.6 GFLOP per layer and cloud at
// N=271) and then contracts it with q.  Both are linear, so the projection is moved onto q exactly:
//     q_h . (Wp e + bp)_h  =  (Wp_h^T q_h) . e  +  q_h . bp_h
// The host computes qp[n,h,:] = Wp_h^T q[n,h,:] (a tiny GEMM) and this kernel streams E once per layer:
// scores, softmax and P.V never leave the SM.  HBM/L2-bound on the E read (N*M*C*4 bytes per cloud and layer).
#include "common.cuh"
#include "geob200.h"

namespace geob200 {

// R query rows per CTA.  C = channels (multiple of 32, <= 256 handled by 256 threads), H heads (divides 32).
template <int R>
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, const float* __restrict__ qp,
                                                        const float* __restrict__ qb, const float* __restrict__ E, int N, int M,
                                                        int C, int H, float inv_scale_div, float* __restrict__ out) {
    extern __shared__ float sm[];
    float* q_s = sm;                       // [R][C]
    float* qp_s = q_s + R * C;             // [R][H][C]
    float* sc = qp_s + R * H * C;          // [R][H][M]
    const int n0 = blockIdx.x * R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cpl = C / 32;                // contiguous channels per lane
    const int lph = 32 / H;                // lanes per head
    for (int t = threadIdx.x; t < R * C; t += blockDim.x) {
        const int r = t / C, n = n0 + r;
        q_s[t] = (n < N) ? q[(long long)n * C + (t % C)] : 0.f;
    }
    if (qp != nullptr)
        for (int t = threadIdx.x; t < R * H * C; t += blockDim.x) {
            const int r = t / (H * C), n = n0 + r;
            qp_s[t] = (n < N) ? qp[(long long)n * H * C + (t % (H * C))] : 0.f;
        }
    __syncthreads();
    const int my_head = lane / lph;
    for (int m = warp; m < M; m += 8) {
        float kv[8];
#pragma unroll 8
        for (int u = 0; u < 8; ++u) kv[u] = (u < cpl) ? k[(long long)m * C + lane * cpl + u] : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = n0 + r;
            float se = 0.f;
#pragma unroll 8
            for (int u = 0; u < 8; ++u)
                if (u < cpl) se = fmaf(q_s[r * C + lane * cpl + u], kv[u], se);
            for (int o = lph >> 1; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);   // per-head q.k
            float tot = se;
            if (E != nullptr && n < N) {
                const float* erow = E + ((long long)n * M + m) * C + lane * cpl;
                float ev[8];
#pragma unroll 8
                for (int u = 0; u < 8; ++u) ev[u] = (u < cpl) ? erow[u] : 0.f;
                for (int h = 0; h < H; ++h) {
                    float sp = 0.f;
                    const float* w = qp_s + (r * H + h) * C + lane * cpl;
#pragma unroll 8
                    for (int u = 0; u < 8; ++u)
                        if (u < cpl) sp = fmaf(w[u], ev[u], sp);
                    sp = warp_sum(sp);
                    if (h == my_head) tot += sp + qb[(long long)n * H + h];
                }
            }
            if ((lane % lph) == 0) sc[(r * H + my_head) * M + m] = tot / inv_scale_div;
        }
    }
    __syncthreads();
    // softmax over m, one warp per (row, head)
    for (int rh = warp; rh < R * H; rh += 8) {
        float* s = sc + rh * M;
        float mx = -INFINITY;
        for (int m = lane; m < M; m += 32) mx = fmaxf(mx, s[m]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int m = lane; m < M; m += 32) {
            const float e = expf(s[m] - mx);
            s[m] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        for (int m = lane; m < M; m += 32) s[m] = s[m] / sum;
    }
    __syncthreads();
    // out[n][c] = sum_m P[h(c)][m] v[m][c]
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int h = c / (C / H);
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        for (int m = 0; m < M; ++m) {
            const float vv = v[(long long)m * C + c];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(sc[(r * H + h) * M + m], vv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (n0 + r < N) out[(long long)(n0 + r) * C + c] = acc[r];
    }
}

// qb[n][h] = sum_c q[n][h*d + c] * bp[h*d + c]
__global__ void __launch_bounds__(256) head_bias_kernel(const float* __restrict__ q, const float* __restrict__ bp, int N, int C,
                                                        int H, float* __restrict__ qb) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * H) return;
    const int n = t / H, h = t % H, d = C / H;
    float s = 0.f;
    for (int c = 0; c < d; ++c) s = fmaf(q[(long long)n * C + h * d + c], bp[h * d + c], s);
    qb[t] = s;
}

// y = LayerNorm(a + b) * gamma + beta, one warp per row (torch.nn.LayerNorm, eps inside the sqrt, biased variance)
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int N, int C, float eps, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float vals[32];   // C <= 1024
    float s = 0.f;
    int cnt = 0;
    for (int c = lane; c < C; c += 32, ++cnt) {
        const float x = a[(long long)n * C + c] + (b != nullptr ? b[(long long)n * C + c] : 0.f);
        vals[cnt] = x;
        s += x;
    }
    const float mean = warp_sum(s) / (float)C;
    float s2 = 0.f;
    for (int i = 0; i < cnt; ++i) { const float d = vals[i] - mean; s2 = fmaf(d, d, s2); }
    const float rstd = rsqrtf(warp_sum(s2) / (float)C + eps);
    cnt = 0;
    for (int c = lane; c < C; c += 32, ++cnt) y[(long long)n * C + c] = (vals[cnt] - mean) * rstd * gamma[c] + beta[c];
}

// F.normalize(x, p=2, dim=1): x / max(||x||, 1e-12)
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, int N, int C, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = x[(long long)n * C + c]; s = fmaf(v, v, s); }
    const float nrm = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
    for (int c = lane; c < C; c += 32) y[(long long)n * C + c] = x[(long long)n * C + c] / nrm;
}

}  // namespace geob200

using namespace geob200;

extern "C" {

int geob200_attention(const float* q, const float* k, const float* v, const float* qp, const float* qb, const float* embed,
                      int64_t n_query, int64_t n_key, int64_t channels, int64_t heads, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GEOB_REQUIRE(n_query > 0 && n_key > 0, "attention: empty input");
    GEOB_REQUIRE(channels % 32 == 0 && channels <= 256 && heads > 0 && 32 % heads == 0 && channels % heads == 0,
                 "attention: unsupported channels=%lld heads=%lld", (long long)channels, (long long)heads);
    GEOB_REQUIRE((embed == nullptr) == (qp == nullptr) && (embed == nullptr) == (qb == nullptr), "attention: qp/qb/embed must come together");
    constexpr int R = 2;
    const size_t smem = sizeof(float) * (R * channels + R * heads * channels + R * heads * n_key);
    GEOB_REQUIRE(smem <= 200 * 1024, "attention: too many keys (%lld)", (long long)n_key);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        GEOB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const float div = sqrtf((float)(channels / heads));   // d_model_per_head ** 0.5
    attention_kernel<R><<<(unsigned)((n_query + R - 1) / R), 256, smem, st>>>(q, k, v, qp, qb, embed, (int)n_query, (int)n_key,
                                                                             (int)channels, (int)heads, div, out);
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

int geob200_head_bias(const float* q, const float* bias_p, int64_t n, int64_t channels, int64_t heads, float* qb, void* stream) {
    head_bias_kernel<<<(unsigned)((n * heads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q, bias_p, (int)n, (int)channels, (int)heads, qb);
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

int geob200_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, int64_t n, int64_t channels,
                          float eps, float* y, void* stream) {
    GEOB_REQUIRE(channels <= 1024, "add_layernorm: channels > 1024");
    add_layernorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a, b, gamma, beta, (int)n, (int)channels, eps, y);
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

int geob200_l2_normalize(const float* x, int64_t n, int64_t channels, float* y, void* stream) {
    l2_normalize_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, (int)n, (int)channels, y);
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

}  // extern "C"
