// Coarse (superpoint) matching, patch gathering + fine matching scores, log-domain Sinkhorn with dustbins.
//
// Reference:
//   geotransformer/modules/geotransformer/superpoint_matching.py:13-50
//   experiments/*/model.py:105-108,169-188 (patch gathers and the 'bnd,bmd->bnm' einsum / sqrt(C))
//   geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66 (100 Python-loop iterations, ~600 launches)
// Here the 100 Sinkhorn iterations of one patch pair run inside one CTA with the (K+1)x(K+1) score matrix in
// shared memory; nothing but the input scores and the final log-assignment touches HBM.
#include "common.cuh"
#include "geob200.h"

namespace geob200 {

// ---- superpoint matching ---------------------------------------------------------------------------------

// valid (non-empty) node lists, in index order (torch.nonzero): single CTA, chunked ordered compaction
__global__ void __launch_bounds__(1024) compact_masks_kernel(const unsigned char* __restrict__ masks, int n, int* __restrict__ idx,
                                                             int* __restrict__ count) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int f = (i < n && masks[i]) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        const int pre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (f) idx[off + pre] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += warp_tot[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

// S[i][j] = exp(-clamp(2 - 2 <fr_i, fs_j>, 0)) over the valid nodes; row sums.  One CTA per (compacted) row.
__global__ void __launch_bounds__(256) spm_scores_kernel(const float* __restrict__ fr, const float* __restrict__ fs, int C,
                                                         const int* __restrict__ ridx, const int* __restrict__ rcount,
                                                         const int* __restrict__ sidx, const int* __restrict__ scount,
                                                         int ld, float* __restrict__ S, float* __restrict__ rowsum) {
    extern __shared__ float sm[];
    float* a = sm;            // [C]
    float* srow = sm + C;     // [ld]
    const int i = blockIdx.x;
    if (i >= *rcount) return;
    const int ns = *scount;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) a[c] = fr[(long long)ridx[i] * C + c];
    __syncthreads();
    for (int j = warp; j < ns; j += 8) {
        const float* b = fs + (long long)sidx[j] * C;
        float d = 0.f;
        for (int c = lane; c < C; c += 32) d = fmaf(a[c], b[c], d);
        d = warp_sum(d);
        if (lane == 0) {
            const float v = expf(-fmaxf(2.0f - 2.0f * d, 0.0f));
            srow[j] = v;
            S[(long long)i * ld + j] = v;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
        for (int j = lane; j < ns; j += 32) s += srow[j];
        s = warp_sum(s);
        if (lane == 0) rowsum[i] = s;
    }
}

__global__ void __launch_bounds__(256) spm_colsum_kernel(const float* __restrict__ S, int ld, const int* __restrict__ rcount,
                                                         const int* __restrict__ scount, float* __restrict__ colsum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *scount) return;
    const int nr = *rcount;
    float s = 0.f;
    for (int i = 0; i < nr; ++i) s += S[(long long)i * ld + j];
    colsum[j] = s;
}

// dual normalisation into a dense flat (nr*ns) array
__global__ void __launch_bounds__(256) spm_dual_kernel(const float* __restrict__ S, int ld, const int* __restrict__ rcount,
                                                       const int* __restrict__ scount, const float* __restrict__ rowsum,
                                                       const float* __restrict__ colsum, int dual, float* __restrict__ flat) {
    const int nr = *rcount, ns = *scount;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nr * ns) return;
    const int i = (int)(t / ns), j = (int)(t % ns);
    const float v = S[(long long)i * ld + j];
    flat[t] = dual ? (v / rowsum[i]) * (v / colsum[j]) : v;
}

// top-k (largest) of a flat positive array: MSB-first 8-bit radix select of the k-th value, then an ordered
// sweep that keeps everything above the threshold and the lowest-index ties, then a bitonic sort of the k winners.
template <int KMAX>
__global__ void __launch_bounds__(1024) topk_flat_kernel(const float* __restrict__ flat, const int* __restrict__ rcount,
                                                         const int* __restrict__ scount, int k_req, const int* __restrict__ ridx,
                                                         const int* __restrict__ sidx, long long* __restrict__ ref_out,
                                                         long long* __restrict__ src_out, float* __restrict__ score_out,
                                                         int* __restrict__ k_out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned long long keys[KMAX];
    __shared__ unsigned prefix_s, kth_s;
    __shared__ int warp_tot[32], carry_gt, carry_eq;
    const int nr = *rcount, ns = *scount;
    const long long total = (long long)nr * ns;
    const int k = (int)min((long long)min(k_req, KMAX), total);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rows past the number of existing candidates are padding: index -1 (gather_patches turns it into an empty patch)
    for (int i = k + threadIdx.x; i < k_req; i += blockDim.x) { ref_out[i] = -1; src_out[i] = -1; score_out[i] = 0.f; }
    if (k == 0) { if (threadIdx.x == 0) *k_out = 0; return; }
    unsigned prefix = 0, mask = 0;
    int remaining = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (long long t = threadIdx.x; t < total; t += blockDim.x) {
            const unsigned b = __float_as_uint(flat[t]);
            if ((b & mask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int rem = remaining;
            int d = 255;
            for (; d > 0; --d) {
                if ((int)hist[d] >= rem) break;
                rem -= (int)hist[d];
            }
            prefix_s = prefix | ((unsigned)d << shift);
            kth_s = (unsigned)rem;
        }
        __syncthreads();
        prefix = prefix_s;
        remaining = (int)kth_s;
        mask |= (255u << shift);
        __syncthreads();
    }
    const unsigned thr = prefix;          // bit pattern of the k-th largest value
    const int need_eq = remaining;        // how many ties at the threshold are kept (lowest flat indices)
    const int n_gt = k - need_eq;
    if (threadIdx.x == 0) { carry_gt = 0; carry_eq = 0; }
    __syncthreads();
    for (long long base = 0; base < total; base += 1024) {
        const long long t = base + threadIdx.x;
        unsigned b = 0;
        if (t < total) b = __float_as_uint(flat[t]);
        const int fg = (t < total && b > thr) ? 1 : 0;
        const int fe = (t < total && b == thr) ? 1 : 0;
        const unsigned bg = __ballot_sync(0xffffffffu, fg), be = __ballot_sync(0xffffffffu, fe);
        const int pg = __popc(bg & ((1u << lane) - 1u)), pe = __popc(be & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = (__popc(bg) << 16) | __popc(be);
        __syncthreads();
        int og = carry_gt, oe = carry_eq;
        for (int w = 0; w < warp; ++w) { og += warp_tot[w] >> 16; oe += warp_tot[w] & 0xFFFF; }
        const unsigned long long key = ((unsigned long long)b << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)t);
        if (fg) keys[og + pg] = key;
        if (fe && oe + pe < need_eq) keys[n_gt + oe + pe] = key;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tg = 0, te = 0;
            for (int w = 0; w < 32; ++w) { tg += warp_tot[w] >> 16; te += warp_tot[w] & 0xFFFF; }
            carry_gt += tg; carry_eq += te;
        }
        __syncthreads();
    }
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    for (int i = k + threadIdx.x; i < n2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const unsigned long long a = keys[t], b = keys[p];
                    const bool desc = ((t & kk) == 0);
                    if ((a < b) == desc) { keys[t] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const unsigned long long key = keys[i];
        const unsigned t = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
        ref_out[i] = ridx[t / ns];
        src_out[i] = sidx[t % ns];
        score_out[i] = __uint_as_float((unsigned)(key >> 32));
    }
    if (threadIdx.x == 0) *k_out = k;
}

// ---- patch gathers ---------------------------------------------------------------------------------------
// out_idx[p][i] = knn[corr[p]][i]; out_mask likewise; out_pts = padded_points[idx]
__global__ void __launch_bounds__(256) gather_patches_kernel(const long long* __restrict__ corr, int P, const long long* __restrict__ knn,
                                                             const unsigned char* __restrict__ knn_masks, int K,
                                                             const float* __restrict__ pts, int n_pts,
                                                             long long* __restrict__ out_idx, unsigned char* __restrict__ out_mask,
                                                             float* __restrict__ out_pts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * K) return;
    const int p = t / K, i = t % K;
    const long long node = corr[p];
    const long long idx = node >= 0 ? knn[node * K + i] : (long long)n_pts;
    out_idx[t] = idx;
    out_mask[t] = node >= 0 ? knn_masks[node * K + i] : 0;
    const bool ok = idx < n_pts;
    out_pts[3 * t + 0] = ok ? pts[3 * idx + 0] : 0.f;
    out_pts[3 * t + 1] = ok ? pts[3 * idx + 1] : 0.f;
    out_pts[3 * t + 2] = ok ? pts[3 * idx + 2] : 0.f;
}

// scores[p][i][j] = <fr[ridx[p][i]], fs[sidx[p][j]]> / sqrt(C); rows of the sentinel index are zero.
// One CTA per patch; T = K/16 outputs per thread per dimension.
template <int T>
__global__ void __launch_bounds__(256) patch_scores_kernel(const float* __restrict__ fr, int nr, const float* __restrict__ fs, int ns,
                                                           int C, const long long* __restrict__ ridx, const long long* __restrict__ sidx,
                                                           float inv_div, float* __restrict__ out) {
    constexpr int K = 16 * T, CH = 32;
    __shared__ float A[K][CH + 1], B[K][CH + 1];
    const int p = blockIdx.x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[T][T];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) acc[a][b] = 0.f;
    for (int c0 = 0; c0 < C; c0 += CH) {
        for (int e = threadIdx.x; e < K * CH; e += 256) {
            const int r = e / CH, c = e % CH;
            const long long ia = ridx[(long long)p * K + r], ib = sidx[(long long)p * K + r];
            A[r][c] = (ia < nr && c0 + c < C) ? fr[ia * C + c0 + c] : 0.f;
            B[r][c] = (ib < ns && c0 + c < C) ? fs[ib * C + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < CH; ++c) {
            float av[T], bv[T];
#pragma unroll
            for (int a = 0; a < T; ++a) av[a] = A[ty + 16 * a][c];
#pragma unroll
            for (int b = 0; b < T; ++b) bv[b] = B[tx + 16 * b][c];
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b)
            out[((long long)p * K + ty + 16 * a) * K + tx + 16 * b] = acc[a][b] / inv_div;
}

// ---- Sinkhorn --------------------------------------------------------------------------------------------
// One CTA per patch pair.  Z (K+1)x(K+1) padded scores in shared memory, row/col potentials u, v.
// learnable_sinkhorn.py:13-18:  u = log_mu - LSE_j(Z + v) ; v = log_nu - LSE_i(Z + u)   x num_iterations
__global__ void __launch_bounds__(1024) sinkhorn_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ row_masks,
                                                       const unsigned char* __restrict__ col_masks, const float* __restrict__ alpha_p,
                                                       int K, int iters, float inf, float* __restrict__ out) {
    extern __shared__ float sm[];
    const int K1 = K + 1;
    const int ld = K1 | 1;                   // odd row stride: conflict-free column walks
    float* Z = sm;                           // [K1][ld]
    float* u = Z + K1 * ld;
    float* v = u + K1;
    float* lmu = v + K1;
    float* lnu = lmu + K1;
    __shared__ float norm_s;
    __shared__ int cnt_s[2];
    const int p = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float alpha = *alpha_p;
    if (threadIdx.x < 2) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    {
        int cr = 0, cc = 0;
        for (int i = threadIdx.x; i < K; i += blockDim.x) { cr += row_masks[(long long)p * K + i] ? 1 : 0; cc += col_masks[(long long)p * K + i] ? 1 : 0; }
        atomicAdd(&cnt_s[0], cr);
        atomicAdd(&cnt_s[1], cc);
    }
    __syncthreads();
    const float nvr = (float)cnt_s[0], nvc = (float)cnt_s[1];
    if (threadIdx.x == 0) norm_s = -logf(nvr + nvc);
    __syncthreads();
    const float norm = norm_s;
    for (int e = threadIdx.x; e < K1 * K1; e += blockDim.x) {
        const int i = e / K1, j = e % K1;
        float z = (i < K && j < K) ? scores[((long long)p * K + i) * K + j] : alpha;
        const bool rm = (i < K) && !row_masks[(long long)p * K + i];
        const bool cm = (j < K) && !col_masks[(long long)p * K + j];
        if (rm || cm) z = -inf;
        Z[i * ld + j] = z;
    }
    for (int i = threadIdx.x; i < K1; i += blockDim.x) {
        float mu = (i < K) ? norm : logf(nvc) + norm;
        float nu = (i < K) ? norm : logf(nvr) + norm;
        if (i < K && !row_masks[(long long)p * K + i]) mu = -inf;
        if (i < K && !col_masks[(long long)p * K + i]) nu = -inf;
        lmu[i] = mu; lnu[i] = nu; u[i] = 0.f; v[i] = 0.f;
    }
    __syncthreads();
    // 8 lanes per row/column (4 rows per warp at once): 3-step shuffle reductions and ~K/8 independent exp per lane keep the
    // dependent chain of one half-iteration short (the kernel is latency-, not throughput-bound)
    const int nslot = (blockDim.x >> 5) * 4;
    const int grp = lane >> 3, sub = lane & 7;
    for (int it = 0; it < iters; ++it) {
        for (int base = warp * 4; base < K1; base += nslot) {     // warp-uniform trip count: every lane joins the shuffles
            const bool act = base + grp < K1;
            const int i = act ? base + grp : K1 - 1;
            const float* zr = Z + i * ld;
            float mx = -INFINITY;
            for (int j = sub; j < K1; j += 8) mx = fmaxf(mx, zr[j] + v[j]);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float s = 0.f;
            for (int j = sub; j < K1; j += 8) s += expf(zr[j] + v[j] - mx);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (act && sub == 0) u[i] = lmu[i] - (logf(s) + mx);
        }
        __syncthreads();
        for (int base = warp * 4; base < K1; base += nslot) {
            const bool act = base + grp < K1;
            const int j = act ? base + grp : K1 - 1;
            float mx = -INFINITY;
            for (int i = sub; i < K1; i += 8) mx = fmaxf(mx, Z[i * ld + j] + u[i]);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float s = 0.f;
            for (int i = sub; i < K1; i += 8) s += expf(Z[i * ld + j] + u[i] - mx);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (act && sub == 0) v[j] = lnu[j] - (logf(s) + mx);
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < K1 * K1; e += blockDim.x) {
        const int i = e / K1, j = e % K1;
        out[(long long)p * K1 * K1 + e] = Z[i * ld + j] + u[i] + v[j] - norm;
    }
}

// Register-resident variant (K+1 <= MAXT / LPR rows): every thread keeps its NE entries of its row AND its NE entries of its
// column in registers for all iterations, in the log2 domain (Z * log2 e), so that a half-iteration is one shared-memory load
// (the other potential), one ex2 and ~4 ALU instructions per entry.  The generic kernel above re-reads Z from shared memory
// twice per entry and pays expf's range reduction; it is issue-bound at 64 resident warps per SM.
template <int LPR, int NE, int MAXT>
__global__ void __launch_bounds__(MAXT) sinkhorn_reg_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ row_masks,
                                                           const unsigned char* __restrict__ col_masks, const float* __restrict__ alpha_p,
                                                           int K, int iters, float inf, float* __restrict__ out) {
    extern __shared__ float sm[];
    const int K1 = K + 1;
    const int ld = K1 | 1;
    float* Z = sm;                           // [K1][ld], natural-log domain (also the source of the final output)
    float* u = Z + K1 * ld;                  // log2 domain
    float* v = u + K1;
    __shared__ float norm_s;
    __shared__ int cnt_s[2];
    const int p = blockIdx.x;
    const float alpha = *alpha_p;
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    if (threadIdx.x < 2) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    {
        int cr = 0, cc = 0;
        for (int i = threadIdx.x; i < K; i += blockDim.x) { cr += row_masks[(long long)p * K + i] ? 1 : 0; cc += col_masks[(long long)p * K + i] ? 1 : 0; }
        if (cr) atomicAdd(&cnt_s[0], cr);
        if (cc) atomicAdd(&cnt_s[1], cc);
    }
    __syncthreads();
    const float nvr = (float)cnt_s[0], nvc = (float)cnt_s[1];
    if (threadIdx.x == 0) norm_s = -logf(nvr + nvc);
    __syncthreads();
    const float norm = norm_s;
    for (int e = threadIdx.x; e < K1 * K1; e += blockDim.x) {
        const int i = e / K1, j = e % K1;
        float z = (i < K && j < K) ? scores[((long long)p * K + i) * K + j] : alpha;
        const bool rm = (i < K) && !row_masks[(long long)p * K + i];
        const bool cm = (j < K) && !col_masks[(long long)p * K + j];
        if (rm || cm) z = -inf;
        Z[i * ld + j] = z;
    }
    for (int i = threadIdx.x; i < K1; i += blockDim.x) { u[i] = 0.f; v[i] = 0.f; }
    __syncthreads();
    const int slot = threadIdx.x / LPR, sub = threadIdx.x % LPR;
    const bool act = slot < K1;
    const int ij = act ? slot : K1 - 1;
    float zr[NE], zc[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int o = sub + LPR * k;
        zr[k] = (o < K1) ? Z[ij * ld + o] * LOG2E : -INFINITY;      // row ij, column o
        zc[k] = (o < K1) ? Z[o * ld + ij] * LOG2E : -INFINITY;      // column ij, row o
    }
    float lmu2, lnu2;                                               // log2-domain marginals of row / column ij
    {
        float mu = (ij < K) ? norm : logf(nvc) + norm;
        float nu = (ij < K) ? norm : logf(nvr) + norm;
        if (ij < K && !row_masks[(long long)p * K + ij]) mu = -inf;
        if (ij < K && !col_masks[(long long)p * K + ij]) nu = -inf;
        lmu2 = mu * LOG2E;
        lnu2 = nu * LOG2E;
    }
    for (int it = 0; it < iters; ++it) {
        {
            float t[NE], mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                const int o = sub + LPR * k;
                t[k] = zr[k] + v[o < K1 ? o : 0];
                mx = fmaxf(mx, t[k]);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NE; ++k) s += exp2f(t[k] - mx);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (act && sub == 0) u[ij] = lmu2 - (log2f(s) + mx);
        }
        __syncthreads();
        {
            float t[NE], mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                const int o = sub + LPR * k;
                t[k] = zc[k] + u[o < K1 ? o : 0];
                mx = fmaxf(mx, t[k]);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NE; ++k) s += exp2f(t[k] - mx);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (act && sub == 0) v[ij] = lnu2 - (log2f(s) + mx);
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < K1 * K1; e += blockDim.x) {
        const int i = e / K1, j = e % K1;
        out[(long long)p * K1 * K1 + e] = Z[i * ld + j] + (u[i] + v[j]) * LN2 - norm;
    }
}

}  // namespace geob200

using namespace geob200;

extern "C" {

size_t geob200_superpoint_matching_workspace_bytes(int64_t n_ref, int64_t n_src) {
    size_t nr = (size_t)n_ref, ns = (size_t)n_src;
    return align_up(4 * nr, 256) * 2 + align_up(4 * ns, 256) * 2 + align_up(4 * nr * ns, 256) * 2 + 1024 + 4096;
}

int geob200_superpoint_matching(const float* ref_feats, const float* src_feats, int64_t n_ref, int64_t n_src, int64_t channels,
                                const uint8_t* ref_masks, const uint8_t* src_masks, int64_t num_correspondences, int dual,
                                int64_t* ref_corr_indices, int64_t* src_corr_indices, float* corr_scores, int32_t* num_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GEOB_REQUIRE(n_ref > 0 && n_src > 0, "superpoint_matching: empty input");
    GEOB_REQUIRE(num_correspondences <= 1024, "superpoint_matching: num_correspondences > 1024 unsupported");
    GEOB_REQUIRE(workspace_bytes >= geob200_superpoint_matching_workspace_bytes(n_ref, n_src), "superpoint_matching: workspace too small");
    GEOB_REQUIRE(n_ref * n_src < (1ll << 31), "superpoint_matching: too many node pairs");
    Arena ar(workspace, workspace_bytes);
    int* ridx = ar.take<int>(n_ref);
    int* sidx = ar.take<int>(n_src);
    float* rowsum = ar.take<float>(n_ref);
    float* colsum = ar.take<float>(n_src);
    float* S = ar.take<float>((size_t)n_ref * n_src);
    float* flat = ar.take<float>((size_t)n_ref * n_src);
    int* counts = ar.take<int>(64);
    compact_masks_kernel<<<1, 1024, 0, st>>>(ref_masks, (int)n_ref, ridx, counts);
    compact_masks_kernel<<<1, 1024, 0, st>>>(src_masks, (int)n_src, sidx, counts + 1);
    const size_t smem = sizeof(float) * (channels + n_src);
    GEOB_REQUIRE(smem <= 48 * 1024, "superpoint_matching: row does not fit shared memory");
    spm_scores_kernel<<<(unsigned)n_ref, 256, smem, st>>>(ref_feats, src_feats, (int)channels, ridx, counts, sidx, counts + 1,
                                                         (int)n_src, S, rowsum);
    spm_colsum_kernel<<<(unsigned)((n_src + 255) / 256), 256, 0, st>>>(S, (int)n_src, counts, counts + 1, colsum);
    spm_dual_kernel<<<(unsigned)((n_ref * n_src + 255) / 256), 256, 0, st>>>(S, (int)n_src, counts, counts + 1, rowsum, colsum, dual, flat);
    topk_flat_kernel<1024><<<1, 1024, 0, st>>>(flat, counts, counts + 1, (int)num_correspondences, ridx, sidx,
                                               (long long*)ref_corr_indices, (long long*)src_corr_indices, corr_scores, num_out);
    GEOB_CHECK_LAUNCH();
    count_launches(6);
    return 0;
}

int geob200_gather_patches(const int64_t* corr_indices, int64_t n_corr, const int64_t* node_knn_indices,
                           const uint8_t* node_knn_masks, int64_t k, const float* points, int64_t n_points,
                           int64_t* out_indices, uint8_t* out_masks, float* out_points, void* stream) {
    if (n_corr == 0) return 0;
    gather_patches_kernel<<<(unsigned)((n_corr * k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)corr_indices, (int)n_corr, (const long long*)node_knn_indices, node_knn_masks, (int)k, points,
        (int)n_points, (long long*)out_indices, out_masks, out_points);
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

int geob200_patch_scores(const float* ref_feats, int64_t n_ref, const float* src_feats, int64_t n_src, int64_t channels,
                         const int64_t* ref_knn_indices, const int64_t* src_knn_indices, int64_t n_patches, int64_t k,
                         float* scores, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_patches == 0) return 0;
    const float div = sqrtf((float)channels);      // feats_f.shape[1] ** 0.5
    if (k == 64)
        patch_scores_kernel<4><<<(unsigned)n_patches, 256, 0, st>>>(ref_feats, (int)n_ref, src_feats, (int)n_src, (int)channels,
                                                                    (const long long*)ref_knn_indices, (const long long*)src_knn_indices, div, scores);
    else if (k == 128)
        patch_scores_kernel<8><<<(unsigned)n_patches, 256, 0, st>>>(ref_feats, (int)n_ref, src_feats, (int)n_src, (int)channels,
                                                                    (const long long*)ref_knn_indices, (const long long*)src_knn_indices, div, scores);
    else if (k == 32)
        patch_scores_kernel<2><<<(unsigned)n_patches, 256, 0, st>>>(ref_feats, (int)n_ref, src_feats, (int)n_src, (int)channels,
                                                                    (const long long*)ref_knn_indices, (const long long*)src_knn_indices, div, scores);
    else
        GEOB_REQUIRE(false, "patch_scores: num_points_in_patch=%lld unsupported (32, 64, 128)", (long long)k);
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

int geob200_sinkhorn(const float* scores, const uint8_t* row_masks, const uint8_t* col_masks, const float* alpha,
                     int64_t n_patches, int64_t k, int64_t num_iterations, float inf, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_patches == 0) return 0;
    const int K1 = (int)k + 1, ld = K1 | 1;
    const size_t smem = sizeof(float) * ((size_t)K1 * ld + 4 * K1);
    GEOB_REQUIRE(smem <= 200 * 1024, "sinkhorn: patch too large (k=%lld)", (long long)k);
    if (smem > 48 * 1024 && ensure_max_smem((const void*)sinkhorn_kernel)) return -1;
#define LAUNCH_SK_REG(LPRV, NEV, MT)                                                                                                  \
    {                                                                                                                            \
        if (smem > 48 * 1024 && ensure_max_smem((const void*)sinkhorn_reg_kernel<LPRV, NEV, MT>)) return -1;                         \
        const int threads = ((K1 * LPRV + 31) / 32) * 32;                                                                        \
        sinkhorn_reg_kernel<LPRV, NEV, MT><<<(unsigned)n_patches, threads, smem, st>>>(scores, row_masks, col_masks, alpha, (int)k,  \
                                                                                  (int)num_iterations, inf, out);               \
    }
    if (K1 <= 40) LAUNCH_SK_REG(8, 5, 320)
    else if (K1 <= 72) LAUNCH_SK_REG(8, 9, 576)
    else if (K1 <= 132) LAUNCH_SK_REG(4, 33, 544)
    else
        sinkhorn_kernel<<<(unsigned)n_patches, 1024, smem, st>>>(scores, row_masks, col_masks, alpha, (int)k, (int)num_iterations, inf, out);
#undef LAUNCH_SK_REG
    GEOB_CHECK_LAUNCH();
    count_launches(1);
    return 0;
}

}  // extern "C"
