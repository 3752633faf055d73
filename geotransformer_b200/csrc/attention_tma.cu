// Self-attention with the geometric structure term, TMA-staged (C = 128 or 256, batched over clouds).
//
// Reference: geotransformer/modules/transformer/rpe_transformer.py:51-70 -- scores = (q.k^T + q.proj_p(E)^T) / sqrt(d),
// softmax over the keys, hidden = P.v.  With proj_p moved onto q (attention.cu) the structure term is qp[n,h,:] . E[n,m,:]:
// the only large operand is E (N x M x C fp32, 105 MB per cloud at N = 320, C = 256), read exactly once per layer.
//
// Three launches per layer for ALL clouds of a batch:
//   att_qk_kernel      S[n,h,m] = q_n,h . k_m,h                      (tiled FFMA, 26 MFLOP per cloud: k is read once per tile,
//                                                                    not once per query as in the lanes<->channels kernel)
//   att_stream_kernel  streams E[n] (a contiguous M*C*4-byte block) with cp.async.bulk (TMA, UBLKCP) into a ring of 16-key
//                      stages, one elected producer lane, mbarrier full/empty pairs; 8 consumer warps (lanes <-> channels,
//                      qp slices in registers) add the structure term, then softmax over the keys in shared memory;
//                      S is overwritten by the probabilities P.  HBM-bound on E.
//   att_pv_kernel      out[n, h*d:(h+1)*d] = P[n,h,:] . v[:, h*d:(h+1)*d]   (tiled FFMA, v read once per 32-query tile)
#include "attention.cuh"
#include "common.cuh"
#include "geob200.h"

namespace geob200 {

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(at_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void at_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(at_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(at_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = at_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void at_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(at_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(at_smem_u32(bar))
                 : "memory");
}

// ---- S = q k^T per head ----------------------------------------------------------------------------------------------------
// grid (ceil(maxM/64), ceil(maxN/64), items * H), 256 threads = (16, 16); thread (tx, ty) owns the 4 x 4 block
// S[n0 + 4 ty .. +3][m0 + 4 tx .. +3].  Operands are staged transposed ([d][row]) so that the inner product reads one float4 of
// each per depth step: 2 LDS.128 per 16 FMAs.
constexpr int ATQ_T = 64;                        // tile edge (queries and keys)
constexpr int ATQ_LD = ATQ_T + 4;                // padded row (floats), keeps float4 alignment

template <int H>
__global__ void __launch_bounds__(256) att_qk_kernel(const __grid_constant__ AttBatch b, int ldq, int ldk, int D) {
    __shared__ __align__(16) float Qs[32][ATQ_LD], Ks[32][ATQ_LD];      // depth chunk of 32
    const int item = blockIdx.z / H, h = blockIdx.z % H;
    const AttItem& it = b.it[item];
    const int N = it.N, M = it.M;
    const int n0 = blockIdx.y * ATQ_T, m0 = blockIdx.x * ATQ_T;
    if (n0 >= N || m0 >= M) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d0 = 0; d0 < D; d0 += 32) {
        // 64 rows x 32 depth per operand: thread t loads depth d = t % 32 of rows t / 32 + 8 i (coalesced over d)
        const int dd = threadIdx.x & 31, r8 = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r8 + 8 * i;
            Qs[dd][r] = (n0 + r < N && d0 + dd < D) ? it.q[(long long)(n0 + r) * ldq + h * D + d0 + dd] : 0.f;
            Ks[dd][r] = (m0 + r < M && d0 + dd < D) ? it.k[(long long)(m0 + r) * ldk + h * D + d0 + dd] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
            const float4 qv = *reinterpret_cast<const float4*>(&Qs[d][4 * ty]);
            const float4 kv = *reinterpret_cast<const float4*>(&Ks[d][4 * tx]);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qa[i], ka[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + 4 * ty + i;
        if (n >= N) continue;
        float* row = it.S + ((long long)n * H + h) * M + m0 + 4 * tx;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (m0 + 4 * tx + j < M) row[j] = acc[i][j];
    }
}

// ---- E stream + softmax ------------------------------------------------------------------------------------------------------
constexpr int ATM_KCH = 16;                      // keys per stage
constexpr int ATM_NST = 4;                       // stages in flight per CTA (64 KB at C = 256: three CTAs per SM)
constexpr int ATM_CW = 8;                        // consumer warps (2 keys of a stage each)
constexpr int ATM_THREADS = (ATM_CW + 1) * 32;   // + one producer warp

template <int H, int J>
__global__ void __launch_bounds__(ATM_THREADS) att_stream_kernel(const __grid_constant__ AttBatch b, float div, int max_m) {
    constexpr int C = 128 * J;
    constexpr int NV = 2 * H;                     // (key u < 2, head h) values reduced together
    static_assert(NV == 2 || NV == 4 || NV == 8 || NV == 16, "heads must be 1, 2, 4 or 8");
    extern __shared__ unsigned char atm_smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)atm_smem_raw + 127) & ~(uintptr_t)127);
    float4* ring = reinterpret_cast<float4*>(smem);                                  // [NST][KCH][C/4]
    float* sc = reinterpret_cast<float*>(smem + (size_t)ATM_NST * ATM_KCH * C * 4);  // [H][mp]
    const int mp = (max_m + 3) & ~3;
    uint64_t* full = reinterpret_cast<uint64_t*>(sc + (size_t)H * mp);
    uint64_t* empty = full + ATM_NST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ATM_NST; ++s) { at_mbar_init(&full[s], 1); at_mbar_init(&empty[s], ATM_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long units = b.uprefix[b.n_items];                 // one unit = one query row of one item
    const long long u_begin = units * blockIdx.x / gridDim.x, u_end = units * (blockIdx.x + 1) / gridDim.x;

    if (warp == ATM_CW) {                          // producer
        if (lane == 0) {
            int s = 0, ci = 0;
            uint32_t ph = 0;
            for (long long u = u_begin; u < u_end; ++u) {
                while (u >= b.uprefix[ci + 1]) ++ci;
                const int M = b.it[ci].M;
                const int n = (int)(u - b.uprefix[ci]);
                const float* e_row = b.it[ci].E + (long long)n * M * C;
                for (int m0 = 0; m0 < M; m0 += ATM_KCH) {
                    const uint32_t bytes = (uint32_t)(min(ATM_KCH, M - m0) * C * 4);
                    at_mbar_wait(&empty[s], ph ^ 1u);
                    at_mbar_arrive_expect_tx(&full[s], bytes);
                    at_bulk_g2s(ring + (size_t)s * ATM_KCH * (C / 4), e_row + (long long)m0 * C, bytes, &full[s]);
                    if (++s == ATM_NST) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }

    // consumers: lanes <-> channels (lane owns channels j*128 + 4*lane .. +3 of every 128-channel block j)
    int s = 0, ci = 0;
    uint32_t ph = 0;
    for (long long u = u_begin; u < u_end; ++u) {
        while (u >= b.uprefix[ci + 1]) ++ci;
        const AttItem& it = b.it[ci];
        const int M = it.M;
        const int n = (int)(u - b.uprefix[ci]);
        float4 qpv[H][J];
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int j = 0; j < J; ++j)
                qpv[h][j] = *reinterpret_cast<const float4*>(it.qp + ((long long)n * H + h) * C + j * 128 + 4 * lane);
        // value index handled by this lane after the transposing butterfly: idx = lane >> (5 - log2 NV)
        constexpr int SH = (NV == 16) ? 1 : (NV == 8) ? 2 : (NV == 4) ? 3 : 4;
        const int vidx = lane >> SH, vu = vidx / H, vh = vidx % H;
        const float qb_h = it.qb[(long long)n * H + vh];
        float* S_row = it.S + ((long long)n * H + vh) * M;
        for (int m0 = 0; m0 < M; m0 += ATM_KCH) {
            const int mk = m0 + 2 * warp + vu;                  // the key this lane writes the score of
            const bool writer = (lane & ((1 << SH) - 1)) == 0 && mk < M && (2 * warp + vu) < ATM_KCH;
            const float sqk = writer ? S_row[mk] : 0.f;         // q.k term (att_qk_kernel), fetched while the stage lands
            at_mbar_wait(&full[s], ph);
            const float4* st = ring + (size_t)s * ATM_KCH * (C / 4);
            float v[NV];
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
                const int kk = 2 * warp + uu;                   // key within the stage (rows past M hold stale data: never written out)
#pragma unroll
                for (int h = 0; h < H; ++h) v[uu * H + h] = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 e = st[(size_t)kk * (C / 4) + j * 32 + lane];
#pragma unroll
                    for (int h = 0; h < H; ++h)
                        v[uu * H + h] = fmaf(e.x, qpv[h][j].x, fmaf(e.y, qpv[h][j].y, fmaf(e.z, qpv[h][j].z, fmaf(e.w, qpv[h][j].w, v[uu * H + h]))));
                }
            }
            __syncwarp();
            if (lane == 0) at_mbar_arrive(&empty[s]);           // this warp has read its two keys of the stage
            if (++s == ATM_NST) { s = 0; ph ^= 1u; }
            warp_butterfly(v, lane);
            if (writer) sc[vh * mp + mk] = (sqk + v[0] + qb_h) / div;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ATM_CW * 32) : "memory");       // all scores of this query are in shared memory
        // softmax over the keys: warp h < H normalises head h and writes the probabilities over S
        for (int h = warp; h < H; h += ATM_CW) {
            float* sh = sc + h * mp;
            float mx = -INFINITY;
            for (int m = lane; m < M; m += 32) mx = fmaxf(mx, sh[m]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int m = lane; m < M; m += 32) {
                const float e = expf(sh[m] - mx);
                sh[m] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            float* dst = it.S + ((long long)n * H + h) * M;
            for (int m = lane; m < M; m += 32) dst[m] = sh[m] / sum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ATM_CW * 32) : "memory");       // sc is free for the next query
    }
}

// ---- out = P v per head ------------------------------------------------------------------------------------------------------
// grid (ceil(maxN/64), H * ceil(D/64), items); 256 threads = (16, 16): thread (tx, ty) owns out[n0 + 4 ty .. +3][dbase + 4 tx .. +3].
// P is staged transposed ([key][query]) and v as is ([key][channel]): 2 LDS.128 per 16 FMAs, v read once per 64-query tile.
template <int H>
__global__ void __launch_bounds__(256) att_pv_kernel(const __grid_constant__ AttBatch b, int ldv, int ldo, int D) {
    __shared__ __align__(16) float Ps[32][ATQ_LD], Vs[32][ATQ_LD];      // key chunk of 32
    const AttItem& it = b.it[blockIdx.z];
    const int dtiles = (D + 63) / 64;
    const int h = blockIdx.y / dtiles, d0 = (blockIdx.y % dtiles) * 64;
    const int N = it.N, M = it.M;
    const int n0 = blockIdx.x * ATQ_T;
    if (n0 >= N) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int m0 = 0; m0 < M; m0 += 32) {
        {   // P tile: 64 queries x 32 keys (thread: key = t % 32, queries t / 32 + 8 i; coalesced over the keys)
            const int km = threadIdx.x & 31, r8 = threadIdx.x >> 5;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r8 + 8 * i;
                Ps[km][r] = (n0 + r < N && m0 + km < M) ? it.S[((long long)(n0 + r) * H + h) * M + m0 + km] : 0.f;
            }
            // V tile: 32 keys x 64 channels (thread: channel = t % 64, keys t / 64 + 4 i)
            const int dc = threadIdx.x & 63, r4 = threadIdx.x >> 6;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r4 + 4 * i;
                Vs[r][dc] = (m0 + r < M && d0 + dc < D) ? it.v[(long long)(m0 + r) * ldv + h * D + d0 + dc] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int mm = 0; mm < 32; ++mm) {
            const float4 pv = *reinterpret_cast<const float4*>(&Ps[mm][4 * ty]);
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[mm][4 * tx]);
            const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pa[i], va[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + 4 * ty + i;
        if (n >= N) continue;
        float* row = it.out + (long long)n * ldo + h * D + d0 + 4 * tx;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (d0 + 4 * tx + j < D) row[j] = acc[i][j];
    }
}

// ---- plain softmax over the keys (cross-attention: no structure term) --------------------------------------------------------
// one warp per (query, head) row of S: p = softmax(s / div), in place
template <int H>
__global__ void __launch_bounds__(256) att_softmax_kernel(const __grid_constant__ AttBatch b, float div) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);          // over all (item, n, h)
    if (row >= b.uprefix[b.n_items] * H) return;
    const long long u = row / H;
    int ci = 0;
    while (u >= b.uprefix[ci + 1]) ++ci;
    const AttItem& it = b.it[ci];
    const int M = it.M;
    float* s = it.S + ((u - b.uprefix[ci]) * H + (row % H)) * M;
    float mx = -INFINITY;
    for (int m = lane; m < M; m += 32) mx = fmaxf(mx, s[m] / div);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int m = lane; m < M; m += 32) {
        const float e = expf(s[m] / div - mx);
        s[m] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    for (int m = lane; m < M; m += 32) s[m] = s[m] / sum;
}

template <int H>
static int launch_cross(const AttBatch& b, int ldq, int ldk, int ldv, int ldo, int D, float div, cudaStream_t st) {
    int max_n = 0, max_m = 0;
    for (int i = 0; i < b.n_items; ++i) { max_n = b.it[i].N > max_n ? b.it[i].N : max_n; max_m = b.it[i].M > max_m ? b.it[i].M : max_m; }
    const dim3 qk_grid((unsigned)((max_m + ATQ_T - 1) / ATQ_T), (unsigned)((max_n + ATQ_T - 1) / ATQ_T), (unsigned)(b.n_items * H));
    att_qk_kernel<H><<<qk_grid, 256, 0, st>>>(b, ldq, ldk, D);
    const long long rows = b.uprefix[b.n_items] * H;
    att_softmax_kernel<H><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(b, div);
    const dim3 pv_grid((unsigned)((max_n + ATQ_T - 1) / ATQ_T), (unsigned)(H * ((D + 63) / 64)), (unsigned)b.n_items);
    att_pv_kernel<H><<<pv_grid, 256, 0, st>>>(b, ldv, ldo, D);
    return 0;
}

template <int H, int J>
static int launch_tma(const AttBatch& b, int ldq, int ldk, int ldv, int ldo, float div, cudaStream_t st) {
    constexpr int C = 128 * J, D = C / H;
    int max_n = 0, max_m = 0;
    for (int i = 0; i < b.n_items; ++i) { max_n = b.it[i].N > max_n ? b.it[i].N : max_n; max_m = b.it[i].M > max_m ? b.it[i].M : max_m; }
    const dim3 qk_grid((unsigned)((max_m + ATQ_T - 1) / ATQ_T), (unsigned)((max_n + ATQ_T - 1) / ATQ_T), (unsigned)(b.n_items * H));
    att_qk_kernel<H><<<qk_grid, 256, 0, st>>>(b, ldq, ldk, D);
    const size_t smem = (size_t)ATM_NST * ATM_KCH * C * 4 + sizeof(float) * H * (size_t)((max_m + 3) & ~3) + 2 * ATM_NST * 8 + 256;
    if (ensure_max_smem((const void*)att_stream_kernel<H, J>)) return -1;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, att_stream_kernel<H, J>, ATM_THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
    const long long units = b.uprefix[b.n_items];
    long long grid = (long long)occ * num_sms();
    if (grid > units) grid = units;
    att_stream_kernel<H, J><<<(unsigned)grid, ATM_THREADS, smem, st>>>(b, div, max_m);
    const dim3 pv_grid((unsigned)((max_n + ATQ_T - 1) / ATQ_T), (unsigned)(H * ((D + 63) / 64)), (unsigned)b.n_items);
    att_pv_kernel<H><<<pv_grid, 256, 0, st>>>(b, ldv, ldo, D);
    return 0;
}

// returns 1 when the shape is not handled here (caller falls back to the lanes<->channels kernels of attention.cu)
int attention_tma_batch(const AttBatch& b, int ldq, int ldk, int ldv, int ldo, int channels, int heads, float div, cudaStream_t st) {
    if (b.it[0].E == nullptr) {                                       // cross-attention: q.k^T -> softmax -> P.v as three tiled passes
        int rc = -2;
        switch (heads) {
            case 1: rc = launch_cross<1>(b, ldq, ldk, ldv, ldo, channels / heads, div, st); break;
            case 2: rc = launch_cross<2>(b, ldq, ldk, ldv, ldo, channels / heads, div, st); break;
            case 4: rc = launch_cross<4>(b, ldq, ldk, ldv, ldo, channels / heads, div, st); break;
            case 8: rc = launch_cross<8>(b, ldq, ldk, ldv, ldo, channels / heads, div, st); break;
            default: return 1;
        }
        if (rc != 0) return rc;
        GEOB_CHECK_LAUNCH();
        count_launches(3);
        return 0;
    }
    if (!(channels == 128 || channels == 256)) return 1;
    int max_m = 0;
    for (int i = 0; i < b.n_items; ++i) max_m = b.it[i].M > max_m ? b.it[i].M : max_m;
    const size_t smem = (size_t)ATM_NST * ATM_KCH * channels * 4 + sizeof(float) * heads * (size_t)((max_m + 3) & ~3) + 2 * ATM_NST * 8 + 256;
    if (smem > 200 * 1024) return 1;
    int rc = -2;
#define LAUNCH_TMA(HV) rc = (channels == 256) ? launch_tma<HV, 2>(b, ldq, ldk, ldv, ldo, div, st) : launch_tma<HV, 1>(b, ldq, ldk, ldv, ldo, div, st)
    switch (heads) {
        case 1: LAUNCH_TMA(1); break;
        case 2: LAUNCH_TMA(2); break;
        case 4: LAUNCH_TMA(4); break;
        case 8: LAUNCH_TMA(8); break;
        default: return 1;
    }
#undef LAUNCH_TMA
    if (rc != 0) return rc;
    GEOB_CHECK_LAUNCH();
    count_launches(3);
    return 0;
}

}  // namespace geob200
