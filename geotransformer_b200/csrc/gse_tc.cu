// tcgen05 (5th-gen tensor core) contraction of the geometric structure embedding for C = 256.
//
// Reference semantics: geotransformer/modules/geotransformer/geotransformer.py:57-72
//     E[p,:] = (Wd s(d_p) + bd) + max_k (Wa s(a_pk) + ba)          s(.) = 256-wide interleaved sin/cos embedding
// Because the distance term does not depend on k it can be folded under the max exactly:
//     E[p,:] = max_k ( [Wa | Wd] . [s(a_pk) ; s(d_p)] ) + (ba + bd)
// so EVERY row (p,k) of the problem is one homogeneous GEMM row with K = 512 against ONE operand B = [Wa | Wd]
// (256 x 512).  A tile is 42 pairs = 126 rows (M = 128, two idle rows), N = 256, K = 512.
//
// Pipeline inside one persistent CTA (one per SM), 704 threads:
//   warps 0-15 generators : compute the sinusoid chunk (128 rows x 32 K) with sincosf and store it straight into
//                           the canonical K-major SWIZZLE_128B shared-memory layout the UMMA descriptor expects
//                           (the A operand never exists in HBM)
//   warp 16    copier     : one lane streams the pre-swizzled B chunk (256 x 32 K = 32 KB, a contiguous image packed
//                           by gse_pack_b_kernel) with cp.async.bulk (TMA, UBLKCP) onto the stage's mbarrier
//   warp 17    MMA issuer : one lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 N256 K8) into a TMEM accumulator,
//                           tcgen05.commit releases the stage / publishes the accumulator
//   warps 18-21 epilogue   : tcgen05.ld the accumulator (32 lanes x 32 columns per warp), exchange through shared memory,
//                           max over the three rows of a pair, add the bias, store E -- while the MMA warp already works
//                           on the other accumulator (2 x 256 TMEM columns)
// NPASS = 3 runs the error-compensated 3xTF32 scheme (a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulate in TMEM), which
// is fp32-accurate; NPASS = 1 is plain TF32.
#include <cuda_fp16.h>

#include "common.cuh"
#include "geob200.h"

namespace geob200 {
namespace tc {

constexpr int C = 256;                 // output channels = MMA N
constexpr int KTOT = 512;              // [angle | distance] sinusoids
constexpr int KC = 32;                 // K elements per chunk = 128 bytes = one swizzle atom row
constexpr int NCHUNK = KTOT / KC;      // 16
constexpr int PAIRS = 42;              // pairs per tile
constexpr int ROWS = 128;              // MMA M
constexpr int A_BYTES = ROWS * 128;    // 16 KB
constexpr int B_BYTES = C * 128;       // 32 KB
constexpr int STAGE_LD = 33;           // epilogue exchange row stride (floats)
constexpr int NGEN_WARPS = 16;          // sinusoid generator warps (4 per SM sub-partition: latency hiding)
constexpr int NTHREADS = (NGEN_WARPS + 2 + 4) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!done);
}
// sin and cos of x for moderate |x| (embedding arguments are < ~100 rad): 3-term Cody-Waite reduction by pi/2 and
// degree-7/8 minimax polynomials on [-pi/4, pi/4]; ~1 ulp, about a third of the instructions of sincosf.
__device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
    if (fabsf(x) > 48000.f) { sincosf(x, &s, &c); return; }
    const float k = rintf(x * 0.636619772367581343f);
    float r = fmaf(k, -1.57079601287841796875f, x);
    r = fmaf(k, -3.1391647326017846353352069854736328125e-7f, r);
    r = fmaf(k, -5.390302529957764765e-15f, r);
    const float r2 = r * r;
    const float sp = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
    const float cp = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                          fmaf(r2, -0.5f, 1.0f));
    const int q = (int)k;
    const float ss = (q & 1) ? cp : sp;
    const float cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// Cheaper variant for the fp16 kernels: same Cody-Waite reduction to [-pi/4, pi/4], then the SFU (sin.approx / cos.approx, abs
// error ~4e-7 on that interval -- below the 2^-12 half-ulp of the fp16 hi/lo split that consumes the values).
__device__ __forceinline__ void sincos_sfu(float x, float& s, float& c) {
    if (fabsf(x) > 48000.f) { sincosf(x, &s, &c); return; }
    const float k = rintf(x * 0.636619772367581343f);
    float r = fmaf(k, -1.57079601287841796875f, x);
    r = fmaf(k, -3.1391647326017846353352069854736328125e-7f, r);
    r = fmaf(k, -5.390302529957764765e-15f, r);
    const float sp = __sinf(r), cp = __cosf(r);
    const int q = (int)k;
    const float ss = (q & 1) ? cp : sp;
    const float cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // waits for remote (cluster-scope) arrivals
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
// start address >> 4 [0,14) | LBO >> 4 [16,30) (=1, unused for swizzled K-major) | SBO >> 4 [32,46) (8 rows x 128 B = 1024)
// | version [46,48) = 1 | layout_type [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10), both K-major,
// N >> 3 @ bit 17, M >> 4 @ bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Packs B = [Wa | Wd] (row n = output channel, 512 K values) into per-chunk shared-memory images:
// image[kc][n/8][n%8][(e/4) ^ (n%8)][e%4], hi part (tf32, round-to-nearest) and lo part (remainder).
__global__ void __launch_bounds__(256) gse_pack_b_kernel(const float* __restrict__ Wd, const float* __restrict__ Wa,
                                                         const float* __restrict__ bd, const float* __restrict__ ba,
                                                         float* __restrict__ img_hi, float* __restrict__ img_lo,
                                                         float* __restrict__ bias_sum) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < C) bias_sum[t] = ba[t] + bd[t];
    if (t >= C * KTOT) return;
    const int n = t / KTOT, k = t % KTOT;
    const float w = (k < 256) ? Wa[n * 256 + k] : Wd[n * 256 + (k - 256)];
    const float hi = tf32_rn(w);
    const int kc = k / KC, e = k % KC;
    const int dst = kc * (C * KC) + (n >> 3) * 256 + (n & 7) * 32 + (((e >> 2) ^ (n & 7)) << 2) + (e & 3);
    img_hi[dst] = hi;
    img_lo[dst] = w - hi;
}

template <int NPASS>
struct Cfg {
    static constexpr int STAGE_BYTES = (NPASS == 3) ? 2 * (A_BYTES + B_BYTES) : (A_BYTES + B_BYTES);
    static constexpr int NSTAGE = (NPASS == 3) ? 2 : 4;
    static constexpr int SMEM = NSTAGE * STAGE_BYTES + ROWS * STAGE_LD * 4 + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int NPASS>
__global__ void __launch_bounds__(NTHREADS, 1) gse_embed_tc_kernel(const float* __restrict__ d_idx, const float* __restrict__ a_idx,
                                                                   long long n_pairs, const float* __restrict__ div_term,
                                                                   const float* __restrict__ img_hi, const float* __restrict__ img_lo,
                                                                   const float* __restrict__ bias_sum, float* __restrict__ E) {
    using CF = Cfg<NPASS>;
    constexpr int NSTAGE = CF::NSTAGE;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B: 1024 B alignment
    float* xstage = (float*)(smem + NSTAGE * CF::STAGE_BYTES);                                   // [128][33]
    uint64_t* bars = (uint64_t*)(xstage + ROWS * STAGE_LD);
    uint64_t* full = bars;                 // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
    uint64_t* tfull = bars + 2 * NSTAGE;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (n_pairs + PAIRS - 1) / PAIRS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], NGEN_WARPS + 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NGEN_WARPS + 1) {   // TMEM: 512 columns = two 128 x 256 fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NGEN_WARPS) {
        // ===================== generators =====================
        // work item = (tile row, 16-byte chunk c) -> frequencies f0+2c, f0+2c+1 (sin, cos interleaved).  Angle chunks
        // (kc < 8) have 126 x 8 items; distance chunks (kc >= 8) have 42 x 8 items whose result is stored to the three
        // rows (k = 0,1,2) of the pair.
        {   // padding rows 126,127 of every stage: zero once (their outputs are never read)
            for (int e = threadIdx.x; e < NSTAGE * ((NPASS == 3) ? 2 : 1) * 64; e += NGEN_WARPS * 32) {
                const int sidx = e / (((NPASS == 3) ? 2 : 1) * 64), rem = e % (((NPASS == 3) ? 2 : 1) * 64);
                const int part = rem / 64, w = rem % 64;
                float* base = (float*)(smem + sidx * CF::STAGE_BYTES + part * A_BYTES + 15 * 1024 + 6 * 128);
                base[w] = 0.f;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(NGEN_WARPS * 32) : "memory");
        }
        int s = 0;
        uint32_t ph = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long p0 = tile * PAIRS;
            for (int kc = 0; kc < NCHUNK; ++kc) {
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* st = smem + s * CF::STAGE_BYTES;
                const int f0 = (kc & 7) * 16;
                const bool angle = kc < 8;
                const int n_items = (angle ? 3 * PAIRS : PAIRS) * 8;
                for (int it = threadIdx.x; it < n_items; it += NGEN_WARPS * 32) {
                    const int c = it & 7, rr = it >> 3;          // rr: tile row (angle) or pair slot (distance)
                    const int j = angle ? (rr % PAIRS) : rr;
                    const long long p = p0 + j;
                    float x = 0.f;
                    if (p < n_pairs) x = angle ? __ldg(a_idx + p * 3 + rr / PAIRS) : __ldg(d_idx + p);
                    float s0, c0, s1, c1;
                    sincos_cw(__fmul_rn(x, __ldg(div_term + f0 + 2 * c)), s0, c0);
                    sincos_cw(__fmul_rn(x, __ldg(div_term + f0 + 2 * c + 1)), s1, c1);
                    float4 hi = make_float4(s0, c0, s1, c1), lo = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (NPASS == 3) {
                        hi.x = tf32_rn(s0); lo.x = s0 - hi.x;
                        hi.y = tf32_rn(c0); lo.y = c0 - hi.y;
                        hi.z = tf32_rn(s1); lo.z = s1 - hi.z;
                        hi.w = tf32_rn(c1); lo.w = c1 - hi.w;
                    }
                    const int nrep = angle ? 1 : 3;
                    for (int rep = 0; rep < nrep; ++rep) {
                        const int r = angle ? rr : (rep * PAIRS + j);
                        const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                        *reinterpret_cast<float4*>(st + off) = hi;
                        if (NPASS == 3) *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
                    }
                }
                fence_proxy_async();               // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                if (++s == NSTAGE) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == NGEN_WARPS) {
        // ===================== B copier (TMA bulk) =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < NCHUNK; ++kc) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* st = smem + s * CF::STAGE_BYTES;
                    const int a_bytes = (NPASS == 3) ? 2 * A_BYTES : A_BYTES;
                    mbar_arrive_expect_tx(&full[s], (NPASS == 3) ? 2 * B_BYTES : B_BYTES);
                    bulk_g2s(st + a_bytes, img_hi + (size_t)kc * (C * KC), B_BYTES, &full[s]);
                    if (NPASS == 3) bulk_g2s(st + a_bytes + B_BYTES, img_lo + (size_t)kc * (C * KC), B_BYTES, &full[s]);
                    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == NGEN_WARPS + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph[2] = {0, 0};
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_ph[acc] ^ 1u);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C);
                for (int kc = 0; kc < NCHUNK; ++kc) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * CF::STAGE_BYTES);
                    const uint32_t a_bytes = (NPASS == 3) ? 2 * A_BYTES : A_BYTES;
                    const uint64_t da_hi = make_desc(st), db_hi = make_desc(st + a_bytes);
#pragma unroll
                    for (int kk = 0; kk < KC / 8; ++kk) {       // 4 MMAs of K = 8 (32 bytes) per chunk
                        const uint64_t adv = (uint64_t)(kk * 2);     // +32 bytes in 16-byte units
                        const uint32_t first = (kc == 0 && kk == 0) ? 0u : 1u;
                        umma_tf32(d_tmem, da_hi + adv, db_hi + adv, first);
                        if (NPASS == 3) {
                            const uint64_t da_lo = make_desc(st + A_BYTES), db_lo = make_desc(st + a_bytes + B_BYTES);
                            umma_tf32(d_tmem, da_hi + adv, db_lo + adv, 1u);
                            umma_tf32(d_tmem, da_lo + adv, db_hi + adv, 1u);
                        }
                    }
                    umma_commit(&empty[s]);                      // stage reusable once these MMAs have read it
                    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
                }
                umma_commit(&tfull[acc]);                        // accumulator complete
                acc_ph[acc] ^= 1u;
                acc ^= 1;
            }
        }
    } else {
        // ===================== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
        const int q = warp & 3;
        const int et = (warp - (NGEN_WARPS + 2)) * 32 + lane;                    // 0..127 within the epilogue group
        int acc = 0;
        uint32_t acc_ph[2] = {0, 0};
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tfull[acc], acc_ph[acc]);
            tc_fence_after();
            const long long p0 = tile * PAIRS;
            for (int cc = 0; cc < C / 32; ++cc) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + cc * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* xr = xstage + (q * 32 + lane) * STAGE_LD;
#pragma unroll
                for (int c = 0; c < 32; ++c) xr[c] = __uint_as_float(v[c]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int idx = et; idx < PAIRS * 32; idx += 128) {
                    const int jj = idx >> 5, c = idx & 31;
                    const long long p = p0 + jj;
                    if (p < n_pairs) {
                        const float m = fmaxf(fmaxf(xstage[jj * STAGE_LD + c], xstage[(PAIRS + jj) * STAGE_LD + c]),
                                              xstage[(2 * PAIRS + jj) * STAGE_LD + c]);
                        E[p * C + cc * 32 + c] = m + __ldg(bias_sum + cc * 32 + c);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc_ph[acc] ^= 1u;
            acc ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NGEN_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ============================================================================================================
// 3xFP16 variant (default): x = x_hi + x_lo with x_hi = fp16(x), x_lo = fp16(x - x_hi).  fp16 carries the same 11 significant
// bits as tf32, so hi.hi + hi.lo + lo.hi has the same ~2^-22 relative accuracy as 3xTF32, but kind::f16 MMAs consume K = 16
// per instruction at the same rate at which kind::tf32 consumes K = 8: half the tensor-pipe time, half the B bytes, half
// the generator stores.  The sinusoids are in [-1, 1] (lo <= 2^-12: fp16 subnormal spacing 6e-8 = fp32 epsilon); the weights
// are pre-scaled by a power of two to [0.5, 1) (exact), undone in the epilogue.
// ============================================================================================================
constexpr int KC16 = 64;                  // halves per 128-byte row
constexpr int NCHUNK16 = KTOT / KC16;     // 8
constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);   // A = B = F16 (0), D = F32

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate, uint32_t idesc = kIdescF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// scale[0] = 2^-e with max|W| = m 2^e, m in [0.5,1)  ; scale[1] = 2^e
template <int CC>
__global__ void __launch_bounds__(1024) gse_absmax_kernel(const float* __restrict__ Wd, const float* __restrict__ Wa, float* __restrict__ scale) {
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < CC * CC; i += blockDim.x) m = fmaxf(m, fmaxf(fabsf(Wd[i]), fabsf(Wa[i])));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) m = fmaxf(m, red[w]);
        int e = 0;
        if (m > 0.f && isfinite(m)) frexpf(m, &e);
        scale[0] = ldexpf(1.0f, -e);
        scale[1] = ldexpf(1.0f, e);
    }
}

// CC = hidden dim (256: 3DMatch / ModelNet, 128: KITTI); B = [Wa | Wd] is (CC rows) x (2 CC halves)
template <int CC>
__global__ void __launch_bounds__(256) gse_pack_b_f16_kernel(const float* __restrict__ Wd, const float* __restrict__ Wa,
                                                             const float* __restrict__ bd, const float* __restrict__ ba,
                                                             const float* __restrict__ scale, __half* __restrict__ img_hi,
                                                             __half* __restrict__ img_lo, float* __restrict__ bias_sum) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < CC) bias_sum[t] = ba[t] + bd[t];
    if (t >= CC * 2 * CC) return;
    const int n = t / (2 * CC), k = t % (2 * CC);
    const float w = ((k < CC) ? Wa[n * CC + k] : Wd[n * CC + (k - CC)]) * scale[0];
    const __half hi = __float2half_rn(w);
    const int kc = k / KC16, e = k % KC16;
    const int dst = kc * (CC * KC16) + (n >> 3) * 512 + (n & 7) * 64 + (((e >> 3) ^ (n & 7)) << 3) + (e & 7);
    img_hi[dst] = hi;
    img_lo[dst] = __float2half_rn(w - __half2float(hi));
}

constexpr int F16_STAGE_BYTES = 2 * (A_BYTES + B_BYTES);      // hi + lo of both operands = 96 KB
constexpr int F16_NSTAGE = 2;
constexpr int F16_SMEM = F16_NSTAGE * F16_STAGE_BYTES + ROWS * STAGE_LD * 4 + 1024 + 256;

template <int CC>
__global__ void __launch_bounds__(NTHREADS, 1) gse_embed_f16_kernel(const float* __restrict__ d_idx, const float* __restrict__ a_idx,
                                                                    long long n_pairs, const float* __restrict__ div_term,
                                                                    const __half* __restrict__ img_hi, const __half* __restrict__ img_lo,
                                                                    const float* __restrict__ bias_sum, const float* __restrict__ scale,
                                                                    float* __restrict__ E) {
    constexpr int NSTAGE = F16_NSTAGE;
    constexpr int CP = CC / KC16;                         // chunks per part (angle, distance)
    constexpr int NCH = 2 * CP;                           // K chunks per tile
    constexpr int BB = CC * 128;                          // bytes of one B chunk (CC rows x 64 halves)
    constexpr int STB = 2 * (A_BYTES + BB);               // stage: hi + lo of both operands
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(CC >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* xstage = (float*)(smem + NSTAGE * STB);
    uint64_t* bars = (uint64_t*)(xstage + ROWS * STAGE_LD);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* tfull = bars + 2 * NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (n_pairs + PAIRS - 1) / PAIRS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], NGEN_WARPS + 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NGEN_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * CC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NGEN_WARPS) {
        {   // zero the two padding rows (126, 127) of every A buffer once
            for (int e = threadIdx.x; e < NSTAGE * 2 * 64; e += NGEN_WARPS * 32) {
                const int sidx = e / 128, rem = e % 128, part = rem / 64, w = rem % 64;
                float* base = (float*)(smem + sidx * STB + part * A_BYTES + 15 * 1024 + 6 * 128);
                base[w] = 0.f;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(NGEN_WARPS * 32) : "memory");
        }
        int s = 0;
        uint32_t ph = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long p0 = tile * PAIRS;
            for (int kc = 0; kc < NCH; ++kc) {
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* st = smem + s * STB;
                const int f0 = (kc % CP) * 32;               // 32 frequencies (64 halves) per chunk
                const bool angle = kc < CP;
                const int n_items = (angle ? 3 * PAIRS : PAIRS) * 8;
                for (int it = threadIdx.x; it < n_items; it += NGEN_WARPS * 32) {
                    const int c = it & 7, rr = it >> 3;     // 16-byte chunk c = frequencies f0+4c .. f0+4c+3
                    const int j = angle ? (rr % PAIRS) : rr;
                    const long long p = p0 + j;
                    float x = 0.f;
                    if (p < n_pairs) x = angle ? __ldg(a_idx + p * 3 + rr / PAIRS) : __ldg(d_idx + p);
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) sincos_sfu(__fmul_rn(x, __ldg(div_term + f0 + 4 * c + u)), v[2 * u], v[2 * u + 1]);
                    __half2 hi[4], lo[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const __half h0 = __float2half_rn(v[2 * u]), h1 = __float2half_rn(v[2 * u + 1]);
                        hi[u] = __halves2half2(h0, h1);
                        lo[u] = __halves2half2(__float2half_rn(v[2 * u] - __half2float(h0)), __float2half_rn(v[2 * u + 1] - __half2float(h1)));
                    }
                    const uint4 hv = make_uint4(*reinterpret_cast<uint32_t*>(&hi[0]), *reinterpret_cast<uint32_t*>(&hi[1]),
                                                *reinterpret_cast<uint32_t*>(&hi[2]), *reinterpret_cast<uint32_t*>(&hi[3]));
                    const uint4 lv = make_uint4(*reinterpret_cast<uint32_t*>(&lo[0]), *reinterpret_cast<uint32_t*>(&lo[1]),
                                                *reinterpret_cast<uint32_t*>(&lo[2]), *reinterpret_cast<uint32_t*>(&lo[3]));
                    const int nrep = angle ? 1 : 3;
                    for (int rep = 0; rep < nrep; ++rep) {
                        const int r = angle ? rr : (rep * PAIRS + j);
                        const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                        *reinterpret_cast<uint4*>(st + off) = hv;
                        *reinterpret_cast<uint4*>(st + A_BYTES + off) = lv;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                if (++s == NSTAGE) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == NGEN_WARPS) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < NCH; ++kc) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* st = smem + s * STB;
                    mbar_arrive_expect_tx(&full[s], 2 * BB);
                    bulk_g2s(st + 2 * A_BYTES, img_hi + (size_t)kc * (CC * KC16), BB, &full[s]);
                    bulk_g2s(st + 2 * A_BYTES + BB, img_lo + (size_t)kc * (CC * KC16), BB, &full[s]);
                    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == NGEN_WARPS + 1) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph[2] = {0, 0};
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_ph[acc] ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CC);
                for (int kc = 0; kc < NCH; ++kc) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * STB);
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + A_BYTES);
                    const uint64_t db_hi = make_desc(st + 2 * A_BYTES), db_lo = make_desc(st + 2 * A_BYTES + BB);
#pragma unroll
                    for (int kk = 0; kk < KC16 / 16; ++kk) {        // 4 MMAs of K = 16 (32 bytes) per operand pair
                        const uint64_t adv = (uint64_t)(kk * 2);
                        umma_f16(d_tmem, da_hi + adv, db_hi + adv, (kc == 0 && kk == 0) ? 0u : 1u, IDESC);
                        umma_f16(d_tmem, da_hi + adv, db_lo + adv, 1u, IDESC);
                        umma_f16(d_tmem, da_lo + adv, db_hi + adv, 1u, IDESC);
                    }
                    umma_commit(&empty[s]);
                    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
                }
                umma_commit(&tfull[acc]);
                acc_ph[acc] ^= 1u;
                acc ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int et = (warp - (NGEN_WARPS + 2)) * 32 + lane;
        const float inv_scale = scale[1];
        int acc = 0;
        uint32_t acc_ph[2] = {0, 0};
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tfull[acc], acc_ph[acc]);
            tc_fence_after();
            const long long p0 = tile * PAIRS;
            for (int cc = 0; cc < CC / 32; ++cc) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CC + cc * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* xr = xstage + (q * 32 + lane) * STAGE_LD;
#pragma unroll
                for (int c = 0; c < 32; ++c) xr[c] = __uint_as_float(v[c]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int idx = et; idx < PAIRS * 32; idx += 128) {
                    const int jj = idx >> 5, c = idx & 31;
                    const long long p = p0 + jj;
                    if (p < n_pairs) {
                        const float m = fmaxf(fmaxf(xstage[jj * STAGE_LD + c], xstage[(PAIRS + jj) * STAGE_LD + c]),
                                              xstage[(2 * PAIRS + jj) * STAGE_LD + c]);
                        E[p * CC + cc * 32 + c] = fmaf(m, inv_scale, __ldg(bias_sum + cc * 32 + c));
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc_ph[acc] ^= 1u;
            acc ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NGEN_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * CC) : "memory");
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) gse_embed_f16_cluster_kernel(const float* __restrict__ d_idx, const float* __restrict__ a_idx,
                                                                    long long n_pairs, const float* __restrict__ div_term,
                                                                    const __half* __restrict__ img_hi, const __half* __restrict__ img_lo,
                                                                    const float* __restrict__ bias_sum, const float* __restrict__ scale,
                                                                    float* __restrict__ E) {
    constexpr int NSTAGE = F16_NSTAGE;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* xstage = (float*)(smem + NSTAGE * F16_STAGE_BYTES);
    uint64_t* bars = (uint64_t*)(xstage + ROWS * STAGE_LD);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* tfull = bars + 2 * NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint64_t* peer_empty = tempty + 2;     // [NSTAGE]: the peer CTA's stage is free (remote arrive)
    uint32_t* tmem_slot = (uint32_t*)(peer_empty + NSTAGE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (n_pairs + PAIRS - 1) / PAIRS;
    // the two CTAs of a cluster advance in lockstep (they share every B chunk): same number of tile slots for everybody
    const long long n_iters = (n_tiles + gridDim.x - 1) / gridDim.x;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], NGEN_WARPS + 1); mbar_init(&empty[s], 1); mbar_init(&peer_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NGEN_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peer barriers are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NGEN_WARPS) {
        {   // zero the two padding rows (126, 127) of every A buffer once
            for (int e = threadIdx.x; e < NSTAGE * 2 * 64; e += NGEN_WARPS * 32) {
                const int sidx = e / 128, rem = e % 128, part = rem / 64, w = rem % 64;
                float* base = (float*)(smem + sidx * F16_STAGE_BYTES + part * A_BYTES + 15 * 1024 + 6 * 128);
                base[w] = 0.f;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(NGEN_WARPS * 32) : "memory");
        }
        int s = 0;
        uint32_t ph = 0;
        for (long long itile = 0; itile < n_iters; ++itile) {
            const long long tile = itile * gridDim.x + blockIdx.x;   // tile >= n_tiles: idle slot, pipeline still runs
            const long long p0 = tile * PAIRS;
            for (int kc = 0; kc < NCHUNK16; ++kc) {
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* st = smem + s * F16_STAGE_BYTES;
                const int f0 = (kc & 3) * 32;               // 32 frequencies (64 halves) per chunk
                const bool angle = kc < 4;
                const int n_items = (angle ? 3 * PAIRS : PAIRS) * 8;
                for (int it = threadIdx.x; it < n_items; it += NGEN_WARPS * 32) {
                    const int c = it & 7, rr = it >> 3;     // 16-byte chunk c = frequencies f0+4c .. f0+4c+3
                    const int j = angle ? (rr % PAIRS) : rr;
                    const long long p = p0 + j;
                    float x = 0.f;
                    if (p < n_pairs) x = angle ? __ldg(a_idx + p * 3 + rr / PAIRS) : __ldg(d_idx + p);
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) sincos_sfu(__fmul_rn(x, __ldg(div_term + f0 + 4 * c + u)), v[2 * u], v[2 * u + 1]);
                    __half2 hi[4], lo[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const __half h0 = __float2half_rn(v[2 * u]), h1 = __float2half_rn(v[2 * u + 1]);
                        hi[u] = __halves2half2(h0, h1);
                        lo[u] = __halves2half2(__float2half_rn(v[2 * u] - __half2float(h0)), __float2half_rn(v[2 * u + 1] - __half2float(h1)));
                    }
                    const uint4 hv = make_uint4(*reinterpret_cast<uint32_t*>(&hi[0]), *reinterpret_cast<uint32_t*>(&hi[1]),
                                                *reinterpret_cast<uint32_t*>(&hi[2]), *reinterpret_cast<uint32_t*>(&hi[3]));
                    const uint4 lv = make_uint4(*reinterpret_cast<uint32_t*>(&lo[0]), *reinterpret_cast<uint32_t*>(&lo[1]),
                                                *reinterpret_cast<uint32_t*>(&lo[2]), *reinterpret_cast<uint32_t*>(&lo[3]));
                    const int nrep = angle ? 1 : 3;
                    for (int rep = 0; rep < nrep; ++rep) {
                        const int r = angle ? rr : (rep * PAIRS + j);
                        const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                        *reinterpret_cast<uint4*>(st + off) = hv;
                        *reinterpret_cast<uint4*>(st + A_BYTES + off) = lv;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                if (++s == NSTAGE) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == NGEN_WARPS) {
        if (lane == 0) {
            // B copier: rank 0 streams the hi image, rank 1 the lo image, each MULTICAST into both CTAs of the cluster
            // (cp.async.bulk ... .multicast::cluster): every B byte leaves L2 once per CTA pair instead of once per CTA.
            const uint32_t peer = cta_rank ^ 1u;
            int s = 0;
            uint32_t ph = 0;
            for (long long itile = 0; itile < n_iters; ++itile) {
                for (int kc = 0; kc < NCHUNK16; ++kc) {
                    mbar_wait(&empty[s], ph ^ 1u);                         // my MMAs are done with stage s
                    {   // tell the peer, then wait until the peer's stage s is free too
                        uint32_t raddr;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(&peer_empty[s])), "r"(peer));
                        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
                    }
                    mbar_wait_cluster(&peer_empty[s], ph);
                    unsigned char* st = smem + s * F16_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], 2 * B_BYTES);          // both halves land here (mine + the peer's multicast)
                    const __half* src = (cta_rank == 0 ? img_hi : img_lo) + (size_t)kc * (C * KC16);
                    unsigned char* dst = st + 2 * A_BYTES + (cta_rank == 0 ? 0 : B_BYTES);
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                            smem_u32(dst)),
                        "l"(src), "r"((uint32_t)B_BYTES), "r"(smem_u32(&full[s])), "h"((unsigned short)3)
                        : "memory");
                    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == NGEN_WARPS + 1) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph[2] = {0, 0};
            for (long long itile = 0; itile < n_iters; ++itile) {
                // slots with itile * gridDim.x + blockIdx.x >= n_tiles are idle, the pipeline still runs
                mbar_wait(&tempty[acc], acc_ph[acc] ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C);
                for (int kc = 0; kc < NCHUNK16; ++kc) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * F16_STAGE_BYTES);
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + A_BYTES);
                    const uint64_t db_hi = make_desc(st + 2 * A_BYTES), db_lo = make_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < KC16 / 16; ++kk) {        // 4 MMAs of K = 16 (32 bytes) per operand pair
                        const uint64_t adv = (uint64_t)(kk * 2);
                        umma_f16(d_tmem, da_hi + adv, db_hi + adv, (kc == 0 && kk == 0) ? 0u : 1u);
                        umma_f16(d_tmem, da_hi + adv, db_lo + adv, 1u);
                        umma_f16(d_tmem, da_lo + adv, db_hi + adv, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
                }
                umma_commit(&tfull[acc]);
                acc_ph[acc] ^= 1u;
                acc ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int et = (warp - (NGEN_WARPS + 2)) * 32 + lane;
        const float inv_scale = scale[1];
        int acc = 0;
        uint32_t acc_ph[2] = {0, 0};
        for (long long itile = 0; itile < n_iters; ++itile) {
            const long long tile = itile * gridDim.x + blockIdx.x;   // tile >= n_tiles: idle slot, pipeline still runs
            mbar_wait(&tfull[acc], acc_ph[acc]);
            tc_fence_after();
            const long long p0 = tile * PAIRS;
            for (int cc = 0; cc < C / 32; ++cc) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + cc * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* xr = xstage + (q * 32 + lane) * STAGE_LD;
#pragma unroll
                for (int c = 0; c < 32; ++c) xr[c] = __uint_as_float(v[c]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int idx = et; idx < PAIRS * 32; idx += 128) {
                    const int jj = idx >> 5, c = idx & 31;
                    const long long p = p0 + jj;
                    if (p < n_pairs) {
                        const float m = fmaxf(fmaxf(xstage[jj * STAGE_LD + c], xstage[(PAIRS + jj) * STAGE_LD + c]),
                                              xstage[(2 * PAIRS + jj) * STAGE_LD + c]);
                        E[p * C + cc * 32 + c] = fmaf(m, inv_scale, __ldg(bias_sum + cc * 32 + c));
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc_ph[acc] ^= 1u;
            acc ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // no CTA leaves while its peer may still write to it
    if (warp == NGEN_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace tc
}  // namespace geob200

using namespace geob200;

int geob200_gse_embed_tc(const float* d_idx, const float* a_idx, long long n_pairs, int C, const float* div_term,
                         const float* Wd, const float* Wa, const float* bd, const float* ba, float* E, int mode,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (mode != 1 && mode != 2 && mode != 3 && mode != 4) return 1;
    if (C == 128 && mode == 3) {
        // hidden 128 (KITTI): the 3xFP16 kernel instantiated for N = 128, K = 2 x 128 (two angle + two distance chunks per tile)
        constexpr int CC = 128;
        const size_t img_halves = (size_t)CC * 2 * CC;
        const size_t need = 2 * img_halves * sizeof(__half) + (CC + 2) * sizeof(float) + 1024;
        GEOB_REQUIRE(workspace_bytes >= need, "gse_embed_tc: workspace too small (%zu < %zu)", workspace_bytes, need);
        __half* h_hi = (__half*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        __half* h_lo = h_hi + img_halves;
        float* bsum = (float*)(h_lo + img_halves);
        float* scale = bsum + CC;
        const long long n_tiles = (n_pairs + tc::PAIRS - 1) / tc::PAIRS;
        const int grid = (int)(n_tiles < (long long)num_sms() ? n_tiles : (long long)num_sms());
        tc::gse_absmax_kernel<CC><<<1, 1024, 0, st>>>(Wd, Wa, scale);
        tc::gse_pack_b_f16_kernel<CC><<<(unsigned)((img_halves + 255) / 256), 256, 0, st>>>(Wd, Wa, bd, ba, scale, h_hi, h_lo, bsum);
        if (ensure_max_smem((const void*)tc::gse_embed_f16_kernel<CC>)) return -1;
        tc::gse_embed_f16_kernel<CC><<<grid, tc::NTHREADS, tc::F16_SMEM, st>>>(d_idx, a_idx, n_pairs, div_term, h_hi, h_lo, bsum, scale, E);
        GEOB_CHECK_LAUNCH();
        count_launches(3);
        return 0;
    }
    if (C != tc::C) return 1;
    const size_t img_floats = (size_t)tc::C * tc::KTOT;
    const size_t need = sizeof(float) * (2 * img_floats + tc::C) + 1024;
    GEOB_REQUIRE(workspace_bytes >= need, "gse_embed_tc: workspace too small (%zu < %zu)", workspace_bytes, need);
    float* img_hi = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* img_lo = img_hi + img_floats;
    float* bias_sum = img_lo + img_floats;
    const long long n_tiles = (n_pairs + tc::PAIRS - 1) / tc::PAIRS;
    const int grid = (int)(n_tiles < (long long)num_sms() ? n_tiles : (long long)num_sms());
    if (mode == 3 || mode == 4) {
        __half* h_hi = (__half*)img_hi;
        __half* h_lo = h_hi + img_floats;
        float* bsum = (float*)(h_lo + img_floats);
        float* scale = bsum + tc::C;
        tc::gse_absmax_kernel<tc::C><<<1, 1024, 0, st>>>(Wd, Wa, scale);
        tc::gse_pack_b_f16_kernel<tc::C><<<(unsigned)((img_floats + 255) / 256), 256, 0, st>>>(Wd, Wa, bd, ba, scale, h_hi, h_lo, bsum);
        if (ensure_max_smem((const void*)tc::gse_embed_f16_kernel<tc::C>) || ensure_max_smem((const void*)tc::gse_embed_f16_cluster_kernel)) return -1;
        if (mode == 4) {
            // CTA pairs (cluster of 2) share every B chunk through TMA multicast; grid = even number of CTAs, one per SM
            int g2 = (int)((n_tiles + 1) / 2 < (long long)(num_sms() / 2) ? (n_tiles + 1) / 2 : (long long)(num_sms() / 2)) * 2;
            tc::gse_embed_f16_cluster_kernel<<<g2, tc::NTHREADS, tc::F16_SMEM, st>>>(d_idx, a_idx, n_pairs, div_term, h_hi, h_lo, bsum, scale, E);
        } else
            tc::gse_embed_f16_kernel<tc::C><<<grid, tc::NTHREADS, tc::F16_SMEM, st>>>(d_idx, a_idx, n_pairs, div_term, h_hi, h_lo, bsum, scale, E);
        GEOB_CHECK_LAUNCH();
        count_launches(3);
        return 0;
    }
    tc::gse_pack_b_kernel<<<(unsigned)((img_floats + 255) / 256), 256, 0, st>>>(Wd, Wa, bd, ba, img_hi, img_lo, bias_sum);
    if (mode == 1) {
        if (ensure_max_smem((const void*)tc::gse_embed_tc_kernel<3>)) return -1;
        tc::gse_embed_tc_kernel<3><<<grid, tc::NTHREADS, tc::Cfg<3>::SMEM, st>>>(d_idx, a_idx, n_pairs, div_term, img_hi, img_lo, bias_sum, E);
    } else {
        if (ensure_max_smem((const void*)tc::gse_embed_tc_kernel<1>)) return -1;
        tc::gse_embed_tc_kernel<1><<<grid, tc::NTHREADS, tc::Cfg<1>::SMEM, st>>>(d_idx, a_idx, n_pairs, div_term, img_hi, img_lo, bias_sum, E);
    }
    GEOB_CHECK_LAUNCH();
    count_launches(2);
    return 0;
}
