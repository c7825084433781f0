// vaxmc_kernels.cuh -- sm_100a device code of the Monte Carlo hot path.
//
// What each kernel replaces in the reference (/root/reference/model.py):
//   vx_param_stats_kernel   sample_param_combinations, rejection bookkeeping    :275-302
//   vx_sim_kernel<..>       run_single_realization (day loop)                    :29-136
//                           + the 30-day-shift/7-day-subsample weekly series     :196-214
//                           + (FOLD) the per-date sum / rank counting that feeds :220-223
//   vx_expected_kernel      run_single_realization with np.random.poisson := identity
//   vx_select_kernel        np.quantile's order statistics + ndarray.mean sums   :220-223
//   vx_extract_kernel       order statistics out of the merged window histograms
//
// Nothing here is a dense contraction: no tensor cores.  The stepper is instruction-issue
// bound (Philox + sampler arithmetic, whole day loop in registers), the DUMP variant and the
// select pass are HBM-bound streaming kernels.  See DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

namespace vx {

// ------------------------------------------------------------------ Philox4x32-10
// Counter-based RNG (Salmon et al., SC'11).  key = job seed; counter =
// (sample_lo, sample_hi, index, tag): a sample's stream depends only on its GLOBAL id, so
// results are bit-identical however the sample range is sharded over GPUs.
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

enum : uint32_t {
    TAG_DAY = 0,     // index = day >> 1: 4 words -> two Box-Muller pairs = the 2 x 2 normals of two days
    TAG_PRO_ANTI = 1,// index = rejection attempt: p_pro, p_anti (2 x 53-bit)
    TAG_PARAMS_A = 2,// pressure, tau
    TAG_PARAMS_B = 3,// nv_0, nv_max
    TAG_PTRS = 4,    // exact-Poisson validation mode: PTRS attempts
    TAG_INV = 5,     // exact-Poisson validation mode: index = day, inversion uniforms (words 0,1)
    TAG_TEST = 7
};

__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi)
{
#ifdef __CUDA_ARCH__
    // one IMAD.WIDE.U32 instead of IMAD.HI + IMAD
    asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0,%1}, p;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
    const uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
#endif
}

__host__ __device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                                      uint32_t &c3, uint32_t k0, uint32_t k1)
{
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(PHILOX_M0, c0, lo0, hi0);
    mulhilo32(PHILOX_M1, c2, lo1, hi1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Round keys precomputed on the host (key + r*W): in the stepper they are kernel-parameter
// constants, so each round is 2 IMAD.WIDE + 2 LOP3 with a constant-bank operand.
struct PhiloxKeys { uint32_t k[20]; };

__host__ __device__ inline PhiloxKeys philox_keys(uint64_t seed)
{
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { K.k[2 * r] = k0; K.k[2 * r + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
    return K;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, K.k[2 * r], K.k[2 * r + 1]);
    return make_uint4(c0, c1, c2, c3);
}

// 53-bit uniform in [0,1) from two words, the construction numpy's legacy stream uses
// (random_sample: (a>>5, b>>6) -> (a*2^26+b)/2^53), so lo+(hi-lo)*u has the same lattice as
// np.random.uniform (model.py:277-294).
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------ fast fp32 building blocks
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 do two lanes' worth of work in ONE issue slot;
// the stepper is issue-bound with the FMA pipes at ~40 %, so packing trades idle pipe for slots).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 splat2(float a) { return pack2(a, a); }

// Two independent N(0,1) from two words (Box-Muller).  u1 in (0,1], angle in [-pi,pi).
__device__ __forceinline__ void normal_pair(uint32_t w0, uint32_t w1, float &z0, float &z1)
{
    const float u1 = fmaf((float)w0, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    // -2 ln(u1) = (-2 ln 2) * lg2(u1)
    const float rad = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));   // u1 <= 1: argument >= +-0
    const float ang = (float)(int32_t)w1 * 1.4629180792671596e-09f; // pi * 2^-31
    z0 = rad * fast_cos(ang);
    z1 = rad * fast_sin(ang);
}

// 1/k for the inversion recurrence p_k = p_{k-1} * lam / k
__constant__ float c_rcp[72];

// Phi(z) in fp32 (Abramowitz & Stegun 26.2.17, |error| < 7.5e-8 plus fp32 rounding): the uniform
// of the inversion regime.  z ~ N(0,1) makes Phi(z) ~ U(0,1), and it is monotone in z like the
// expansion below, so one normal per draw serves both regimes and a Philox block (4 words = two
// Box-Muller pairs) covers TWO days.
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float normal_cdf(float z)
{
    const float x = fabsf(z);
    const float t = fast_rcp(fmaf(0.2316419f, x, 1.0f));
    float poly = fmaf(t, 1.330274429f, -1.821255978f);
    poly = fmaf(t, poly, 1.781477937f);
    poly = fmaf(t, poly, -0.356563782f);
    poly = fmaf(t, poly, 0.319381530f);
    // phi(x) = exp(-x^2/2)/sqrt(2 pi) = 2^(-x^2 log2(e)/2) / sqrt(2 pi)
    const float tail = (0.3989422804f * fast_ex2(x * x * -0.7213475204f)) * (t * poly);
    return z > 0.0f ? 1.0f - tail : tail;
}

// Poisson(lam) variate from ONE standard normal z.
//   lam == 0          -> 0 (numpy returns 0 without consuming the stream, model.py:104,116)
//   0 < lam < 2       -> exact inversion by sequential search on the uniform Phi(z)
//   lam >= 2          -> floor of the continuous inverse Q(z) of the Poisson CDF expanded
//                        around the normal quantile z (Cornish-Fisher with the lattice
//                        correction; the first three orders are Giles' Q_N1, ACM TOMS 42(1),
//                        2016):
//                          Q = lam + s z + (1/3 + z^2/6) + (-z/36 - z^3/72)/s
//                              + (-8/405 + 7 z^2/810 + z^4/270)/lam
//                              + (a1 z + a3 z^3 + a5 z^5)/lam^1.5
//                              + (b0 + b2 z^2 + b4 z^4 + b6 z^6)/lam^2,      s = sqrt(lam)
//                        The two highest orders are a weighted least-squares fit of the exact
//                        jump points Phi^-1(F(k)) -> k+1 over lam in [2,32] (their asymptotic
//                        values are a=(.01725,-.00352,-.00133), b=(-.0015,-.0113,.0013,.0006));
//                        oracle/fit_poisson_cf.py regenerates them and prints the resulting
//                        sup-CDF error against the exact law: <=1.5e-4 on [2,3], 6.4e-5 at 3.5,
//                        <=4.8e-5 for lam>=4, <=1.5e-5 for lam>=10, <2e-6 for lam>=32, and the
//                        bias of the first two moments: |E X - lam| <= 1.2e-4 counts (7.4e-5 sd),
//                        |Var X / lam - 1| <= 3e-4 (tests/test_sampler_math.py pins these; a
//                        refit that weights the moments buys a 5x smaller bias with 30-130 %
//                        relative errors in the far tail at lam ~ 2, so it was not adopted --
//                        see the script).  The threshold is where
//                        the divergent inversion loop stops paying: at 10 (numpy's) it was
//                        43 % of all issued instructions, at 4 it was 10 %, at 2 it is 4 %.
constexpr float POISSON_INV_MAX = 2.0f;
constexpr float CF_A1 = 0.01865529f, CF_A3 = 0.00004626f, CF_A5 = -0.00184102f;
constexpr float CF_B0 = 0.00100592f, CF_B2 = -0.01903893f, CF_B4 = 0.00285482f, CF_B6 = 0.00029421f;

// The z-only part of the expansion for BOTH draws of a day at once (packed fp32x2):
// c0..c4 of  Q = lam + s z + c0 + w (c1 + w (c2 + w (c3 + w c4))),  w = lam^-1/2.
struct CFCoef { f32x2 c0, c1, c2, c3, c4; };

__device__ __forceinline__ CFCoef cf_coefficients(float z1, float z2)
{
    // (no clamp of z in the two fitted orders: |z| > 5 has probability 5.7e-7 and the law's error
    //  there is the same with and without it -- oracle/fit_poisson_cf.py prints both)
    const f32x2 z = pack2(z1, z2), z2p = mul2(z, z);
    CFCoef c;
    c.c0 = fma2(z2p, splat2(0.16666667f), splat2(0.33333334f));
    c.c1 = mul2(z, fma2(z2p, splat2(-0.013888889f), splat2(-0.027777778f)));
    c.c2 = fma2(z2p, fma2(z2p, splat2(0.0037037036f), splat2(0.0086419750f)), splat2(-0.019753087f));
    c.c3 = mul2(z, fma2(z2p, fma2(z2p, splat2(CF_A5), splat2(CF_A3)), splat2(CF_A1)));
    c.c4 = fma2(z2p, fma2(z2p, fma2(z2p, splat2(CF_B6), splat2(CF_B4)), splat2(CF_B2)), splat2(CF_B0));
    return c;
}

// Sequential-search inversion on u in [0,1): smallest x with u < F(x; lam), 0 < lam < 2.
__device__ __forceinline__ int poisson_invert(float lam, float u, int kmax)
{
    u = fminf(u, 0.99999994f);
    float p = fast_ex2(lam * -1.4426950408889634f), cdf = p;
    int x = 0;
    while (u >= cdf && x < kmax) {
        ++x;
        p *= lam * c_rcp[x];
        cdf += p;
    }
    return x;
}

// The Poisson variate of the regime table above, given the pair's coefficients (WHICH = 0: first
// draw of the day, 1: second).
// `cap`: the caller clamps the variate to [0, cap] anyway (model.py:105-108, 117-118), so the
// sequential search of the inversion regime may stop there (late in the campaign lam1 ~ 1.7 n_wait
// with n_wait = 0, 1, 2: one comparison instead of a loop).
template <int WHICH>
__device__ __forceinline__ int poisson_draw_c(float lam, float z, const CFCoef &c, int cap = 40)
{
    float a0, b0, a1, b1, a2, b2, a3, b3, a4, b4;
    unpack2(c.c0, a0, b0); unpack2(c.c1, a1, b1); unpack2(c.c2, a2, b2);
    unpack2(c.c3, a3, b3); unpack2(c.c4, a4, b4);
    const float c0 = WHICH ? b0 : a0, c1 = WHICH ? b1 : a1, c2 = WHICH ? b2 : a2, c3 = WHICH ? b3 : a3,
                c4 = WHICH ? b4 : a4;
    const float w1 = fast_rsqrt(lam), s = lam * w1;
    const float tail = fmaf(w1, fmaf(w1, fmaf(w1, fmaf(w1, c4, c3), c2), c1), c0);
    int x = max(__float2int_rd(fmaf(s, z, lam) + tail), 0);
    if (lam < POISSON_INV_MAX) {
        x = 0;
        if (lam > 0.0f) x = poisson_invert(lam, normal_cdf(z), min(cap, 40));
    }
    return x;
}

// Exact Poisson(lam) variate -- validation mode ("exact_poisson" option), numpy's regime
// structure on the Philox stream: lam < 10 inversion (exact up to fp32 rounding of the CDF),
// lam >= 10 Hoermann's PTRS transformed rejection ("The transformed rejection method for
// generating Poisson random variables", 1993 -- the algorithm numpy's legacy poisson uses).
// The acceptance test is evaluated in fp64 (k*log(lam) - lgamma(k+1) cancels catastrophically
// in fp32 at lam ~ 1e4).  Its uniforms come from Philox blocks of their own (TAG_INV, index = day,
// for the inversion; TAG_PTRS, index = 2*(8*day + attempt block) + draw, for the rejection loop);
// ~11x slower than the default sampler on the bench job.

__device__ __noinline__ int poisson_draw_exact(float lam, uint32_t s_lo, uint32_t s_hi,
                                               uint32_t day, uint32_t which, const PhiloxKeys &K)
{
    if (!(lam > 0.0f)) return 0;
    if (lam < 10.0f) {
        const uint4 r = philox4x32_10(s_lo, s_hi, day, TAG_INV, K);
        return poisson_invert(lam, (float)(which ? r.y : r.x) * 2.3283064365386963e-10f, 70);
    }
    const double L = (double)lam, slam = sqrt(L), loglam = log(L);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
    for (uint32_t blk = 0;; ++blk) {
        const uint4 r = philox4x32_10(s_lo, s_hi, 2u * (8u * day + (blk & 7u)) + which + 16384u * (blk >> 3),
                                      TAG_PTRS, K);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double U = ((double)(h ? r.z : r.x) + 0.5) * 2.3283064365386963e-10 - 0.5;
            const double V = ((double)(h ? r.w : r.y) + 0.5) * 2.3283064365386963e-10;
            const double us = 0.5 - fabs(U);
            const double kf = floor((2 * a / us + b) * U + L + 0.43);
            if (us >= 0.07 && V <= vr) return (int)kf;
            if (kf < 0 || (us < 0.013 && V > us)) continue;
            if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -L + kf * loglam - lgamma(kf + 1))
                return (int)kf;
        }
    }
}

// ------------------------------------------------------------------ parameter draw
struct SampleParams { double p_pro, p_anti, pressure, tau, nv0, nvmax; };

// model.py:275-300 for ONE sample: redraw (p_pro,p_anti) until p_pro+p_anti <= 1, then the
// other four.  Returns the number of rejections this sample consumed.
__device__ __forceinline__ uint32_t draw_sample_params(const double *__restrict__ b, uint32_t k0,
                                                       uint32_t k1, uint64_t sample,
                                                       SampleParams &p, uint32_t max_attempts)
{
    const uint32_t s_lo = (uint32_t)sample, s_hi = (uint32_t)(sample >> 32);
    uint32_t j = 0;
    for (;; ++j) {
        const uint4 r = philox4x32_10(s_lo, s_hi, j, TAG_PRO_ANTI, k0, k1);
        p.p_pro = b[0] + (b[1] - b[0]) * u53(r.x, r.y);
        p.p_anti = b[2] + (b[3] - b[2]) * u53(r.z, r.w);
        if (!(p.p_pro + p.p_anti > 1.0) || j + 1 >= max_attempts) break;
    }
    const uint4 ra = philox4x32_10(s_lo, s_hi, 0u, TAG_PARAMS_A, k0, k1);
    const uint4 rb = philox4x32_10(s_lo, s_hi, 0u, TAG_PARAMS_B, k0, k1);
    p.pressure = b[4] + (b[5] - b[4]) * u53(ra.x, ra.y);
    p.tau = b[6] + (b[7] - b[6]) * u53(ra.z, ra.w);
    p.nv0 = b[8] + (b[9] - b[8]) * u53(rb.x, rb.y);
    p.nvmax = b[10] + (b[11] - b[10]) * u53(rb.z, rb.w);
    return j;
}

struct Bounds { double b[12]; };

// Rejection bookkeeping + moments of p_soft_no = 1-(p_pro+p_anti) (model.py:280-302,339-342)
// over samples [first, first+count).  Per-block partials (deterministic tree), reduced by
// vx_reduce_stats_kernel.  `stop` is a device flag/counter pair: [0] running total of
// rejections (pushed every 64), [1] limit; when the total passes the limit every thread
// bails out (the reference aborts at 10*n_rep rejections, model.py:283-287).
__global__ void __launch_bounds__(256) vx_param_stats_kernel(Bounds bd, uint64_t seed,
                                                             int64_t first, int64_t count,
                                                             unsigned long long *stop,
                                                             double *block_part /*[grid][3]*/)
{
    __shared__ double sh[3][256];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double rej = 0.0, s1 = 0.0, s2 = 0.0;
    if (i < count) {
        const uint64_t sample = (uint64_t)(first + i);
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        const uint32_t s_lo = (uint32_t)sample, s_hi = (uint32_t)(sample >> 32);
        const unsigned long long limit = stop[1];
        uint32_t j = 0, pending = 0;
        double pro, anti;
        bool bail = false;
        for (;; ++j) {
            const uint4 r = philox4x32_10(s_lo, s_hi, j, TAG_PRO_ANTI, k0, k1);
            pro = bd.b[0] + (bd.b[1] - bd.b[0]) * u53(r.x, r.y);
            anti = bd.b[2] + (bd.b[3] - bd.b[2]) * u53(r.z, r.w);
            if (!(pro + anti > 1.0)) break;
            if (++pending == 64) {
                const unsigned long long t = atomicAdd(&stop[0], 64ull) + 64ull;
                pending = 0;
                if (t > limit) { bail = true; ++j; break; }
            }
            if (j == 0xfffffff0u) { bail = true; break; }
        }
        if (pending) atomicAdd(&stop[0], (unsigned long long)pending);
        rej = (double)j;
        if (!bail) {
            const double s = 1 - (pro + anti);
            s1 = s; s2 = s * s;
        }
    }
    sh[0][threadIdx.x] = rej; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + off];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + off];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) block_part[(int64_t)blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(1024) vx_reduce_stats_kernel(const double *block_part,
                                                               int64_t nblocks, double *out3)
{
    __shared__ double sh[3][1024];
    double a[3] = {0.0, 0.0, 0.0};
    for (int64_t i = threadIdx.x; i < nblocks; i += 1024)
        for (int k = 0; k < 3; ++k) a[k] += block_part[i * 3 + k];
    for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if (threadIdx.x < off)
            for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x < 3) out3[threadIdx.x] = sh[threadIdx.x][0];
}

// This shard's statistics (vx_reduce_stats_kernel's out3 = rejections, sum s, sum s^2) into header
// words 2..5 of the accumulator slab as integers / fixed point, so that they are merged by the
// same all-reduce as the histograms and never visit the host on their own.
__global__ void vx_stats_to_header_kernel(const double *out3, long long count, double fix,
                                          unsigned long long *acc)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        acc[2] = (unsigned long long)out3[0];
        acc[3] = (unsigned long long)llround(out3[1] * fix);
        acc[4] = (unsigned long long)llround(out3[2] * fix);
        acc[5] = (unsigned long long)count;
    }
}

// Parameter tuples as drawn on the device, for sample_param_combinations' return value.
__global__ void __launch_bounds__(256) vx_sample_params_kernel(Bounds bd, uint64_t seed,
                                                               int64_t first, int64_t count,
                                                               double *params /*[count][6]*/,
                                                               double *soft /*[count] or null*/)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    SampleParams p;
    draw_sample_params(bd.b, (uint32_t)seed, (uint32_t)(seed >> 32), (uint64_t)(first + i), p,
                       1u << 20);
    double *o = params + i * 6;
    o[0] = p.p_pro; o[1] = p.p_anti; o[2] = p.pressure; o[3] = p.tau; o[4] = p.nv0; o[5] = p.nvmax;
    if (soft) soft[i] = 1 - (p.p_pro + p.p_anti);
}

// ------------------------------------------------- regime sort (counting sort by pressure)
// lam2 = n_agn * f * pressure and, late in the campaign, lam1 ~ the daily conversions: both
// scale with the sample's pressure, so the lanes of a warp share a Poisson regime if the warp's
// samples have similar pressure.  The fold does not care about sample order (integer sums), so
// the local samples are bucketed by the top 10 bits of the very word the pressure is drawn
// from (monotone in the pressure, no bounds needed).  Unsorted, the divergent inversion loops
// run 4.7 iterations per warp-day with 2-3 lanes active; sorted, 1.75 (SURVEY-style numpy
// experiment reproduced the ncu counts).
constexpr int SORT_BUCKETS = 1024;

__device__ __forceinline__ uint32_t pressure_bucket(uint64_t seed, uint64_t sample)
{
    const uint4 ra = philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), 0u, TAG_PARAMS_A,
                                   (uint32_t)seed, (uint32_t)(seed >> 32));
    return ra.x >> 22;
}

__global__ void __launch_bounds__(256) vx_bucket_count_kernel(uint64_t seed, int64_t first,
                                                              int64_t count, uint32_t *gcount)
{
    __shared__ uint32_t sh[SORT_BUCKETS];
    for (int b = threadIdx.x; b < SORT_BUCKETS; b += 256) sh[b] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < count) atomicAdd(&sh[pressure_bucket(seed, (uint64_t)(first + i))], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < SORT_BUCKETS; b += 256)
        if (sh[b]) atomicAdd(&gcount[b], sh[b]);
}

__global__ void __launch_bounds__(SORT_BUCKETS) vx_bucket_scan_kernel(uint32_t *gcount)
{
    __shared__ uint32_t sh[SORT_BUCKETS];
    const int t = threadIdx.x;
    const uint32_t own = gcount[t];
    sh[t] = own;
    __syncthreads();
    for (int off = 1; off < SORT_BUCKETS; off <<= 1) {
        const uint32_t v = (t >= off) ? sh[t - off] : 0u;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    gcount[t] = sh[t] - own;      // exclusive prefix = first slot of the bucket
}

__global__ void __launch_bounds__(256) vx_bucket_scatter_kernel(uint64_t seed, int64_t first,
                                                                int64_t count, uint32_t *goffset,
                                                                int32_t *perm)
{
    __shared__ uint32_t sh[SORT_BUCKETS];
    for (int b = threadIdx.x; b < SORT_BUCKETS; b += 256) sh[b] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t b = 0, r = 0;
    if (i < count) { b = pressure_bucket(seed, (uint64_t)(first + i)); r = atomicAdd(&sh[b], 1u); }
    __syncthreads();
    for (int k = threadIdx.x; k < SORT_BUCKETS; k += 256) {
        const uint32_t c = sh[k];
        sh[k] = c ? atomicAdd(&goffset[k], c) : 0u;   // this block's slice of bucket k
    }
    __syncthreads();
    if (i < count) perm[sh[b] + r] = (int32_t)i;
}

// ------------------------------------------------------------------------ stepper
// Window descriptor of the streaming fold: values v with 0 <= v-a < len are histogrammed
// at bin (v-a)>>shift into acc[hist + off + bin]; values v < a are counted.
struct __align__(16) Win { int a; uint32_t len; uint32_t shift; uint32_t off; };

struct SimArgs {
    Bounds bd;
    const double *params;   // explicit tuples [count][6] (then bd is unused) or nullptr
    uint64_t seed;
    PhiloxKeys keys;        // philox_keys(seed)
    int64_t first;          // GLOBAL id of local sample 0 (Philox counter)
    int64_t count;          // samples in this launch
    int64_t N;
    int T, W;
    int32_t *dump;          // DUMP: [R][dump_stride] raw rows, column = position in this launch
    int64_t dump_stride;
    const int32_t *perm;    // optional order of the local samples (pressure-sorted): position -> local id, or nullptr
    const Win *wins;        // FOLD: [R][2]
    unsigned long long *acc;// FOLD: header(8) | sums[R] | below[R][2] | hist bins
    int bins64;             // FOLD: histogram bins are 64-bit words (global n >= 2^32), else 32-bit
};

// One more count in histogram bin `idx` of the slab.  Bins are 32-bit unless the job has 2^32 or
// more samples: they are packed two per 64-bit slab word and, below 2^32 samples, can never
// carry into their neighbour, so the ranks still merge the whole slab with ONE int64
// all-reduce(sum).
__device__ __forceinline__ void bin_add(unsigned long long *hist, uint32_t idx, int bins64, uint32_t count = 1u)
{
    if (bins64) atomicAdd(hist + idx, (unsigned long long)count);
    else atomicAdd(reinterpret_cast<unsigned int *>(hist) + idx, count);
}

// The same as a PREDICATED reduction (no branch): the stepper's fold revisits quads of samples in
// which some value hit a window and tests all eight (value, window) pairs; a branch per pair
// costs more issue slots than the eight predicated REDs.
__device__ __forceinline__ void bin_add_if(bool pred, unsigned long long *hist, uint32_t idx, int bins64)
{
    if (bins64)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.global.add.u64 [%1], 1;\n\t}"
                     :: "r"((uint32_t)pred), "l"(hist + idx) : "memory");
    else
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.global.add.u32 [%1], 1;\n\t}"
                     :: "r"((uint32_t)pred), "l"(reinterpret_cast<unsigned int *>(hist) + idx) : "memory");
}

constexpr int ACC_HDR = 8;
constexpr int WIN_DAYS = 14;       // one warp tile = 2 weeks = 14+14+2+2 = 32 rows
constexpr int TILE_LD = 36;        // padded, 16 B-aligned rows: sample-major STS and row-major LDS.128 are both conflict-free
constexpr int SIM_THREADS = 128;       // warps are independent: small CTAs = short tail
#ifndef VX_SIM_MINBLOCKS
#define VX_SIM_MINBLOCKS 8            // CTAs per SM the register allocation is capped for
#endif

enum { SIM_DUMP = 0, SIM_FOLD = 1 };

// PIPE = 1 is the LATENCY variant for launches that cannot fill the GPU (the pilot: 128 CTAs, one
// warp per scheduler, a serial T-day chain): the state-independent half of a day pair -- Philox,
// Box-Muller, the packed z-polynomials -- is computed one pair ahead, so that it sits in the
// same straight-line code as the state-dependent chain of the current pair and fills its stall
// slots.  It needs ~50 more registers, which only an occupancy-bound launch would miss.
template <int MODE, int EXACT = 0, int PIPE = 0>
__global__ void __launch_bounds__(SIM_THREADS, EXACT ? 4 : (PIPE ? 2 : VX_SIM_MINBLOCKS)) vx_sim_kernel(const SimArgs a)
{
    __shared__ __align__(16) int tile_all[MODE == SIM_FOLD ? (SIM_THREADS / 32) * 32 * TILE_LD : 4];
    __shared__ double arr_sm[3][SIM_THREADS];   // nv0, nvmax, tau*7: read once a week (until the cap is reached)
    __shared__ int q_sm[5][SIM_THREADS];        // delay line of delta_n_vacc, one push a week
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (a persistent launch with warps claiming 32-sample batches from a counter was tried:
    //  4 % slower than letting the hardware schedule 128-thread CTAs)
    const int64_t warp_base = (int64_t)blockIdx.x * SIM_THREADS + warp * 32;
    if (warp_base >= a.count) return;                       // warp-uniform
    const int nvalid = (int)min((int64_t)32, a.count - warp_base);
    const bool live = lane < nvalid;
    const int64_t col = warp_base + (live ? lane : nvalid - 1);   // dead lanes shadow a live one
    const int64_t local = a.perm ? (int64_t)a.perm[col] : col;    // position -> local sample id
    const uint64_t sample = (uint64_t)(a.first + local);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const uint32_t s_lo = (uint32_t)sample, s_hi = (uint32_t)(sample >> 32);
    int *tile = tile_all + (MODE == SIM_FOLD ? warp * 32 * TILE_LD : 0);

    SampleParams p;
    if (a.params) {
        const double *src = a.params + local * 6;
        p.p_pro = src[0]; p.p_anti = src[1]; p.pressure = src[2];
        p.tau = src[3]; p.nv0 = src[4]; p.nvmax = src[5];
    } else {
        draw_sample_params(a.bd.b, k0, k1, sample, p, 1u << 20);
    }

    // model.py:63-75.  int() truncates toward zero.
    const double Nd = (double)a.N;
    const double p_agn = 1 - (p.p_pro + p.p_anti);
    int n_wait = (int)(p.p_pro * Nd);
    int n_agn = (int)(p_agn * Nd);
    int stock = 0, n_vacc = 0;
    const double tau7 = p.tau * 7;
    const double ln2 = 0.6931471805599453;                 // np.log(2)
    const bool monotone = (p.tau > 0.0) && (p.nv0 > 0.0);
    bool capped = false;
    const int arr_cap = (int)(p.nvmax * Nd);
    const float c1 = (float)((2.0 / 7.0) / Nd);            // (2/7)/N   model.py:102
    const float c2 = (float)(p.pressure / Nd);             // pressure/N model.py:112-115
    // delta_n_vacc of days = 5 (mod 7), 5 weeks deep: the weekly series adds the value of day 7k-30,
    // i.e. the push of week k-5, to day 7k's (model.py:196-209)
    // ... kept as a ring in shared memory (slot = week % 5; one load and one store a week), together
    // with the fp64 arrival parameters (read once a week): 11 registers the daily chain does not
    // need, which leaves ptxas room to schedule it (62 -> 55 registers, fold 1.7 % faster on B200)
    arr_sm[0][threadIdx.x] = p.nv0; arr_sm[1][threadIdx.x] = p.nvmax; arr_sm[2][threadIdx.x] = tau7;
#pragma unroll
    for (int k = 0; k < 5; ++k) q_sm[k][threadIdx.x] = 0;
    auto q_oldest = [&](const int week) { return q_sm[week % 5][threadIdx.x]; };
    auto q_push = [&](const int week, const int dv) { q_sm[week % 5][threadIdx.x] = dv; };

    const int T = a.T, W = a.W;
    struct PairRng { float z1, z2, z3, z4; CFCoef cfA, cfB; };
    auto pair_ahead = [&](const uint32_t pair) {
        PairRng o;
        const uint4 r = philox4x32_10(s_lo, s_hi, pair, TAG_DAY, a.keys);
        normal_pair(r.x, r.y, o.z1, o.z2);
        normal_pair(r.z, r.w, o.z3, o.z4);
        o.cfA = cf_coefficients(o.z1, o.z2);
        o.cfB = cf_coefficients(o.z3, o.z4);
        return o;
    };
    PairRng nx;
    if (PIPE) nx = pair_ahead(0u);
    // DUMP: this sample's column in the rows of day 0 (n_vacc) and week 0 (doses)
    int32_t *dump_day = MODE == SIM_DUMP ? a.dump + col : nullptr;
    int32_t *dump_week = MODE == SIM_DUMP ? a.dump + (int64_t)a.T * a.dump_stride + col : nullptr;
    const int64_t dump_stock_off = ((int64_t)a.T + 2 * a.W) * a.dump_stride, dump_recv_off = (int64_t)a.W * a.dump_stride;
    for (int d0 = 0; d0 < T; d0 += WIN_DAYS) {
        int *tp = tile + lane;                              // this sample's column of the tile
        const int nd = min(WIN_DAYS, T - d0);               // days in this tile
        // ---- weekly arrivals, model.py:69, 85-92 (fp64: the truncation must match)
        auto arrivals = [&](const int day) {
            int arr;
            if (capped) {
                arr = arr_cap;
            } else {
                const double nv0 = arr_sm[0][threadIdx.x], nvmax = arr_sm[1][threadIdx.x], t7 = arr_sm[2][threadIdx.x];
                const double g = nv0 * exp(ln2 * (double)day / t7);
                const bool over = nvmax < g;                // Python min(g, nv_max)
                arr = (int)((over ? nvmax : g) * Nd);
                capped = over && monotone;
            }
            stock += arr;                                   // received == n_vacc + stock at all times
        };
        // one day of model.py:82-127; (z1, z2) = the day's two normals.  WEEK0: day % 7 == 0
        // (arrivals already added; the weekly rows are recorded), wk = week slot of the tile
        auto day_step = [&](const int day, const float z1, const float z2, const CFCoef &cf, const bool week0, const int wk) {
            // ---- vaccination draw, model.py:102-112
            const float lam1 = (float)n_wait * ((float)stock * c1);
            const int cap1 = min(stock, n_wait);
            int dv = EXACT ? poisson_draw_exact(lam1, s_lo, s_hi, (uint32_t)day, 0u, a.keys)
                           : poisson_draw_c<0>(lam1, z1, cf, cap1);
            dv = min(dv, cap1);
            n_vacc += dv; n_wait -= dv; stock -= dv;
            // ---- conversion draw, model.py:115-120
            const float lam2 = (float)n_agn * ((float)n_vacc * c2);
            int da = EXACT ? poisson_draw_exact(lam2, s_lo, s_hi, (uint32_t)day, 1u, a.keys)
                           : poisson_draw_c<1>(lam2, z2, cf, n_agn);
            da = min(da, n_agn);
            n_agn -= da; n_wait += da;
            // ---- recording, model.py:124-127 and the weekly series :196-214
            if (MODE == SIM_FOLD) {
                tp[0] = n_vacc;
                tp[14 * TILE_LD] = stock;
                tp += TILE_LD;
                if (week0) {
                    tile[(28 + wk) * TILE_LD + lane] = dv + q_oldest(d0 / 7 + wk);
                    tile[(30 + wk) * TILE_LD + lane] = n_vacc + stock;
                }
            } else {
                // row pointers advance by one row a day / a week (no 64-bit index arithmetic per store)
                if (live) {
                    dump_day[0] = n_vacc;
                    dump_day[dump_stock_off] = stock;
                    if (week0) {
                        dump_week[0] = dv + q_oldest(d0 / 7 + wk);
                        dump_week[dump_recv_off] = n_vacc + stock;
                    }
                }
                dump_day += a.dump_stride;
                if (week0) dump_week += a.dump_stride;
            }
            return dv;
        };
        // Two days per Philox block (tile days 2pp and 2pp+1): the pair loop is kept rolled (an
        // unrolled week, ~47 KB, misses the instruction cache and measured 1 % slower); the
        // events of the week (arrivals at tile days 0 and 7, the delay-line shift at 5 and 12)
        // are warp-uniform branches on pp.
#pragma unroll 1
        for (int pp = 0; 2 * pp < nd; ++pp) {
            const int dayA = d0 + 2 * pp;
            if (PIPE) {
                const PairRng cur = nx;
                nx = pair_ahead(((uint32_t)dayA >> 1) + 1u);                     // next pair, state-independent
                if (pp == 0) arrivals(dayA);
                const int dvA = day_step(dayA, cur.z1, cur.z2, cur.cfA, pp == 0, 0);
                if (pp == 6) q_push(d0 / 7 + 1, dvA);
                if (2 * pp + 1 < nd) {
                    if (pp == 3) arrivals(dayA + 1);
                    const int dvB = day_step(dayA + 1, cur.z3, cur.z4, cur.cfB, pp == 3, 1);
                    if (pp == 2) q_push(d0 / 7, dvB);
                }
            } else {
                const uint4 r = philox4x32_10(s_lo, s_hi, (uint32_t)dayA >> 1, TAG_DAY, a.keys);
                float z1, z2;
                normal_pair(r.x, r.y, z1, z2);
                if (pp == 0) arrivals(dayA);
                const int dvA = day_step(dayA, z1, z2, cf_coefficients(z1, z2), pp == 0, 0);
                if (pp == 6) q_push(d0 / 7 + 1, dvA);                             // tile day 12 = 5 (mod 7)
                if (2 * pp + 1 < nd) {
                    // day B's normals are made only now: ten fewer live registers across day A
                    normal_pair(r.z, r.w, z1, z2);
                    if (pp == 3) arrivals(dayA + 1);                                // tile day 7
                    const int dvB = day_step(dayA + 1, z1, z2, cf_coefficients(z1, z2), pp == 3, 1);
                    if (pp == 2) q_push(d0 / 7, dvB);                            // tile day 5
                }
            }
        }
        if (MODE == SIM_FOLD) {
            // ---- fold: lane r owns tile row r (one date of one series) over the warp's samples
            __syncwarp();
            int grow = -1;
            if (lane < 14) { const int d = d0 + lane; if (d < T) grow = d; }
            else if (lane < 28) { const int d = d0 + lane - 14; if (d < T) grow = T + 2 * W + d; }
            else if (lane < 30) { const int k = d0 / 7 + lane - 28; if (k < W) grow = T + k; }
            else { const int k = d0 / 7 + lane - 30; if (k < W) grow = T + W + k; }
            if (grow >= 0) {
                const int R = 2 * T + 2 * W;
                const Win wl = a.wins[grow * 2], wh = a.wins[grow * 2 + 1];
                unsigned long long *hist = a.acc + ACC_HDR + 3 * (int64_t)R;
                uint32_t sum = 0, clo = 0, chi = 0;
                const int *row = tile + lane * TILE_LD;
                // a value inside one of the row's two windows: one more count in its bin
                auto hit = [&](int v) {
                    const uint32_t dl = (uint32_t)(v - wl.a), dh = (uint32_t)(v - wh.a);
                    if (dl < wl.len) bin_add(hist, wl.off + (dl >> wl.shift), a.bins64);
                    if (dh < wh.len) bin_add(hist, wh.off + (dh >> wh.shift), a.bins64);
                };
                // In-window values are rare (the windows hold ~1-3 % of the mass), but with 32
                // rows in flight SOME lane has one in almost every step: the scan is branch-free
                // and only marks the samples that hit (one bit each); the marked values are revisited
                // afterwards, so the divergent part runs max-over-lanes(#hits) short iterations per
                // tile (~4) instead of a long one per scan step (8).
                uint32_t hits = 0;
                auto scan4 = [&](const int s) {
                    const int4 v4 = *reinterpret_cast<const int4 *>(row + s);
                    const int v0 = v4.x, v1 = v4.y, v2 = v4.z, v3 = v4.w;
                    sum += (uint32_t)v0 + (uint32_t)v1 + (uint32_t)v2 + (uint32_t)v3;
                    const uint32_t l0 = v0 - wl.a, l1 = v1 - wl.a, l2 = v2 - wl.a, l3 = v3 - wl.a;
                    const uint32_t h0 = v0 - wh.a, h1 = v1 - wh.a, h2 = v2 - wh.a, h3 = v3 - wh.a;
                    clo += (l0 >> 31) + (l1 >> 31) + (l2 >> 31) + (l3 >> 31);
                    chi += (h0 >> 31) + (h1 >> 31) + (h2 >> 31) + (h3 >> 31);
#ifndef VX_HITS_PER_QUAD
                    if (l0 < wl.len || h0 < wh.len) hits |= 1u << s;
                    if (l1 < wl.len || h1 < wh.len) hits |= 2u << s;
                    if (l2 < wl.len || h2 < wh.len) hits |= 4u << s;
                    if (l3 < wl.len || h3 < wh.len) hits |= 8u << s;
#else
                    // alternative: one combined test per 4 samples, quads revisited (10 fewer instructions per
                    // scan step, ~2.5x longer revisit loop: measured 0.8 % slower on B200)
                    const uint32_t ml = min(min(l0, l1), min(l2, l3)), mh = min(min(h0, h1), min(h2, h3));
                    if (ml < wl.len || mh < wh.len) hits |= 1u << (s >> 2);
#endif
                };
                if (nvalid == 32) {
#pragma unroll
                    for (int s = 0; s < 32; s += 4) scan4(s);
                } else {
                    for (int s = 0; s < nvalid; ++s) {
                        const int v = row[s];
                        sum += (uint32_t)v;
                        clo += (uint32_t)(v - wl.a) >> 31;
                        chi += (uint32_t)(v - wh.a) >> 31;
                        hit(v);
                    }
                }
#ifndef VX_HITS_PER_QUAD
                if (__popc(hits) >= 16) {
                    // a row whose samples (almost) all sit in a window is a flat or nearly flat row --
                    // single-value bounds make the delivery rows deterministic, for one: run-length
                    // aggregate the bins, or 32 atomics per warp and tile would queue on one address
                    uint32_t pl = 0, ph = 0, nl = 0, nh = 0;
                    while (hits) {
                        const int v = row[__ffs((int)hits) - 1];
                        hits &= hits - 1;
                        const uint32_t dl = (uint32_t)(v - wl.a), dh = (uint32_t)(v - wh.a);
                        if (dl < wl.len) {
                            const uint32_t b = wl.off + (dl >> wl.shift);
                            if (nl && b != pl) { bin_add(hist, pl, a.bins64, nl); nl = 0; }
                            pl = b; ++nl;
                        }
                        if (dh < wh.len) {
                            const uint32_t b = wh.off + (dh >> wh.shift);
                            if (nh && b != ph) { bin_add(hist, ph, a.bins64, nh); nh = 0; }
                            ph = b; ++nh;
                        }
                    }
                    if (nl) bin_add(hist, pl, a.bins64, nl);
                    if (nh) bin_add(hist, ph, a.bins64, nh);
                }
#endif
                while (hits) {
                    const int sidx = __ffs((int)hits) - 1;
                    hits &= hits - 1;
#ifndef VX_HITS_PER_QUAD
                    hit(row[sidx]);
#else
                    const int4 v4 = *reinterpret_cast<const int4 *>(row + 4 * sidx);
                    hit(v4.x); hit(v4.y); hit(v4.z); hit(v4.w);
#endif
                }
                atomicAdd(a.acc + ACC_HDR + grow, (unsigned long long)sum);
                if (clo) atomicAdd(a.acc + ACC_HDR + R + 2 * grow, (unsigned long long)clo);
                if (chi) atomicAdd(a.acc + ACC_HDR + R + 2 * grow + 1, (unsigned long long)chi);
            }
            __syncwarp();
        }
    }
    if (MODE == SIM_FOLD && lane == 0)
        atomicAdd(a.acc + 0, (unsigned long long)nvalid);  // header word 0: samples folded
}

// Fold of an already materialised block of raw rows (the pilot matrix, and the samples stepped
// beside the pilot) into the accumulator slab: same sums / counts-below / window histograms as the
// FOLD stepper produces, so those samples are not simulated twice.  HBM-streaming: grid =
// (row, column chunk), every thread has its FM_VEC 16-byte loads in flight before it touches a
// value; the launch runs on the low-priority side stream beside the issue-bound fold stepper (the
// slab only ever sees integer atomics, so the two commute).
constexpr int FM_THREADS = 256, FM_VEC = 8, FM_COLS = FM_THREADS * FM_VEC * 4;   // 8192 columns per CTA

__global__ void __launch_bounds__(FM_THREADS) vx_fold_matrix_kernel(const int32_t *__restrict__ dump,
                                                                    int64_t ld, int64_t c0, int64_t c1,
                                                                    const Win *__restrict__ wins,
                                                                    unsigned long long *acc, int R, int bins64)
{
    __shared__ unsigned long long sh[3][FM_THREADS / 32];
    const int row = blockIdx.x;
    const Win wl = wins[row * 2], wh = wins[row * 2 + 1];
    unsigned long long *hist = acc + ACC_HDR + 3 * (int64_t)R;
    const int32_t *src = dump + (int64_t)row * ld;
    unsigned long long sum = 0;
    uint32_t clo = 0, chi = 0;
    auto one = [&](const int v) {
        sum += (unsigned long long)(uint32_t)v;
        const uint32_t dl = (uint32_t)(v - wl.a), dh = (uint32_t)(v - wh.a);
        clo += dl >> 31; chi += dh >> 31;
        if (dl < wl.len) bin_add(hist, wl.off + (dl >> wl.shift), bins64);
        if (dh < wh.len) bin_add(hist, wh.off + (dh >> wh.shift), bins64);
    };
    const bool aligned = ((ld | c0) & 3) == 0;
    for (int64_t cb = c0 + (int64_t)blockIdx.y * FM_COLS; cb < c1; cb += (int64_t)gridDim.y * FM_COLS) {
        const int64_t ce = min(c1, cb + FM_COLS);
        if (aligned && ce - cb == FM_COLS) {
            const int4 *src4 = reinterpret_cast<const int4 *>(src + cb) + threadIdx.x;
            int4 v[FM_VEC];
#pragma unroll
            for (int k = 0; k < FM_VEC; ++k) v[k] = src4[k * FM_THREADS];
#pragma unroll
            for (int k = 0; k < FM_VEC; ++k) { one(v[k].x); one(v[k].y); one(v[k].z); one(v[k].w); }
        } else {
            for (int64_t c = cb + threadIdx.x; c < ce; c += FM_THREADS) one(src[c]);
        }
    }
    unsigned long long lo64 = clo, hi64 = chi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        lo64 += __shfl_xor_sync(0xffffffffu, lo64, o);
        hi64 += __shfl_xor_sync(0xffffffffu, hi64, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = sum; sh[1][wid] = lo64; sh[2][wid] = hi64; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t[3] = {0, 0, 0};
        for (int w = 0; w < FM_THREADS / 32; ++w) { t[0] += sh[0][w]; t[1] += sh[1][w]; t[2] += sh[2][w]; }
        atomicAdd(acc + ACC_HDR + row, t[0]);
        if (t[1]) atomicAdd(acc + ACC_HDR + R + 2 * row, t[1]);
        if (t[2]) atomicAdd(acc + ACC_HDR + R + 2 * row + 1, t[2]);
        if (row == 0 && blockIdx.y == 0) atomicAdd(acc + 0, (unsigned long long)(c1 - c0));
    }
}

// Expected-value stepper: np.random.poisson := identity, state carried in fp64 exactly as
// Python carries it (model.py:29-136 with lam substituted for the draw).  One thread per
// sample; rows [0,4T) are the four daily series in the reference's float units, rows
// [4T,4T+W) the weekly series daily_pm[7k] + daily_pm[7k-30] (model.py:196-214).
__global__ void __launch_bounds__(128) vx_expected_kernel(const SimArgs a, double *out,
                                                          int64_t ld)
{
    const int64_t local = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (local >= a.count) return;
    const uint64_t sample = (uint64_t)(a.first + local);
    SampleParams p;
    if (a.params) {
        const double *src = a.params + local * 6;
        p.p_pro = src[0]; p.p_anti = src[1]; p.pressure = src[2];
        p.tau = src[3]; p.nv0 = src[4]; p.nvmax = src[5];
    } else {
        draw_sample_params(a.bd.b, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), sample, p, 1u << 20);
    }
    const double Nd = (double)a.N;
    const double p_agn = 1 - (p.p_pro + p.p_anti);
    double n_pro = (double)(int64_t)(p.p_pro * Nd);
    double n_agn = (double)(int64_t)(p_agn * Nd);
    double stock = 0, received = 0, n_vacc = 0;
    double n_wait = n_pro - n_vacc;
    const double ln2 = 0.6931471805599453;
    const int T = a.T;
    double q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;
    for (int day = 0; day < T; ++day) {
        double arr = 0;
        const int dd = day % 7;
        if (dd == 0) {
            const double g = p.nv0 * exp(ln2 * (double)day / (p.tau * 7));
            const double m = (p.nvmax < g) ? p.nvmax : g;
            arr = (double)(int64_t)(m * Nd);
        }
        stock += arr; received += arr;
        const double avail = (2.0 / 7.0) * stock / Nd;
        double dv = n_wait * avail;
        if (stock < dv) dv = stock;
        if (n_wait < dv) dv = n_wait;
        n_vacc += dv; n_wait -= dv; stock -= dv;
        const double f = n_vacc / Nd;
        const double pc = f * p.pressure;
        double da = n_agn * pc;
        if (n_agn < da) da = n_agn;
        n_agn -= da; n_wait += da;
        const double dpm = dv * 1e6 / Nd;
        out[(int64_t)day * ld + local] = f * 100;
        out[((int64_t)T + day) * ld + local] = dpm;
        out[((int64_t)2 * T + day) * ld + local] = received * 100 / Nd;
        out[((int64_t)3 * T + day) * ld + local] = stock * 100 / Nd;
        if (dd == 0) {
            const double wkv = (day >= 30) ? dpm + q0 : dpm;
            out[((int64_t)4 * T + day / 7) * ld + local] = wkv;
        }
        if (dd == 5) { q0 = q1; q1 = q2; q2 = q3; q3 = q4; q4 = dpm; }
    }
}

// ------------------------------------------------------------- exact order statistics
// One CTA per row of a [rows][ld] matrix.  Finds the order statistics at up to 8 ranks by
// MSD radix select (8-bit digits, shared-memory histograms; ranks whose prefixes coincide
// share a histogram) and the row sum.  Exact for any input: this is np.quantile's
// partition step (model.py:220) done on integers / order-preserving keys.
template <typename T> struct KeyTraits;
template <> struct KeyTraits<int32_t> {
    typedef uint32_t key_t; typedef long long sum_t;
    static constexpr int BITS = 32;
    __device__ static key_t to_key(int32_t v) { return (uint32_t)v ^ 0x80000000u; }
    __device__ static int32_t from_key(key_t k) { return (int32_t)(k ^ 0x80000000u); }
    __device__ static sum_t to_sum(int32_t v) { return (long long)v; }
};
template <> struct KeyTraits<double> {
    typedef unsigned long long key_t; typedef double sum_t;
    static constexpr int BITS = 64;
    __device__ static key_t to_key(double v) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    __device__ static double from_key(key_t k) {
        const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        return __longlong_as_double((long long)b);
    }
    __device__ static sum_t to_sum(double v) { return v; }
};

constexpr int SEL_THREADS = 512;
constexpr int SEL_MAXR = 8;

template <typename T> struct SelScratch {
    typedef typename KeyTraits<T>::key_t key_t;
    uint32_t hist[SEL_MAXR][256];
    key_t prefix[SEL_MAXR], dprefix[SEL_MAXR];
    long long rem[SEL_MAXR];
    int hid[SEL_MAXR];
    int ndist;
};

// MSD radix select of `nranks` order statistics of src[0..n) (global or shared memory) by the
// whole CTA.  In: S.rem[k] = rank k; first_key / diff = any key of the data and the OR of
// (key ^ first_key) over the data (bits that vary).  Out: S.prefix[k] = key of the order
// statistic.  Ends with a __syncthreads().
template <typename T>
__device__ void block_radix_select(const T *__restrict__ src, int64_t n, int nranks, SelScratch<T> &S,
                                   typename KeyTraits<T>::key_t first_key,
                                   typename KeyTraits<T>::key_t diff)
{
    typedef KeyTraits<T> KT;
    typedef typename KT::key_t key_t;
    const int tid = threadIdx.x;
    // keys agree above bit `top`: those bits go straight into the prefix
    int top = 0;
    if (diff) top = (KT::BITS == 64) ? 64 - __clzll((long long)diff) : 32 - __clz((int)(uint32_t)diff);
    const int npass = (top + 7) / 8;
    if (tid < nranks) S.prefix[tid] = (npass * 8 >= KT::BITS) ? (key_t)0 : (first_key >> (npass * 8));
    __syncthreads();

    for (int pass = npass - 1; pass >= 0; --pass) {
        const int shift = pass * 8;
        if (tid == 0) {
            int nd = 0;
            for (int k = 0; k < nranks; ++k) {
                int f = -1;
                for (int d = 0; d < nd; ++d) if (S.dprefix[d] == S.prefix[k]) { f = d; break; }
                if (f < 0) { f = nd; S.dprefix[nd++] = S.prefix[k]; }
                S.hid[k] = f;
            }
            S.ndist = nd;
        }
        for (int i = tid; i < SEL_MAXR * 256; i += SEL_THREADS) (&S.hist[0][0])[i] = 0;
        __syncthreads();
        const int nd = S.ndist;
        const bool whole = (shift + 8 >= KT::BITS);
        for (int64_t i0 = 0; i0 < n; i0 += SEL_THREADS) {
            const int64_t i = i0 + tid;
            const bool in = i < n;
            key_t key = 0, hi = 0;
            uint32_t bin = 0;
            if (in) {
                key = KT::to_key(src[i]);
                hi = whole ? (key_t)0 : (key >> (shift + 8));
                bin = (uint32_t)(key >> shift) & 255u;
            }
            for (int d = 0; d < nd; ++d) {
                const bool hit = in && (hi == S.dprefix[d]);
                // warp-aggregate the common all-lanes-same-bin case (flat rows)
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (m == 0) continue;
                const int leader = __ffs(m) - 1;
                const uint32_t lbin = __shfl_sync(0xffffffffu, bin, leader);
                const uint32_t same = __ballot_sync(0xffffffffu, hit && bin == lbin);
                if (same == m) {
                    if ((tid & 31) == leader) atomicAdd(&S.hist[d][lbin], (uint32_t)__popc(m));
                } else if (hit) {
                    atomicAdd(&S.hist[d][bin], 1u);
                }
            }
        }
        __syncthreads();
        if ((tid >> 5) < nranks) {           // warp w locates rank w in its 256-bin histogram
            const int w = tid >> 5, l = tid & 31;
            const uint32_t *h = S.hist[S.hid[w]] + 8 * l;
            uint32_t c[8], lsum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = h[i]; lsum += c[i]; }
            long long incl = lsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, incl, o);
                if (l >= o) incl += y;
            }
            const long long excl = incl - lsum, r = S.rem[w];
            const uint32_t m = __ballot_sync(0xffffffffu, r >= excl && r < incl);
            const int src_lane = m ? __ffs(m) - 1 : 31;
            if (l == src_lane) {
                long long cum = excl;
                int i = 0;
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    if (i == q && r >= cum + (long long)c[q]) { cum += c[q]; i = q + 1; }
                }
                S.rem[w] = r - cum;
                S.prefix[w] = (S.prefix[w] << 8) | (key_t)(8 * l + i);
            }
        }
        __syncthreads();
    }
}

// Row sum (fixed-order tree -> deterministic) and the varying key bits of a row; all threads
// return the same (sum in red_sum[0], diff).
template <typename T>
__device__ typename KeyTraits<T>::key_t block_sum_and_diff(const T *__restrict__ row, int64_t n,
                                                           typename KeyTraits<T>::sum_t *red_sum,
                                                           typename KeyTraits<T>::key_t *red_xor,
                                                           typename KeyTraits<T>::key_t first_key)
{
    typedef KeyTraits<T> KT;
    const int tid = threadIdx.x;
    typename KT::sum_t acc = 0;
    typename KT::key_t diff = 0;
    for (int64_t i = tid; i < n; i += SEL_THREADS) {
        const T v = row[i];
        acc += KT::to_sum(v);
        diff |= KT::to_key(v) ^ first_key;
    }
    red_sum[tid] = acc; red_xor[tid] = diff;
    __syncthreads();
    for (int off = SEL_THREADS / 2; off > 0; off >>= 1) {
        if (tid < off) { red_sum[tid] += red_sum[tid + off]; red_xor[tid] |= red_xor[tid + off]; }
        __syncthreads();
    }
    return red_xor[0];
}

template <typename T>
__global__ void __launch_bounds__(SEL_THREADS) vx_select_kernel(
    const T *__restrict__ data, int64_t ld, int64_t n, const int64_t *__restrict__ ranks,
    int nranks, T *__restrict__ out /*[rows][nranks]*/,
    typename KeyTraits<T>::sum_t *__restrict__ sums /*[rows]*/)
{
    typedef KeyTraits<T> KT;
    typedef typename KT::key_t key_t;
    typedef typename KT::sum_t sum_t;
    __shared__ SelScratch<T> S;
    __shared__ sum_t red_sum[SEL_THREADS];
    __shared__ key_t red_xor[SEL_THREADS];
    const T *row = data + (int64_t)blockIdx.x * ld;
    const int tid = threadIdx.x;
    const key_t first_key = KT::to_key(row[0]);
    const key_t diff = block_sum_and_diff<T>(row, n, red_sum, red_xor, first_key);
    if (tid == 0 && sums) sums[blockIdx.x] = red_sum[0];
    if (tid < nranks) S.rem[tid] = ranks[tid];
    __syncthreads();
    block_radix_select<T>(row, n, nranks, S, first_key, diff);
    if (tid < nranks) out[(int64_t)blockIdx.x * nranks + tid] = KT::from_key(S.prefix[tid]);
}

// Order statistics of int32 rows by histogram refinement (the default for int32; the radix kernel
// above stays for fp64 rows and as a cross-check).  One CTA per row:
//   pass 0   row sum, min and max (the only pass that looks at every bit of every value)
//   level 0  SELS_BINS bins over [min, max]
//   level k  all still-open ranks at once, SELS_BINS / #ranks bins each over their bins: every
//            level divides the open ranges by >= 256, so four levels resolve any int32 row.
// STAGED: the row fits in shared memory (the pilot matrix: 16384 columns = 64 KB per row): it is
//         read from HBM/L2 ONCE and the levels run on chip.
// else:   the row is streamed from global memory once per pass with 16-byte loads (1 + levels
//         reads, typically 3, ~6 instructions per element and pass; the 8-bit radix select needs
//         1 + 4 reads at ~25 instructions per element and pass).
constexpr int SELS_THREADS = 512;
// 2048 bins (8 KB) next to a 64 KB pilot row = three CTAs per SM (40 registers): measured 0.116 ms
// for the 1290 x 16384 pilot select against 0.136 ms with 4096 bins and two CTAs per SM
constexpr int SELS_BINS = 2048;
constexpr int SELS_CTAS = 3;

template <bool STAGED>
__global__ void __launch_bounds__(SELS_THREADS, SELS_CTAS) vx_select_hist_kernel(
    const int32_t *__restrict__ data, int64_t ld, int64_t n, const int64_t *__restrict__ ranks, int nranks,
    int32_t *__restrict__ out /*[rows][nranks]*/, long long *__restrict__ sums /*[rows]*/)
{
    extern __shared__ __align__(16) int sels_sm[];
    int *vals = sels_sm;                                            // STAGED: [n rounded up to 4]
    uint32_t *hist = reinterpret_cast<uint32_t *>(sels_sm + (STAGED ? (int)((n + 3) & ~(int64_t)3) : 0));   // [SELS_BINS]
    __shared__ long long red_sum[SELS_THREADS / 32];
    __shared__ int red_min[SELS_THREADS / 32], red_max[SELS_THREADS / 32];
    __shared__ uint32_t warp_tot[SELS_THREADS / 32];
    __shared__ long long t_lo[SEL_MAXR], t_w[SEL_MAXR], t_rem[SEL_MAXR];   // open interval [lo, lo+w) and rank inside it
    __shared__ int t_shift[SEL_MAXR], t_slot[SEL_MAXR], t_open[SEL_MAXR];
    __shared__ long long s_lo[SEL_MAXR], s_w[SEL_MAXR];             // distinct intervals of this level
    __shared__ int s_shift[SEL_MAXR], s_n;
    __shared__ int s_small;                                         // row range < 2^31: 32-bit interval tests are exact
    __shared__ long long u_lo[SEL_MAXR], u_w[SEL_MAXR], u_rem[SEL_MAXR];   // next level's intervals (written by the bin owners)
    __shared__ uint32_t slot_base[SEL_MAXR];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t *row = data + (int64_t)blockIdx.x * ld;
    const bool vec = ((ld | n) & 3) == 0;
    const int64_t n4 = n / 4;

    // Every value of the row, four per thread and step where the layout allows, with a uniform trip
    // count (f may use warp collectives); `ok` is false for the padding of the last step.
    auto for_each4 = [&](auto &&f) {
        if (vec) {
            const int4 *src = STAGED ? reinterpret_cast<const int4 *>(vals) : reinterpret_cast<const int4 *>(row);
            int64_t i0 = 0;
            if (!STAGED)       // two 16-byte loads in flight per thread
                for (; i0 + 2 * SELS_THREADS <= n4; i0 += 2 * SELS_THREADS) {
                    const int4 a = src[i0 + tid], b = src[i0 + SELS_THREADS + tid];
                    f(a, true); f(b, true);
                }
            for (; i0 < n4; i0 += SELS_THREADS) {
                const bool ok = i0 + tid < n4;
                const int4 a = ok ? src[i0 + tid] : make_int4(0, 0, 0, 0);
                f(a, ok);
            }
        } else {
            const int32_t *src = STAGED ? vals : row;
            for (int64_t i0 = 0; i0 < n; i0 += SELS_THREADS) {
                const bool ok = i0 + tid < n;
                const int v = ok ? src[i0 + tid] : 0;
                f(make_int4(v, v, v, v), ok, 1);
            }
        }
    };

    // ---- pass 0: (stage the row;) sum / min / max
    long long acc = 0;
    int mn = INT_MAX, mx = INT_MIN;
    if (vec) {
        const int4 *src = reinterpret_cast<const int4 *>(row);
        int64_t i = tid;
        if (!STAGED)
            for (; i + SELS_THREADS < n4; i += 2 * SELS_THREADS) {
                const int4 v = src[i], w = src[i + SELS_THREADS];
                acc += (long long)v.x + v.y + v.z + v.w + w.x + w.y + w.z + w.w;
                mn = min(min(min(mn, v.x), min(v.y, min(v.z, v.w))), min(min(w.x, w.y), min(w.z, w.w)));
                mx = max(max(max(mx, v.x), max(v.y, max(v.z, v.w))), max(max(w.x, w.y), max(w.z, w.w)));
            }
        for (; i < n4; i += SELS_THREADS) {
            const int4 v = src[i];
            if (STAGED) reinterpret_cast<int4 *>(vals)[i] = v;
            acc += (long long)v.x + v.y + v.z + v.w;
            mn = min(min(mn, v.x), min(v.y, min(v.z, v.w)));
            mx = max(max(mx, v.x), max(v.y, max(v.z, v.w)));
        }
    } else {
        for (int64_t i = tid; i < n; i += SELS_THREADS) {
            const int v = row[i];
            if (STAGED) vals[i] = v;
            acc += v; mn = min(mn, v); mx = max(mx, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red_sum[wid] = acc; red_min[wid] = mn; red_max[wid] = mx; }
    __syncthreads();
    if (tid == 0) {
        long long a = 0; int lo = INT_MAX, hi = INT_MIN;
        for (int w = 0; w < SELS_THREADS / 32; ++w) { a += red_sum[w]; lo = min(lo, red_min[w]); hi = max(hi, red_max[w]); }
        if (sums) sums[blockIdx.x] = a;
        s_small = ((long long)hi - lo) < (1ll << 31);
        for (int k = 0; k < nranks; ++k) {
            t_lo[k] = lo; t_w[k] = (long long)hi - lo + 1; t_rem[k] = ranks[k]; t_open[k] = t_w[k] > 1;
        }
    }
    __syncthreads();

    for (int level = 0; level < 8; ++level) {
        // ---- distinct open intervals of this level, SELS_BINS / (their number, rounded up to 2^k) bins each
        if (tid == 0) {
            int nd = 0;
            for (int k = 0; k < nranks; ++k) {
                if (!t_open[k]) { t_slot[k] = -1; continue; }
                int f = -1;
                for (int d = 0; d < nd; ++d) if (s_lo[d] == t_lo[k]) { f = d; break; }
                if (f < 0) { f = nd; s_lo[nd] = t_lo[k]; s_w[nd] = t_w[k]; ++nd; }
                t_slot[k] = f;
            }
            s_n = nd;
            int per = SELS_BINS;
            while (per * nd > SELS_BINS) per >>= 1;            // nd <= 8 -> per >= 256
            for (int d = 0; d < nd; ++d) {
                int sh = 0;
                while (((s_w[d] + ((1ll << sh) - 1)) >> sh) > per) ++sh;
                s_shift[d] = sh;
            }
            for (int k = 0; k < nranks; ++k) if (t_slot[k] >= 0) t_shift[k] = s_shift[t_slot[k]];
        }
        __syncthreads();
        const int nd = s_n;
        if (nd == 0) break;
        int per = SELS_BINS;
        while (per * nd > SELS_BINS) per >>= 1;
        for (int i = tid; i < SELS_BINS; i += SELS_THREADS) hist[i] = 0;
        __syncthreads();
        // ---- histogram of the values inside the open intervals
        // The common shapes take a lean path: one or two open intervals of a row whose range fits 31
        // bits, tested in 32-bit wrap-around arithmetic ((uint32)(v - lo) < w is exact then), no warp
        // collectives -- ~5 instructions per value instead of ~20.  A single interval of at most 64
        // values (a flat or nearly flat row, where same-bin atomics would serialise) and everything
        // else go through the general code below.
        if (nd == 1 && s_small && s_w[0] > 64) {
            const uint32_t lo = (uint32_t)(int)s_lo[0], w = (uint32_t)s_w[0];
            const int sh = s_shift[0];
            auto one = [&](const int v) {
                const uint32_t d = (uint32_t)v - lo;
                if (d < w) atomicAdd(&hist[d >> sh], 1u);
            };
            for_each4([&](const int4 v, const bool ok, const int cnt = 4) {
                if (!ok) return;
                one(v.x);
                if (cnt == 4) { one(v.y); one(v.z); one(v.w); }
            });
        } else if (nd == 2 && s_small) {
            const uint32_t lo0 = (uint32_t)(int)s_lo[0], w0 = (uint32_t)s_w[0], lo1 = (uint32_t)(int)s_lo[1], w1 = (uint32_t)s_w[1];
            const int sh0 = s_shift[0], sh1 = s_shift[1];
            uint32_t *hist1 = hist + per;
            auto one = [&](const int v) {
                const uint32_t d0 = (uint32_t)v - lo0, d1 = (uint32_t)v - lo1;
                if (d0 < w0) atomicAdd(&hist[d0 >> sh0], 1u);
                else if (d1 < w1) atomicAdd(&hist1[d1 >> sh1], 1u);
            };
            for_each4([&](const int4 v, const bool ok, const int cnt = 4) {
                if (!ok) return;
                one(v.x);
                if (cnt == 4) { one(v.y); one(v.z); one(v.w); }
            });
        } else if (nd == 1) {
            const long long lo = s_lo[0], w = s_w[0];
            const int sh = s_shift[0];
            auto one = [&](const int v, const bool ok) {
                const long long d = ok ? (long long)v - lo : -1;
                const bool hit = ok && d >= 0 && d < w;
                const uint32_t b = hit ? (uint32_t)(d >> sh) : 0xffffffffu;
                // flat rows: the whole warp in one bin -> one atomic instead of 32 serialised ones
                const uint32_t b0 = __shfl_sync(0xffffffffu, b, 0);
                if (__all_sync(0xffffffffu, b == b0)) {
                    if (lane == 0 && b0 != 0xffffffffu) atomicAdd(&hist[b0], 32u);
                } else if (hit) {
                    atomicAdd(&hist[b], 1u);
                }
            };
            for_each4([&](const int4 v, const bool ok, const int cnt = 4) {
                one(v.x, ok);
                if (cnt == 4) { one(v.y, ok); one(v.z, ok); one(v.w, ok); }
            });
        } else {
            auto one = [&](const int vi) {
                const long long v = vi;
                for (int d = 0; d < nd; ++d) {
                    const long long dd = v - s_lo[d];
                    if (dd >= 0 && dd < s_w[d]) { atomicAdd(&hist[d * per + (int)(dd >> s_shift[d])], 1u); break; }
                }
            };
            for_each4([&](const int4 v, const bool ok, const int cnt = 4) {
                if (!ok) return;
                one(v.x);
                if (cnt == 4) { one(v.y); one(v.z); one(v.w); }
            });
        }
        __syncthreads();
        // ---- exclusive scan of the SELS_BINS counts (8 per thread), restarted at every interval's first bin
        uint32_t c[SELS_BINS / SELS_THREADS], tsum = 0;
#pragma unroll
        for (int q = 0; q < SELS_BINS / SELS_THREADS; ++q) { c[q] = hist[tid * (SELS_BINS / SELS_THREADS) + q]; tsum += c[q]; }
        uint32_t incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < wid; ++w) base += warp_tot[w];
        const uint32_t excl_global = base + incl - tsum;        // values in bins before this thread's first bin
        // the interval that owns this thread's bins starts at bin slot*per: its own offset = the count before it
        const int my_slot = (tid * (SELS_BINS / SELS_THREADS)) / per;
        if ((tid * (SELS_BINS / SELS_THREADS)) % per == 0 && my_slot < nd) slot_base[my_slot] = excl_global;
        __syncthreads();
        if (my_slot < nd && tsum > 0) {
            uint32_t run = excl_global - slot_base[my_slot];    // count inside the interval before this thread's bins
            for (int k = 0; k < nranks; ++k) {
                if (t_slot[k] != my_slot) continue;
                const long long r = t_rem[k];
                if (r < (long long)run || r >= (long long)run + tsum) continue;     // not in this thread's bins
                uint32_t before = run;
#pragma unroll
                for (int q = 0; q < SELS_BINS / SELS_THREADS; ++q) {
                    if (r >= (long long)before && r < (long long)before + c[q]) {
                        // exactly one (thread, q) matches rank k; the t_* arrays are still being read by the
                        // other threads, so the new interval goes to u_* and is adopted after the barrier
                        const int bin = tid * (SELS_BINS / SELS_THREADS) + q - my_slot * per;
                        const int sh = t_shift[k];
                        const long long nlo = t_lo[k] + ((long long)bin << sh);
                        u_lo[k] = nlo; u_w[k] = min(1ll << sh, t_lo[k] + t_w[k] - nlo); u_rem[k] = r - before;
                    }
                    before += c[q];
                }
            }
        }
        __syncthreads();
        if (tid < nranks && t_slot[tid] >= 0) {
            t_lo[tid] = u_lo[tid]; t_w[tid] = u_w[tid]; t_rem[tid] = u_rem[tid]; t_open[tid] = u_w[tid] > 1;
        }
        __syncthreads();
    }
    if (tid < nranks) out[(int64_t)blockIdx.x * nranks + tid] = (int32_t)t_lo[tid];
}

// --------------------------------------------- order statistics out of merged windows
// np.quantile(method="linear"): virtual index (n-1)*q -> the two neighbouring ranks
// (numpy/lib/_function_base_impl.py: _get_indexes).  Same arithmetic on host and device.
__host__ __device__ inline void quantile_ranks(long long n, double q, long long &prev, long long &next)
{
    const double vi = (double)(n - 1) * q;
    const double p = floor(vi);
    prev = (long long)p; next = prev + 1;
    if (vi >= (double)(n - 1)) { prev = n - 1; next = n - 1; }
    if (vi < 0) { prev = 0; next = 0; }
}

// One warp per (row, window).  kind[(row*2+w)*2 + {0,1}] says which of the row's four wanted
// order statistics the window serves (0: lower band rank j, 1: j+1, 2: upper band rank j, 3: j+1;
// -1: none); the ranks themselves follow from n = acc[0] (samples folded by all ranks) on the
// device.  res gets, per target, (status << 32) | bin with status 0 found, 1 the rank lies
// below the window (rank < #values < a), 2 above it.
__global__ void __launch_bounds__(256) vx_extract_kernel(const unsigned long long *__restrict__ acc,
                                                         const Win *__restrict__ wins,
                                                         const int *__restrict__ kind,
                                                         long long *__restrict__ res, int R,
                                                         double q_lo, double q_hi, int bins64)
{
    const int gw = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= 2 * R) return;
    const Win w = wins[gw];
    const uint32_t nb = w.len >> w.shift;
    const unsigned long long *hist64 = acc + ACC_HDR + 3 * (int64_t)R;
    const unsigned int *hist32 = reinterpret_cast<const unsigned int *>(hist64) + w.off;
    hist64 += w.off;
    auto bin = [&](uint32_t b) -> long long { return bins64 ? (long long)hist64[b] : (long long)hist32[b]; };
    const long long below = (long long)acc[ACC_HDR + R + gw];
    long long rk[4];
    quantile_ranks((long long)acc[0], q_lo, rk[0], rk[1]);
    quantile_ranks((long long)acc[0], q_hi, rk[2], rk[3]);
    long long t[2], r[2] = {-1, -1};
    for (int k = 0; k < 2; ++k) {
        const int kd = kind[gw * 2 + k];
        t[k] = kd < 0 ? -1 : (kd == 0 ? rk[0] : kd == 1 ? rk[1] : kd == 2 ? rk[2] : rk[3]);
        if (t[k] >= 0 && t[k] < below) { r[k] = (1ll << 32); t[k] = -1; }
    }
    long long cum = below;
    for (uint32_t b0 = 0; b0 < nb && (t[0] >= 0 || t[1] >= 0); b0 += 128) {
        const uint32_t b = b0 + 4 * lane;                    // 4 bins per lane: 4 loads in flight
        const long long h0 = (b < nb) ? bin(b) : 0;
        const long long h1 = (b + 1 < nb) ? bin(b + 1) : 0;
        const long long h2 = (b + 2 < nb) ? bin(b + 2) : 0;
        const long long h3 = (b + 3 < nb) ? bin(b + 3) : 0;
        const long long lsum = h0 + h1 + h2 + h3;
        long long incl = lsum;                                // inclusive warp scan
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const long long c_excl = cum + incl - lsum, c_incl = cum + incl;
        for (int k = 0; k < 2; ++k) {
            if (t[k] < 0) continue;
            const uint32_t m = __ballot_sync(0xffffffffu, t[k] < c_incl);
            if (m) {
                long long c = c_excl;
                int i = 3;
                if (t[k] < c + h0) i = 0;
                else if (t[k] < c + h0 + h1) i = 1;
                else if (t[k] < c + h0 + h1 + h2) i = 2;
                r[k] = (long long)__shfl_sync(0xffffffffu, (int)(b + i), __ffs(m) - 1);
                t[k] = -1;
            }
        }
        cum = __shfl_sync(0xffffffffu, c_incl, 31);
    }
    for (int k = 0; k < 2; ++k)
        if (t[k] >= 0) r[k] = (2ll << 32);
    if (lane == 0) { res[gw * 2] = r[0]; res[gw * 2 + 1] = r[1]; }
}

// ------------------------------------------------------------------- self-test kernels
__global__ void vx_test_philox_kernel(uint4 ctr, uint2 key, uint32_t *out)
{
    const uint4 r = philox4x32_10(ctr.x, ctr.y, ctr.z, ctr.w, key.x, key.y);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

__global__ void vx_test_poisson_kernel(float lam, int64_t n, uint64_t seed, int32_t *out, int exact)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, TAG_TEST, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
    float z1, z2;
    normal_pair(r.x, r.y, z1, z2);
    if (exact) {
        const PhiloxKeys K = philox_keys(seed);
        out[i] = poisson_draw_exact(lam, (uint32_t)i, (uint32_t)(i >> 32), 3u, 0u, K);
        return;
    }
    // the stepper's own code path: packed coefficients for the pair, first / second draw by parity
    const CFCoef cf = cf_coefficients(z1, z2);
    out[i] = (i & 1) ? poisson_draw_c<1>(lam, z2, cf) : poisson_draw_c<0>(lam, z1, cf);
}

__global__ void vx_test_normal_kernel(int64_t n, uint64_t seed, float *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i >= n) return;
    const uint4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 1u, TAG_TEST, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
    float z1, z2;
    normal_pair(r.x, r.y, z1, z2);
    out[2 * i] = z1;
    if (2 * i + 1 < n) out[2 * i + 1] = z2;
}

} // namespace vx
