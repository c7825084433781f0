/*
 * oracle/vaxmodel_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the Monte Carlo hot path of the reference
 * /root/reference/model.py, used only as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product path (covid19-vaccination-model_b200/) never links, imports or
 * calls anything in this file.
 *
 * What is restated, with the reference line each piece follows:
 *   vo_run_single_realization   model.py:29-136  (init 63-80, arrivals 69/85-92,
 *                               vaccination draw 102-112, conversion draw 115-120,
 *                               recording 124-127)
 *   vo_sample_param_combinations model.py:231-303 (draw order, rejection, abort)
 *   vo_weekly_series            model.py:196-214 (30-day shift + 7-day subsample)
 *
 * Third-party arithmetic the reference calls (NOT under /root/reference; numpy is
 * unpinned in requirements.txt:4, numpy 2.3.5 in this image) is restated from
 * numpy's published algorithms so that the oracle reproduces the reference's
 * random stream bit for bit:
 *   MT19937 (init_genrand seeding used by np.random.seed(int), genrand_int32,
 *   53-bit doubles from two 32-bit words), legacy RandomState.uniform
 *   (lo + (hi-lo)*double), legacy RandomState.poisson (lam==0 -> 0, lam<10 ->
 *   multiplication method, lam>=10 -> Hoermann PTRS with numpy's loggam).
 *
 * PARITY PIN: the reference ships no tests or golden vectors (Makefile:5 points
 * at a missing ./test).  The oracle is therefore pinned against outputs of the
 * reference itself, generated in the build container by oracle/make_golden.py
 * (imports /root/reference/model.py unmodified with a stub `plot` module) and
 * committed under tests/golden/.  tests/test_oracle_pin.py checks this file
 * against those fixtures bit-exactly (seeded stochastic runs, np.random.seed)
 * and in the deterministic expected-value mode (np.random.poisson := identity).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ MT19937 */
#define MT_N 624
#define MT_M 397

typedef struct {
    uint32_t key[MT_N];
    int pos;
} vo_mt;

/* numpy `_legacy_seeding(int)` -> mt19937_seed == init_genrand */
void vo_mt_seed(vo_mt *s, uint32_t seed)
{
    seed &= 0xffffffffu;
    for (int i = 0; i < MT_N; i++) {
        s->key[i] = seed;
        seed = (1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u) & 0xffffffffu;
    }
    s->pos = MT_N;
}

static void mt_gen(vo_mt *s)
{
    uint32_t y;
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
        y = (s->key[kk] & 0x80000000u) | (s->key[kk + 1] & 0x7fffffffu);
        s->key[kk] = s->key[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < MT_N - 1; kk++) {
        y = (s->key[kk] & 0x80000000u) | (s->key[kk + 1] & 0x7fffffffu);
        s->key[kk] = s->key[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (s->key[MT_N - 1] & 0x80000000u) | (s->key[0] & 0x7fffffffu);
    s->key[MT_N - 1] = s->key[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->pos = 0;
}

static inline uint32_t mt_next32(vo_mt *s)
{
    uint32_t y;
    if (s->pos == MT_N) mt_gen(s);
    y = s->key[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

double vo_mt_double(vo_mt *s)
{
    int32_t a = (int32_t)(mt_next32(s) >> 5), b = (int32_t)(mt_next32(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* legacy RandomState.uniform(low, high): low + (high-low)*U  (model.py:277-294) */
double vo_uniform(vo_mt *s, double lo, double hi)
{
    double range = hi - lo;
    return lo + range * vo_mt_double(s);
}

/* ------------------------------------------------- numpy legacy poisson law */
static double np_loggam(double x)
{
    static const double a[10] = {
        8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04,
        -5.952380952380952e-04, 8.417508417508418e-04, -1.917526917526918e-03,
        6.410256410256410e-03, -2.955065359477124e-02, 1.796443723688307e-01,
        -1.39243221690590e+00};
    double x0, x2, lg2pi, gl, gl0;
    int64_t k, n;

    if ((x == 1.0) || (x == 2.0)) return 0.0;
    n = (x < 7.0) ? (int64_t)(7 - x) : 0;
    x0 = x + n;
    x2 = (1.0 / x0) * (1.0 / x0);
    lg2pi = 1.8378770664093453e+00;
    gl0 = a[9];
    for (k = 8; k >= 0; k--) {
        gl0 *= x2;
        gl0 += a[k];
    }
    gl = gl0 / x0 + 0.5 * lg2pi + (x0 - 0.5) * log(x0) - x0;
    if (x < 7.0) {
        for (k = 1; k <= n; k++) {
            gl -= log(x0 - 1.0);
            x0 -= 1.0;
        }
    }
    return gl;
}

static int64_t np_poisson_mult(vo_mt *s, double lam)
{
    int64_t X = 0;
    double prod = 1.0, enlam = exp(-lam);
    for (;;) {
        prod *= vo_mt_double(s);
        if (prod > enlam)
            X += 1;
        else
            return X;
    }
}

static int64_t np_poisson_ptrs(vo_mt *s, double lam)
{
    int64_t k;
    double U, V, slam, loglam, a, b, invalpha, vr, us;

    slam = sqrt(lam);
    loglam = log(lam);
    b = 0.931 + 2.53 * slam;
    a = -0.059 + 0.02483 * b;
    invalpha = 1.1239 + 1.1328 / (b - 3.4);
    vr = 0.9277 - 3.6224 / (b - 2);

    for (;;) {
        U = vo_mt_double(s) - 0.5;
        V = vo_mt_double(s);
        us = 0.5 - fabs(U);
        k = (int64_t)floor((2 * a / us + b) * U + lam + 0.43);
        if ((us >= 0.07) && (V <= vr)) return k;
        if ((k < 0) || ((us < 0.013) && (V > us))) continue;
        if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <=
            (-lam + k * loglam - np_loggam((double)k + 1)))
            return k;
    }
}

/* np.random.poisson(lam) for scalar lam >= 0 (model.py:104, 116) */
int64_t vo_poisson(vo_mt *s, double lam)
{
    if (lam >= 10)
        return np_poisson_ptrs(s, lam);
    else if (lam == 0)
        return 0;
    else
        return np_poisson_mult(s, lam);
}

/* ---------------------------------------------- A1: parameter combinations */
/*
 * model.py:231-303.  bounds = {pro_lo,pro_hi, anti_lo,anti_hi, pressure_lo,hi,
 * tau_lo,hi, nv0_lo,hi, nvmax_lo,hi}, already in model units (fractions; the /100
 * of model.py:329-334 is the caller's job).  params_out is [n_rep][6] row-major in
 * the reference's tuple order; p_soft_no_out is [n_rep].
 * Returns the number of rejections, or -1 when the reference returns (None, None)
 * (model.py:283-287).
 */
int64_t vo_sample_param_combinations(vo_mt *s, const double *bounds, int64_t n_rep,
                                     double *params_out, double *p_soft_no_out)
{
    int64_t n = 0, have = 0;
    while (have < n_rep) {
        double p_pro = vo_uniform(s, bounds[0], bounds[1]);
        double p_anti = vo_uniform(s, bounds[2], bounds[3]);
        if (p_pro + p_anti > 1.0) {
            n += 1;
            if (n > n_rep * 10) return -1;
            continue;
        }
        double *p = params_out + 6 * have;
        p[0] = p_pro;
        p[1] = p_anti;
        p[2] = vo_uniform(s, bounds[4], bounds[5]);
        p[3] = vo_uniform(s, bounds[6], bounds[7]);
        p[4] = vo_uniform(s, bounds[8], bounds[9]);
        p[5] = vo_uniform(s, bounds[10], bounds[11]);
        p_soft_no_out[have] = 1 - (p_pro + p_anti);
        have++;
    }
    return n;
}

/* ------------------------------------------- A2-A6: one realization, floats */
/*
 * model.py:29-136.  mode 0: stochastic (np.random.poisson from state s);
 * mode 1: expected-value (np.random.poisson := lambda lam: lam), state carried in
 * doubles exactly as Python carries it (ints are exactly representable).
 * out is [4][T] row-major, rows in the reference's dict order:
 *   0 people_vaccinated_per_hundred   1 daily_vaccinations_per_million
 *   2 cum_number_vac_received_per_hundred   3 vaccines_in_stock_per_hundred
 * counts (optional, may be NULL; mode 0 only) receives the raw integers
 * [4][T]: n_vaccinated, delta_n_vacc, cum_number_vac_received, vaccines_stock.
 * Returns 0, or -1 if the reference's assert at model.py:63 would fire.
 */
int vo_run_single_realization(vo_mt *s, int mode, const double *p, int64_t max_day_number,
                              int64_t N, double *out, int32_t *counts)
{
    const double p_pro = p[0], p_anti = p[1], pressure = p[2], tau = p[3], nv_0 = p[4],
                 nv_max = p[5];
    const int64_t T = max_day_number;
    if (!(p_pro + p_anti <= 1.0)) return -1;
    const double p_agnostics = 1 - (p_pro + p_anti);
    const double Nd = (double)N;

    /* int() truncates toward zero (model.py:66-67) */
    double n_pro = (double)(int64_t)(p_pro * Nd);
    double n_agnostics = (double)(int64_t)(p_agnostics * Nd);

    double vaccines_stock = 0, cum_number_vac_received = 0, n_vaccinated = 0;
    double n_waiting = n_pro - n_vaccinated;
    const double ln2 = log(2.0);

    for (int64_t day = 0; day < T; day++) {
        double nv_arriving = 0;
        if (day % 7 == 0) {
            /* F(t) = min(nv_0*exp(log(2)*t/(tau*7)), nv_max)*N   (model.py:69) */
            double g = nv_0 * exp(ln2 * (double)day / (tau * 7));
            double m = (nv_max < g) ? nv_max : g; /* Python min(g, nv_max) */
            nv_arriving = (double)(int64_t)(m * Nd);
        }
        vaccines_stock += nv_arriving;
        cum_number_vac_received += nv_arriving;

        double proc_vac_available = (2.0 / 7.0) * vaccines_stock / Nd;
        double lam1 = n_waiting * proc_vac_available;
        double delta_n_vacc = mode ? lam1 : (double)vo_poisson(s, lam1);
        if (vaccines_stock < delta_n_vacc) delta_n_vacc = vaccines_stock;
        if (n_waiting < delta_n_vacc) delta_n_vacc = n_waiting;
        n_vaccinated += delta_n_vacc;
        n_waiting -= delta_n_vacc;
        vaccines_stock -= delta_n_vacc;
        double fract_pop_vaccinated = n_vaccinated / Nd;

        double prob_change_mind = fract_pop_vaccinated * pressure;
        double lam2 = n_agnostics * prob_change_mind;
        double delta_n_agnos = mode ? lam2 : (double)vo_poisson(s, lam2);
        if (n_agnostics < delta_n_agnos) delta_n_agnos = n_agnostics;
        n_agnostics -= delta_n_agnos;
        n_waiting += delta_n_agnos;

        out[0 * T + day] = fract_pop_vaccinated * 100;
        out[1 * T + day] = delta_n_vacc * 1e6 / Nd;
        out[2 * T + day] = cum_number_vac_received * 100 / Nd;
        out[3 * T + day] = vaccines_stock * 100 / Nd;
        if (counts) {
            counts[0 * T + day] = (int32_t)n_vaccinated;
            counts[1 * T + day] = (int32_t)delta_n_vacc;
            counts[2 * T + day] = (int32_t)cum_number_vac_received;
            counts[3 * T + day] = (int32_t)vaccines_stock;
        }
    }
    return 0;
}

/* A8: weekly series of one sample, model.py:196-214.
 * daily is the daily_vaccinations_per_million row [T]; weekly receives
 * W = ceil(T/7) values: daily[7k] + (daily[7k-30] if 7k >= 30 else 0). */
void vo_weekly_series(const double *daily, int64_t T, double *weekly)
{
    int64_t W = (T + 6) / 7;
    for (int64_t k = 0; k < W; k++) {
        int64_t d = 7 * k;
        double v = daily[d];
        if (d - 30 >= 0) v = v + daily[d - 30];
        weekly[k] = v;
    }
}

/* --------------------------------------------------------------- bulk runs */
/*
 * A7 (model.py:176-193) over n parameter tuples with ONE sequential stream, exactly
 * as the reference consumes its global RandomState: sample after sample.
 * out is [4][n][T] doubles (reference layout: per series an [n,T] C-order array).
 */
int vo_run_sampling_sequential(vo_mt *s, int mode, const double *params, int64_t n, int64_t T,
                               int64_t N, double *out)
{
    double *tmp = (double *)malloc(sizeof(double) * 4 * T);
    if (!tmp) return -2;
    for (int64_t i = 0; i < n; i++) {
        int rc = vo_run_single_realization(s, mode, params + 6 * i, T, N, tmp, NULL);
        if (rc) {
            free(tmp);
            return rc;
        }
        for (int k = 0; k < 4; k++) memcpy(out + ((int64_t)k * n + i) * T, tmp + k * T, sizeof(double) * T);
    }
    free(tmp);
    return 0;
}

/*
 * Bulk generator for statistical comparisons and CPU-baseline timing: the same law
 * as the reference (every sample is an independent run of model.py:29-136), fanned
 * out over host threads.  Each sample i draws from its own MT19937 stream seeded
 * with init_genrand(seed + i), so the output does not depend on the thread count.
 * Integer counts only, laid out [series][day][sample] (sample fastest):
 *   daily  [3][T][n] int32 : n_vaccinated, cum_number_vac_received, vaccines_stock
 *   weekly [W][n]    int32 : delta_n_vacc[7k] + delta_n_vacc[7k-30]
 * Either pointer may be NULL (timing runs).
 */
typedef struct {
    const double *params;
    int64_t n, T, N, lo, hi;
    uint32_t seed;
    int32_t *daily, *weekly;
    int rc;
} bulk_job;

static void *bulk_worker(void *arg)
{
    bulk_job *j = (bulk_job *)arg;
    const int64_t T = j->T, n = j->n, W = (T + 6) / 7;
    double *tmp = (double *)malloc(sizeof(double) * 4 * T);
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * 4 * T);
    vo_mt st;
    j->rc = 0;
    for (int64_t i = j->lo; i < j->hi; i++) {
        vo_mt_seed(&st, j->seed + (uint32_t)i);
        int rc = vo_run_single_realization(&st, 0, j->params + 6 * i, T, j->N, tmp, cnt);
        if (rc) {
            j->rc = rc;
            break;
        }
        if (j->daily) {
            for (int64_t d = 0; d < T; d++) {
                j->daily[(0 * T + d) * n + i] = cnt[0 * T + d];
                j->daily[(1 * T + d) * n + i] = cnt[2 * T + d];
                j->daily[(2 * T + d) * n + i] = cnt[3 * T + d];
            }
        }
        if (j->weekly) {
            for (int64_t k = 0; k < W; k++) {
                int64_t d = 7 * k;
                int32_t v = cnt[1 * T + d];
                if (d - 30 >= 0) v += cnt[1 * T + d - 30];
                j->weekly[k * n + i] = v;
            }
        }
    }
    free(tmp);
    free(cnt);
    return NULL;
}

int vo_run_bulk_counts(uint32_t seed, const double *params, int64_t n, int64_t T, int64_t N,
                       int32_t *daily, int32_t *weekly, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    bulk_job jobs[256];
    int64_t chunk = (n + n_threads - 1) / n_threads;
    int started = 0, rc = 0;
    for (int t = 0; t < n_threads; t++) {
        int64_t lo = (int64_t)t * chunk, hi = lo + chunk;
        if (lo >= n) break;
        if (hi > n) hi = n;
        jobs[t] = (bulk_job){params, n, T, N, lo, hi, seed, daily, weekly, 0};
        if (pthread_create(&th[t], NULL, bulk_worker, &jobs[t])) return -3;
        started++;
    }
    for (int t = 0; t < started; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    return rc;
}

int vo_sizeof_mt(void) { return (int)sizeof(vo_mt); }
