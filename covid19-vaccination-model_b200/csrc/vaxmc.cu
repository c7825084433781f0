// vaxmc.cu -- host side of libvaxmc.so: the C ABI declared in include/vaxmc.h.
//
// Orchestrates the reference's run_sampling (model.py:139-228) on one GPU (or one shard of
// the sample range per GPU):
//
//   direct job   (n <= direct_max): stepper<DUMP> -> raw int32 rows [R][n] in HBM ->
//                exact radix select per row (order statistics + sums).
//   streamed job (n  > direct_max): the same on a PILOT (first n_p global samples) to place
//                two narrow integer windows per date around the wanted ranks, then
//                stepper<FOLD> over the whole shard accumulating, per date, an int64 sum,
//                the count below each window and a unit-resolution histogram inside it.
//                The slab is all-integer, so ranks merge it with one NCCL all-reduce(sum)
//                and the order statistics read from it are EXACT and independent of the
//                GPU count.  If a wanted rank falls outside its window (or the window had
//                to be coarsened) vx_job_finalize narrows the windows and asks for another
//                pass (VX_REFINE); the loop always terminates with the exact answer.
//
// There is no CPU fallback: every compute entry point fails with VX_ERR_NODEVICE when no
// CUDA device is usable.
#include "../../include/vaxmc.h"
#include "vaxmc_kernels.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace vx;

namespace {

thread_local std::string g_create_error;

// fixed-point scale of the p_soft_no moments when they travel in the int64 slab: sums of up
// to 8e9 samples of values in [0,1] stay below 2^63, the mean keeps 9 digits
constexpr double STATS_FIX = 1073741824.0;   // 2^30

// ---- NCCL, bound at run time -----------------------------------------------------------------
// The only exchange step of the path is one all-reduce(sum) of the integer slab per pass.  It is
// issued by the library itself on its own stream, right behind the last fold kernel; libnccl is
// dlopen'ed (the copy torch already loaded when the process uses torch.distributed for the
// rendezvous, else the system one), so the library has no link-time dependency on it and
// single-GPU users never touch it.  Prototypes follow nccl.h 2.x (stable since 2.0).
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, vx_nccl_id, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string err;
};
constexpr int NCCL_UINT64 = 5, NCCL_SUM = 0;

NcclApi &nccl_api()
{
    static NcclApi api;
    if (api.handle || !api.err.empty()) return api;
    const char *cands[] = {getenv("VAXMC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
        if (!c || !*c) continue;
        api.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return api; }
    auto sym = [&](const char *n) { void *f = dlsym(api.handle, n); if (!f) api.err = std::string("libnccl lacks ") + n; return f; };
    api.GetUniqueId = (int (*)(void *))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, vx_nccl_id, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclAllReduce");
    api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    api.GetVersion = (int (*)(int *))sym("ncclGetVersion");
    return api;
}

struct Target {
    int64_t rank = 0;
    int64_t lo = 0, hi = 0; // the order statistic is known to lie in [lo, hi]
    bool resolved = false;
    int64_t value = 0;
};

struct WinPlan { // host mirror of one device window + which targets it serves
    Win w{0, 0, 0, 0};
    int tgt[2] = {-1, -1};
};

struct Job {
    vx_config cfg{};
    int64_t shard_first = 0, shard_count = 0;
    int T = 0, W = 0;
    int64_t R = 0;
    bool explicit_params = false;
    bool direct = false;
    bool stats_set = false, piloted = false, aborted = false, finished = false;
    double gstats[4] = {0, 0, 0, 0};      // global: rejections, sum s, sum s^2, samples
    double lstats[4] = {0, 0, 0, 0};      // this shard's (vx_job_param_stats)
    bool lstats_set = false, stats_in_slab = false;
    int64_t n_pilot = 0;
    int64_t vmax[4] = {0, 0, 0, 0};       // per raw-row class: n_vacc, dose, received, stock
    // device buffers
    double *d_params = nullptr;           // explicit tuples [n][6]
    int32_t *d_dump = nullptr;
    double *d_exp = nullptr;
    Win *d_wins = nullptr;
    unsigned long long *d_acc = nullptr;
    int64_t acc_words = 0, acc_cap = 0;
    int bins64 = 0;                       // 64-bit histogram bins (global n >= 2^32), else 32-bit
    int *d_tgt = nullptr;                 // [R][2][2] which wanted rank each window serves
    long long *d_res = nullptr;
    // host state of the streaming job
    std::vector<Target> targets;          // [R][4]: lo_j, lo_j+1, hi_j, hi_j+1
    std::vector<WinPlan> plan;            // [R][2]
    std::vector<int64_t> pilot_stats;     // [R][4] pilot order statistics a_lo,b_lo,a_hi,b_hi
    bool ranks_set = false;
    int64_t n_total = 0;                  // samples folded (all ranks) = number_finished_samples
    int64_t shard_done = 0;               // samples of this shard folded in pass 1
    int n_passes = 0;
    int32_t *d_perm = nullptr;            // pressure-bucket order of the shard (first pass, untimed jobs)
    int64_t perm_first = 0, perm_count = 0;
    // samples [side_first, side_first+side_count) of the shard are stepped by a DUMP launch on the
    // side stream WHILE the latency-bound pilot runs (it leaves > 85 % of the SMs idle) and folded
    // from that matrix once the windows exist, instead of being stepped by the fold kernel later
    int32_t *d_side = nullptr;
    int64_t side_first = 0, side_count = 0;
    bool stats_on_device = false;         // d_stats3 holds this shard's statistics (no host copy yet)
    bool stats_on_side = false;           // ... computed on the side stream: consumers wait for ev_stats
    double *d_stats3 = nullptr;
    bool use_comm = false;                // merge the slab with the context's NCCL communicator, in-stream
    int world = 1;                        // ranks the global sample range is sharded over
    int64_t fold_samples = 0;             // samples stepped by the streaming-fold launches (pass 1)
    cudaEvent_t evM0 = nullptr, evM1 = nullptr;
    double merge_ms = 0.0;
    bool merge_timed = false;
    bool fold_timing_pending = false;
    bool stats_fetch_pending = false;
    int pilot_stage = 0;                  // 0 idle, 1 expected-value job done, 2 pilot in flight, 3 no pilot: plan from bounds
    cudaStream_t pilot_stream = nullptr;
    std::vector<int64_t> pilot_ranks;     // the 4 pilot ranks whose order statistics bracket the two quantiles
    // results (host)
    std::vector<double> mean[4], lower[4], upper[4];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;
    double fold_ms = 0.0, pilot_ms = 0.0;
    int64_t fold_launches = 0;
    std::chrono::steady_clock::time_point t_begin;
};

} // namespace

enum Slot { SLOT_DUMP = 0, SLOT_ACC, SLOT_WINS, SLOT_TGT, SLOT_RES, SLOT_SEL_RANKS, SLOT_SEL_OUT,
            SLOT_SEL_SUMS, SLOT_STATS_PART, SLOT_STATS_MISC, SLOT_PARAMS, SLOT_EXP, SLOT_PERM, SLOT_SORT, SLOT_SIDE, SLOT_TMP_A,
            SLOT_TMP_B, SLOT_TMP_C, SLOT_COUNT };

struct vx_ctx {
    // grow-only device workspace reused across jobs (cudaMalloc/cudaFree of the pilot matrix
    // cost more than the pilot itself); released by vx_trim / vx_destroy
    void *pool[SLOT_COUNT] = {nullptr};
    size_t pool_cap[SLOT_COUNT] = {0};
    bool trace = false;
    bool hostprof = false;                // VX_HOSTPROF=1: host timeline of a job on stderr (no extra synchronisation)
    std::vector<std::pair<const char *, double>> marks;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;          // regime sort runs here, under the latency-bound pilot
    cudaEvent_t side_done = nullptr;      // everything queued on the side stream so far has finished
    cudaEvent_t ev_sorted = nullptr;      // regime sort of the fold's samples done (side stream)
    cudaEvent_t ev_wins = nullptr;        // window table uploaded and slab cleared (main stream)
    cudaEvent_t ev_stats = nullptr;       // parameter statistics of the shard reduced (side stream)
    cudaEvent_t jev[6] = {nullptr};       // the timing events a job uses (created once, not per job)
    std::string err;
    int64_t launches = 0;
    Job *job = nullptr;
    // options
    int64_t max_bins = 262144;            // 2^18: keeps N = 1e6 deliveries (1.5 % of ~4e6 doses) at unit resolution in one pass
    int64_t chunk_samples = 1 << 20;
    double pilot_sigmas = 6.0;
    int64_t use_pilot = 1;
    int64_t sort_samples = 1;
    int64_t fold_pad_smem = 0;            // experiment knob: unused dynamic smem to lower occupancy
    int64_t exact_poisson = 0;            // validation mode: inversion < 10, PTRS above (see kernels)
    int64_t pipe_pilot = 1;               // latency variant of the stepper for launches that cannot fill the GPU
    int64_t select_global = 0;            // 1: the 8-bit radix select, 2: the streamed histogram select, whatever the row length (tests: same answers)
    int64_t side_samples = -1;            // samples stepped beside the pilot (-1 = default = 0: none, see pilot_launch)
    int64_t side_pad_smem = 0;            // experiment knob: dynamic smem of the side DUMP launch (limits its CTAs per SM)
    int sm_count = 148;
    // pinned host staging for the copy-backs (a D2H copy into pageable memory blocks the host
    // until it is done, which would serialise the batched grid driver)
    void *hpin[3] = {nullptr, nullptr, nullptr};
    size_t hpin_cap[3] = {0, 0, 0};
    std::vector<vx_ctx *> subs;           // vx_run_grid: one sub-context (streams + workspace + job) per box in flight
    // multi-GPU: one NCCL communicator per context (vx_comm_init), used by vx_run_bounds_comm
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;
};

namespace {

int fail(vx_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? VX_ERR_NOMEM : VX_ERR_CUDA,       \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,        \
                        __LINE__);                                                               \
    } while (0)

#define CKL()                                                                                    \
    do {                                                                                         \
        ctx->launches++;                                                                         \
        CK(cudaGetLastError());                                                                  \
    } while (0)

int pool_get(vx_ctx *ctx, Slot slot, size_t bytes, void **out)
{
    if (bytes == 0) bytes = 16;
    if (ctx->pool_cap[slot] < bytes) {
        if (ctx->pool[slot]) CK(cudaFree(ctx->pool[slot]));
        ctx->pool[slot] = nullptr;
        ctx->pool_cap[slot] = 0;
        // headroom: the slab's size follows the window widths, which vary by ~10 % from job to job
        // (a re-allocation synchronises the device and stalls a pipelined grid)
        const size_t want = bytes + (slot == SLOT_ACC ? bytes / 2 : bytes / 8);
        cudaError_t e = cudaMalloc(&ctx->pool[slot], want);
        if (e != cudaSuccess) { cudaGetLastError(); CK(cudaMalloc(&ctx->pool[slot], bytes)); ctx->pool_cap[slot] = bytes; }
        else ctx->pool_cap[slot] = want;
    }
    *out = ctx->pool[slot];
    return VX_OK;
}

enum HSlot { HSLOT_SEL = 0, HSLOT_FIN = 1, HSLOT_PLAN = 2 };

int hpin_get(vx_ctx *ctx, HSlot slot, size_t bytes, void **out)
{
    if (bytes == 0) bytes = 64;
    if (ctx->hpin_cap[slot] < bytes) {
        if (ctx->hpin[slot]) CK(cudaFreeHost(ctx->hpin[slot]));
        ctx->hpin[slot] = nullptr; ctx->hpin_cap[slot] = 0;
        CK(cudaMallocHost(&ctx->hpin[slot], bytes + bytes / 4));
        ctx->hpin_cap[slot] = bytes + bytes / 4;
    }
    *out = ctx->hpin[slot];
    return VX_OK;
}

void pool_release(vx_ctx *ctx)
{
    cudaSetDevice(ctx->device);
    for (vx_ctx *sub : ctx->subs) vx_destroy(sub);
    ctx->subs.clear();
    for (int i = 0; i < 3; ++i) {
        if (ctx->hpin[i]) cudaFreeHost(ctx->hpin[i]);
        ctx->hpin[i] = nullptr; ctx->hpin_cap[i] = 0;
    }
    for (int i = 0; i < SLOT_COUNT; ++i) {
        if (ctx->pool[i]) cudaFree(ctx->pool[i]);
        ctx->pool[i] = nullptr;
        ctx->pool_cap[i] = 0;
    }
}

inline void mark(vx_ctx *c, const char *what)
{
    if (c->hostprof)
        c->marks.emplace_back(what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count());
}

void marks_flush(vx_ctx *c)
{
    if (!c->hostprof || c->marks.empty()) return;
    std::string line = "[vx host us]";
    char buf[96];
    for (size_t i = 1; i < c->marks.size(); ++i) {
        snprintf(buf, sizeof buf, " %s %.1f |", c->marks[i].first, c->marks[i].second - c->marks[i - 1].second);
        line += buf;
    }
    snprintf(buf, sizeof buf, " total %.1f", c->marks.back().second - c->marks.front().second);
    fprintf(stderr, "%s%s\n", line.c_str(), buf);
    c->marks.clear();
}

struct Phase {   // VX_TRACE=1: host wall time of each phase (after a stream sync) on stderr
    vx_ctx *c; const char *name; std::chrono::steady_clock::time_point t0;
    Phase(vx_ctx *c_, const char *n) : c(c_), name(n) { if (c->trace) t0 = std::chrono::steady_clock::now(); }
    ~Phase() {
        if (!c->trace) return;
        cudaStreamSynchronize(c->stream);
        fprintf(stderr, "[vx trace] %-22s %9.3f ms\n", name,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

void free_job(vx_ctx *ctx)
{
    Job *j = ctx->job;
    if (!j) return;
    cudaSetDevice(ctx->device);
    delete j;                             // (its events belong to the context)
    ctx->job = nullptr;
}

int validate(vx_ctx *ctx, const vx_config *cfg, bool explicit_params, const double *params)
{
    if (!cfg) return fail(ctx, VX_ERR_ARG, "cfg is NULL");
    if (cfg->T < 1 || cfg->T > 100000) return fail(ctx, VX_ERR_ARG, "T=%d out of range", cfg->T);
    if (cfg->N < 1 || cfg->N > (int64_t)1 << 26) return fail(ctx, VX_ERR_ARG, "N out of range [1, 2^26]");
    if (cfg->n_samples < 1) return fail(ctx, VX_ERR_ARG, "n_samples must be >= 1");
    if (cfg->mode != VX_MODE_STOCHASTIC && cfg->mode != VX_MODE_EXPECTED)
        return fail(ctx, VX_ERR_ARG, "unknown mode %d", cfg->mode);
    if (!(cfg->q_lo >= 0.0 && cfg->q_lo <= 1.0 && cfg->q_hi >= 0.0 && cfg->q_hi <= 1.0))
        return fail(ctx, VX_ERR_ARG, "quantile fractions must lie in [0,1]");
    double nvmax_hi = 0.0;
    if (!explicit_params) {
        const double *b = cfg->bounds;
        for (int i = 0; i < 12; ++i)
            if (!std::isfinite(b[i])) return fail(ctx, VX_ERR_ARG, "non-finite bound");
        if (std::min(b[4], b[5]) < 0.0)
            return fail(ctx, VX_ERR_ARG, "lam < 0: negative pressure bound (numpy raises ValueError, model.py:116)");
        if (std::min(b[6], b[7]) <= 0.0) return fail(ctx, VX_ERR_ARG, "tau bounds must be > 0");
        if (std::min(b[0], b[1]) < 0.0 || std::min(b[2], b[3]) < 0.0 || std::min(b[8], b[9]) < 0.0 ||
            std::min(b[10], b[11]) < 0.0)
            return fail(ctx, VX_ERR_ARG, "negative fraction bound");
        nvmax_hi = std::max(b[10], b[11]);
    } else {
        if (!params) return fail(ctx, VX_ERR_ARG, "params is NULL");
        for (int64_t i = 0; i < cfg->n_samples; ++i) {
            const double *p = params + 6 * i;
            for (int k = 0; k < 6; ++k)
                if (!std::isfinite(p[k])) return fail(ctx, VX_ERR_ARG, "non-finite parameter in tuple %lld", (long long)i);
            if (p[0] + p[1] > 1.0) return fail(ctx, VX_ERR_ARG, "assert p_pro + p_anti <= 1.0 (model.py:63), tuple %lld", (long long)i);
            if (p[0] < 0 || p[1] < 0 || p[2] < 0 || p[3] <= 0 || p[4] < 0 || p[5] < 0)
                return fail(ctx, VX_ERR_ARG, "parameter out of range in tuple %lld", (long long)i);
            nvmax_hi = std::max(nvmax_hi, p[5]);
        }
    }
    const double Wd = (double)((cfg->T + 6) / 7);
    if (Wd * (std::floor(nvmax_hi * (double)cfg->N) + 1.0) >= 134217728.0)
        return fail(ctx, VX_ERR_ARG, "cumulative deliveries may exceed 2^27 doses: unsupported size");
    return VX_OK;
}

inline int series_of_row(const Job *j, int64_t row)
{
    if (row < j->T) return 0;
    if (row < j->T + j->W) return 1;
    if (row < j->T + 2 * j->W) return 2;
    return 3;
}

int alloc_results(Job *j)
{
    const int len[4] = {j->T, j->W, j->T, j->T};
    for (int s = 0; s < 4; ++s) {
        j->mean[s].assign(len[s], 0.0);
        j->lower[s].assign(len[s], 0.0);
        j->upper[s].assign(len[s], 0.0);
    }
    return 0;
}

// (series, date index) -> raw row
inline int64_t row_of(const Job *j, int s, int i)
{
    switch (s) {
    case 0: return i;
    case 1: return j->T + i;
    case 2: return j->T + j->W + i / 7;
    default: return j->T + 2 * (int64_t)j->W + i;
    }
}

void fill_from_counts(Job *j, const std::vector<int64_t> &ord /*[R][4]*/,
                      const std::vector<long long> &sums, int64_t n)
{
    int64_t p0, n0, p1, n1;
    double g0, g1;
    vx_host_quantile_index(n, j->cfg.q_lo, &p0, &n0, &g0);
    vx_host_quantile_index(n, j->cfg.q_hi, &p1, &n1, &g1);
    const int len[4] = {j->T, j->W, j->T, j->T};
    for (int s = 0; s < 4; ++s)
        for (int i = 0; i < len[s]; ++i) {
            const int64_t r = row_of(j, s, i);
            const int64_t *o = &ord[r * 4];
            j->lower[s][i] = vx_host_lerp(vx_host_count_to_value(s, o[0], j->cfg.N),
                                          vx_host_count_to_value(s, o[1], j->cfg.N), g0);
            j->upper[s][i] = vx_host_lerp(vx_host_count_to_value(s, o[2], j->cfg.N),
                                          vx_host_count_to_value(s, o[3], j->cfg.N), g1);
            j->mean[s][i] = vx_host_count_to_value(s, sums[r], j->cfg.N) / (double)n;
        }
}

int copy_results(vx_ctx *ctx, vx_result *res)
{
    Job *j = ctx->job;
    const int len[4] = {j->T, j->W, j->T, j->T};
    for (int s = 0; s < 4; ++s) {
        if (res->mean[s]) memcpy(res->mean[s], j->mean[s].data(), sizeof(double) * len[s]);
        if (res->lower[s]) memcpy(res->lower[s], j->lower[s].data(), sizeof(double) * len[s]);
        if (res->upper[s]) memcpy(res->upper[s], j->upper[s].data(), sizeof(double) * len[s]);
    }
    return VX_OK;
}

void fill_result_scalars(vx_ctx *ctx, vx_result *res)
{
    Job *j = ctx->job;
    res->n_finished = j->n_total;
    res->aborted = j->aborted ? 1 : 0;
    res->n_passes = j->n_passes;
    res->n_rejected = (int64_t)j->gstats[0];
    const double n = j->gstats[3];
    if (n > 0) {
        const double m = j->gstats[1] / n;
        const double var = std::max(j->gstats[2] / n - m * m, 0.0);
        res->soft_no_mean = m;
        res->soft_no_std = std::sqrt(var);
    } else {
        res->soft_no_mean = res->soft_no_std = 0.0;
    }
}

SimArgs make_args(const Job *j, int64_t first, int64_t count, const double *d_params)
{
    SimArgs a{};
    memcpy(a.bd.b, j->cfg.bounds, sizeof a.bd.b);
    a.params = d_params;
    a.seed = j->cfg.seed;
    a.keys = philox_keys(j->cfg.seed);
    a.first = first;
    a.count = count;
    a.N = j->cfg.N;
    a.T = j->T;
    a.W = j->W;
    return a;
}

int upload_rcp(vx_ctx *ctx)
{
    float h[72];
    h[0] = 0.f;
    for (int k = 1; k < 72; ++k) h[k] = 1.0f / (float)k;
    CK(cudaMemcpyToSymbol(c_rcp, h, sizeof h));
    return VX_OK;
}

// ---- window planning ---------------------------------------------------------------
inline uint32_t shift_for(int64_t width, int64_t max_bins)
{
    uint32_t s = 0;
    while (((width + ((int64_t)1 << s) - 1) >> s) > max_bins) ++s;
    return s;
}

// Build the device windows of the next pass from the unresolved targets.
int plan_windows(vx_ctx *ctx, bool from_pilot)
{
    Job *j = ctx->job;
    const int64_t R = j->R;
    j->plan.assign(R * 2, WinPlan());
    uint64_t off = 0;
    for (int64_t r = 0; r < R; ++r) {
        Target *t = &j->targets[r * 4];
        int nw = 0;
        if (from_pilot) {
            const int64_t *ps = &j->pilot_stats[r * 4];
            for (int q = 0; q < 2; ++q) {
                int64_t a = ps[2 * q], b = ps[2 * q + 1];
                a = std::max(a, t[2 * q].lo);
                b = std::min(std::max(b, a), t[2 * q].hi);
                WinPlan &wp = j->plan[r * 2 + q];
                const int64_t width = b - a + 1;
                const uint32_t sh = shift_for(width, ctx->max_bins);
                const int64_t nb = (width + ((int64_t)1 << sh) - 1) >> sh;
                wp.w.a = (int)a; wp.w.shift = sh; wp.w.len = (uint32_t)(nb << sh);
                wp.tgt[0] = 2 * q; wp.tgt[1] = 2 * q + 1;
            }
            nw = 2;
        } else {
            bool used[4] = {false, false, false, false};
            for (int k = 0; k < 4 && nw < 2; ++k) {
                if (t[k].resolved || used[k]) continue;
                WinPlan &wp = j->plan[r * 2 + nw];
                int64_t lo = t[k].lo, hi = t[k].hi;
                wp.tgt[0] = k; used[k] = true;
                for (int m = k + 1; m < 4; ++m) {
                    if (t[m].resolved || used[m]) continue;
                    const int64_t ulo = std::min(lo, t[m].lo), uhi = std::max(hi, t[m].hi);
                    if (uhi - ulo + 1 <= (hi - lo + 1) + (t[m].hi - t[m].lo + 1) + 1) {
                        wp.tgt[1] = m; used[m] = true; lo = ulo; hi = uhi;
                    }
                    break; // at most two targets per window, neighbours only
                }
                const int64_t width = hi - lo + 1;
                const uint32_t sh = shift_for(width, ctx->max_bins);
                const int64_t nb = (width + ((int64_t)1 << sh) - 1) >> sh;
                wp.w.a = (int)lo; wp.w.shift = sh; wp.w.len = (uint32_t)(nb << sh);
                ++nw;
            }
        }
        for (int q = 0; q < 2; ++q) {
            WinPlan &wp = j->plan[r * 2 + q];
            if (off + (wp.w.len >> wp.w.shift) > 0xfffffff0ull)
                return fail(ctx, VX_ERR_NOMEM, "window histograms exceed 2^32 bins");
            wp.w.off = (uint32_t)off;
            off += wp.w.len >> wp.w.shift;
        }
    }
    const int64_t words = ACC_HDR + 3 * R + (j->bins64 ? (int64_t)off : (int64_t)((off + 1) / 2));
    { void *p_ = nullptr; int rc_ = pool_get(ctx, SLOT_ACC, sizeof(unsigned long long) * words, &p_); if (rc_) return rc_; j->d_acc = (unsigned long long *)p_; }
    j->acc_words = words;
    // window table and target kinds: built in pinned memory and uploaded with ONE copy (they sit
    // between the pilot and the fold, where every host microsecond is an idle GPU)
    const size_t wins_bytes = sizeof(Win) * R * 2, tgt_bytes = sizeof(int) * R * 4;
    void *hp_ = nullptr;
    { void *p_ = nullptr; int rc_;
      rc_ = pool_get(ctx, SLOT_WINS, wins_bytes + tgt_bytes, &p_); if (rc_) return rc_;
      j->d_wins = (Win *)p_; j->d_tgt = (int *)((char *)p_ + wins_bytes);
      rc_ = pool_get(ctx, SLOT_RES, sizeof(long long) * R * 4, &p_); if (rc_) return rc_; j->d_res = (long long *)p_;
      rc_ = hpin_get(ctx, HSLOT_PLAN, wins_bytes + tgt_bytes, &hp_); if (rc_) return rc_; }
    Win *hw = (Win *)hp_;
    int *hk = (int *)((char *)hp_ + wins_bytes);
    for (int64_t i = 0; i < R * 2; ++i) {
        hw[i] = j->plan[i].w;
        hk[i * 2] = j->plan[i].tgt[0]; hk[i * 2 + 1] = j->plan[i].tgt[1];
    }
    CK(cudaMemcpyAsync(j->d_wins, hp_, wins_bytes + tgt_bytes, cudaMemcpyHostToDevice, ctx->stream));
    return VX_OK;
}

// Exact order statistics + row sums of an int32 matrix: launch (kernel + copy-back into pinned
// host memory, all asynchronous on `st`) and collect (once `st` has been synchronised).
int select_launch(vx_ctx *ctx, const int32_t *d_data, int64_t rows, int64_t ld, int64_t n,
                  const std::vector<int64_t> &ranks, cudaStream_t st)
{
    const int K = (int)ranks.size();
    int64_t *d_ranks = nullptr;
    int32_t *d_out = nullptr;
    long long *d_sums = nullptr;
    void *h = nullptr;
    { void *p_ = nullptr; int rc_;
      rc_ = pool_get(ctx, SLOT_SEL_RANKS, sizeof(int64_t) * K, &p_); if (rc_) return rc_; d_ranks = (int64_t *)p_;
      rc_ = pool_get(ctx, SLOT_SEL_OUT, sizeof(int32_t) * rows * K, &p_); if (rc_) return rc_; d_out = (int32_t *)p_;
      rc_ = pool_get(ctx, SLOT_SEL_SUMS, sizeof(long long) * rows, &p_); if (rc_) return rc_; d_sums = (long long *)p_;
      rc_ = hpin_get(ctx, HSLOT_SEL, sizeof(long long) * rows + sizeof(int32_t) * rows * K, &h); if (rc_) return rc_; }
    CK(cudaMemcpyAsync(d_ranks, ranks.data(), sizeof(int64_t) * K, cudaMemcpyHostToDevice, st));
    // histogram-refinement select: rows that fit in shared memory (the pilot: 64 KB per row) are
    // staged and selected on chip after one read, longer rows are streamed once per pass;
    // "select_global" = 1 forces the 8-bit radix kernel (cross-check), 2 the streamed variant
    const size_t smem = sizeof(int32_t) * (size_t)((n + 3) & ~(int64_t)3) + sizeof(uint32_t) * SELS_BINS;
    if (ctx->select_global == 1) {
        vx_select_kernel<int32_t><<<(unsigned)rows, SEL_THREADS, 0, st>>>(d_data, ld, n, d_ranks, K, d_out, d_sums);
    } else if (smem <= 200 * 1024 && ctx->select_global == 0) {
        static bool attr_set[64] = {false};
        if (ctx->device < 64 && !attr_set[ctx->device]) {
            CK(cudaFuncSetAttribute(vx_select_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            // all of the SM's L1/shared array as shared memory: two 80 KB rows resident per SM
            CK(cudaFuncSetAttribute(vx_select_hist_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr_set[ctx->device] = true;
        }
        vx_select_hist_kernel<true><<<(unsigned)rows, SELS_THREADS, smem, st>>>(d_data, ld, n, d_ranks, K, d_out, d_sums);
    } else {
        vx_select_hist_kernel<false><<<(unsigned)rows, SELS_THREADS, sizeof(uint32_t) * SELS_BINS, st>>>(d_data, ld, n, d_ranks, K, d_out, d_sums);
    }
    CKL();
    CK(cudaMemcpyAsync(h, d_sums, sizeof(long long) * rows, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((char *)h + sizeof(long long) * rows, d_out, sizeof(int32_t) * rows * K, cudaMemcpyDeviceToHost, st));
    return VX_OK;
}

void select_collect(vx_ctx *ctx, int64_t rows, int K, std::vector<int64_t> &ord, std::vector<long long> &sums)
{
    const long long *hs = (const long long *)ctx->hpin[HSLOT_SEL];
    const int32_t *ho = (const int32_t *)((const char *)ctx->hpin[HSLOT_SEL] + sizeof(long long) * rows);
    sums.assign(hs, hs + rows);
    ord.resize(rows * K);
    for (int64_t i = 0; i < rows * K; ++i) ord[i] = ho[i];
}

int run_select_i32(vx_ctx *ctx, const int32_t *d_data, int64_t rows, int64_t ld, int64_t n,
                   const std::vector<int64_t> &ranks, std::vector<int64_t> &ord,
                   std::vector<long long> &sums)
{
    int rc = select_launch(ctx, d_data, rows, ld, n, ranks, ctx->stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    select_collect(ctx, rows, (int)ranks.size(), ord, sums);
    return VX_OK;
}

int need_device(vx_ctx *ctx)
{
    if (!ctx) return VX_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return VX_OK;
}

int launch_dump(vx_ctx *ctx, Job *j, int64_t first, int64_t count, const double *d_params,
                int32_t *d_dump, int64_t ld, cudaStream_t st = nullptr, const int32_t *d_perm = nullptr,
                size_t pad_smem = 0)
{
    if (!st) st = ctx->stream;
    // a launch that cannot fill the GPU is a serial T-day chain per warp: take the latency variant
    const bool latency = ctx->pipe_pilot && count <= (int64_t)ctx->sm_count * 4 * SIM_THREADS;
    SimArgs a = make_args(j, first, count, d_params);
    a.dump = d_dump;
    a.dump_stride = ld;
    a.perm = d_perm;
    const int64_t grid = (count + SIM_THREADS - 1) / SIM_THREADS;
    if (ctx->exact_poisson) vx_sim_kernel<SIM_DUMP, 1><<<(unsigned)grid, SIM_THREADS, 0, st>>>(a);
    else if (latency) vx_sim_kernel<SIM_DUMP, 0, 1><<<(unsigned)grid, SIM_THREADS, 0, st>>>(a);
    else vx_sim_kernel<SIM_DUMP, 0><<<(unsigned)grid, SIM_THREADS, pad_smem, st>>>(a);
    CKL();
    return VX_OK;
}

// P(p_pro + p_anti <= 1) for independent uniforms on the two bound intervals: the acceptance
// probability of model.py:275-287's rejection step (exact area of the box under the line).
double acceptance_probability(const double *b)
{
    const double x0 = std::min(b[0], b[1]), x1 = std::max(b[0], b[1]);
    const double y0 = std::min(b[2], b[3]), y1 = std::max(b[2], b[3]);
    if (x1 + y1 <= 1.0) return 1.0;
    if (x0 + y0 > 1.0) return 0.0;
    if (x1 == x0 || y1 == y0) {      // degenerate box: a segment (or a point)
        if (x1 == x0 && y1 == y0) return 1.0;
        const double lo = (x1 == x0) ? y0 : x0, hi = (x1 == x0) ? y1 : x1, c = 1.0 - ((x1 == x0) ? x0 : y0);
        return std::min(std::max((c - lo) / (hi - lo), 0.0), 1.0);
    }
    // integrate over x the accepted length of y: clamp(1 - x - y0, 0, y1 - y0)
    const double h = y1 - y0, xa = std::min(std::max(1.0 - y1, x0), x1), xb = std::min(std::max(1.0 - y0, x0), x1);
    // x in [x0, xa]: full height; x in [xa, xb]: 1 - x - y0 (linear); beyond xb: nothing
    const double area = (xa - x0) * h + ((1.0 - y0) * (xb - xa) - 0.5 * (xb * xb - xa * xa));
    return area / ((x1 - x0) * h);
}

// This shard's rejection count and moments of p_soft_no, left ON THE DEVICE (d_stats3).
int launch_param_stats(vx_ctx *ctx, Job *j, cudaStream_t st = nullptr)
{
    if (!st) st = ctx->stream;
    const int64_t n = j->shard_count, grid = (n + 255) / 256;
    unsigned long long *d_stop = nullptr;
    double *d_part = nullptr, *d_out = nullptr;
    { void *p_ = nullptr; int rc_;
      rc_ = pool_get(ctx, SLOT_STATS_PART, sizeof(double) * 3 * grid, &p_); if (rc_) return rc_; d_part = (double *)p_;
      rc_ = pool_get(ctx, SLOT_STATS_MISC, 64, &p_); if (rc_) return rc_; d_stop = (unsigned long long *)p_; d_out = (double *)p_ + 4; }
    const unsigned long long hstop[2] = {0ull, (unsigned long long)(10 * j->cfg.n_samples)};
    CK(cudaMemcpyAsync(d_stop, hstop, 16, cudaMemcpyHostToDevice, st));
    Bounds bd;
    memcpy(bd.b, j->cfg.bounds, sizeof bd.b);
    vx_param_stats_kernel<<<(unsigned)grid, 256, 0, st>>>(bd, j->cfg.seed, j->shard_first, n, d_stop, d_part);
    CKL();
    vx_reduce_stats_kernel<<<1, 1024, 0, st>>>(d_part, grid, d_out);
    CKL();
    j->d_stats3 = d_out;
    return VX_OK;
}

// Counting sort of samples [first, first+count) by pressure bucket into d_perm (local ids).
int launch_regime_sort(vx_ctx *ctx, Job *j, int64_t first, int64_t count, cudaStream_t st, int32_t **perm_out)
{
    void *p_ = nullptr;
    int rc = pool_get(ctx, SLOT_PERM, sizeof(int32_t) * count, &p_);
    if (rc) return rc;
    int32_t *d_perm = (int32_t *)p_;
    rc = pool_get(ctx, SLOT_SORT, sizeof(uint32_t) * SORT_BUCKETS, &p_);
    if (rc) return rc;
    uint32_t *d_cnt = (uint32_t *)p_;
    const unsigned sgrid = (unsigned)((count + 255) / 256);
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * SORT_BUCKETS, st));
    vx_bucket_count_kernel<<<sgrid, 256, 0, st>>>(j->cfg.seed, first, count, d_cnt);
    CKL();
    vx_bucket_scan_kernel<<<1, SORT_BUCKETS, 0, st>>>(d_cnt);
    CKL();
    vx_bucket_scatter_kernel<<<sgrid, 256, 0, st>>>(j->cfg.seed, first, count, d_cnt, d_perm);
    CKL();
    *perm_out = d_perm;
    return VX_OK;
}

// Pilot size.  The pilot is replicated on every rank and its stepper is latency-bound (one
// 564-day dependency chain per warp, ~0.4 ms however few samples) plus ~1e-5 ms per sample
// (stepper + select); the window width -- hence the fold's slow path and the slab the ranks
// all-reduce -- goes with n_p^-1/2.  Measured on B200 (USA shape): fold_ms(n_p) = F (1 + 4.1
// n_p^-1/2) with F = 4.1 ms per 1e6 samples, so the cost  0.5 + 1e-5 n_p + fold_ms  is minimal
// at n_p = (0.84 n)^(2/3): 16384 at 1e6 samples (measured optimum), 32768 at 8e6 (measured),
// capped at 2^18.  Independently the +-6 sigma bracket of the outer quantile has to stay
// inside the pilot sample: n_p * q_tail >~ 64.
int64_t default_pilot(double q_lo, double q_hi, int64_t n)
{
    double tail = std::min(std::min(q_lo, 1.0 - q_lo), std::min(q_hi, 1.0 - q_hi));
    tail = std::max(tail, 1e-9);
    int64_t p = 16384;
    while ((double)p * tail < 64.0 && p < ((int64_t)1 << 20)) p <<= 1;
    const double opt = std::pow(0.84 * (double)n, 2.0 / 3.0);
    while ((double)(2 * p) <= opt && p < ((int64_t)1 << 18)) p <<= 1;
    return p;
}

// ---- the expected-value job (always direct) ----------------------------------------
int run_expected(vx_ctx *ctx)
{
    Job *j = ctx->job;
    const int64_t n = j->cfg.n_samples;
    const int64_t RE = 4 * (int64_t)j->T + j->W;
    if ((double)RE * (double)n * 8.0 > 64e9)
        return fail(ctx, VX_ERR_ARG, "expected-value mode materialises fp64 trajectories: n_samples too large");
    { void *p_ = nullptr; int rc_ = pool_get(ctx, SLOT_EXP, sizeof(double) * RE * n, &p_); if (rc_) return rc_; j->d_exp = (double *)p_; }
    SimArgs a = make_args(j, 0, n, j->d_params);
    vx_expected_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(a, j->d_exp, n);
    CKL();
    int64_t p0, n0, p1, n1; double g0, g1;
    vx_host_quantile_index(n, j->cfg.q_lo, &p0, &n0, &g0);
    vx_host_quantile_index(n, j->cfg.q_hi, &p1, &n1, &g1);
    const int64_t hr[4] = {p0, n0, p1, n1};
    int64_t *d_ranks = nullptr; double *d_out = nullptr, *d_sums = nullptr;
    { void *p_ = nullptr; int rc_;       // workspace slots: nothing to free on an error path
      rc_ = pool_get(ctx, SLOT_TMP_A, sizeof hr, &p_); if (rc_) return rc_; d_ranks = (int64_t *)p_;
      rc_ = pool_get(ctx, SLOT_TMP_B, sizeof(double) * RE * 4, &p_); if (rc_) return rc_; d_out = (double *)p_;
      rc_ = pool_get(ctx, SLOT_TMP_C, sizeof(double) * RE, &p_); if (rc_) return rc_; d_sums = (double *)p_; }
    CK(cudaMemcpyAsync(d_ranks, hr, sizeof hr, cudaMemcpyHostToDevice, ctx->stream));
    vx_select_kernel<double><<<(unsigned)RE, SEL_THREADS, 0, ctx->stream>>>(j->d_exp, n, n, d_ranks, 4, d_out, d_sums);
    CKL();
    std::vector<double> ho(RE * 4), hs(RE);
    CK(cudaMemcpyAsync(ho.data(), d_out, sizeof(double) * RE * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hs.data(), d_sums, sizeof(double) * RE, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    alloc_results(j);
    const int len[4] = {j->T, j->W, j->T, j->T};
    for (int s = 0; s < 4; ++s)
        for (int i = 0; i < len[s]; ++i) {
            const int64_t r = (s == 1) ? 4 * (int64_t)j->T + i : (int64_t)s * j->T + i;
            j->lower[s][i] = vx_host_lerp(ho[r * 4], ho[r * 4 + 1], g0);
            j->upper[s][i] = vx_host_lerp(ho[r * 4 + 2], ho[r * 4 + 3], g1);
            j->mean[s][i] = hs[r] / (double)n;
        }
    j->n_total = n;
    j->finished = true;
    return VX_OK;
}

} // namespace

// =====================================================================================
extern "C" {

const char *vx_version(void) { return "vaxmc 0.2 (sm_100a)"; }

int vx_abi_version(void) { return VX_ABI_VERSION; }

int64_t vx_sizeof(int which)
{
    switch (which) {
    case 0: return (int64_t)sizeof(vx_config);
    case 1: return (int64_t)sizeof(vx_result);
    case 2: return (int64_t)sizeof(vx_nccl_id);
    default: return -1;
    }
}

int vx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int create_ctx(vx_ctx **out, int device, bool side_high_priority)
{
    vx_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, VX_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, VX_ERR_NODEVICE, "no CUDA device: libvaxmc has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, VX_ERR_ARG, "device %d not in [0,%d)", device, n);
    vx_ctx *c = new vx_ctx();
    c->device = device;
    c->trace = getenv("VX_TRACE") != nullptr;
    c->hostprof = getenv("VX_HOSTPROF") != nullptr;
    cudaError_t e = cudaSetDevice(device);
    int lo_pri = 0, hi_pri = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    // a grid sub-context runs its box's pilot on the side stream, which then outranks the main
    // streams: the pilot is short and gates a long fold, other boxes' folds should not delay it.
    // A whole-job context is the other way round: pilot and fold stepper on the main stream outrank
    // the filler work (side DUMP, matrix folds) of the side stream.
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, side_high_priority ? lo_pri : hi_pri);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, side_high_priority ? hi_pri : lo_pri);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->side_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_sorted, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_wins, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_stats, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        int sm = 0;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
        if (sm > 0) c->sm_count = sm;
        ctx = c;
        if (upload_rcp(ctx) != VX_OK) { g_create_error = c->err; cudaStreamDestroy(c->stream); delete c; return VX_ERR_CUDA; }
        *out = c;
        return VX_OK;
    }
    delete c;
    return fail(nullptr, VX_ERR_CUDA, "vx_create: %s", cudaGetErrorString(e));
}

int vx_create(vx_ctx **out, int device) { return create_ctx(out, device, false); }

void vx_destroy(vx_ctx *ctx)
{
    if (!ctx) return;
    free_job(ctx);
    vx_comm_destroy(ctx);
    pool_release(ctx);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->side_done) cudaEventDestroy(ctx->side_done);
    if (ctx->ev_sorted) cudaEventDestroy(ctx->ev_sorted);
    if (ctx->ev_wins) cudaEventDestroy(ctx->ev_wins);
    if (ctx->ev_stats) cudaEventDestroy(ctx->ev_stats);
    for (int i = 0; i < 6; ++i) if (ctx->jev[i]) cudaEventDestroy(ctx->jev[i]);
    delete ctx;
}

const char *vx_last_error(vx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int64_t vx_launch_count(vx_ctx *ctx) { return ctx ? ctx->launches : 0; }

int vx_trim(vx_ctx *ctx)
{
    if (!ctx) return VX_ERR_ARG;
    if (ctx->job) return fail(ctx, VX_ERR_STATE, "vx_trim while a job is open");
    pool_release(ctx);
    return VX_OK;
}

int vx_set_option(vx_ctx *ctx, const char *name, double value)
{
    if (!ctx || !name) return VX_ERR_ARG;
    const std::string s(name);
    if (s == "max_bins") { if (value < 2 || value > (1 << 24)) return fail(ctx, VX_ERR_ARG, "max_bins out of range"); ctx->max_bins = (int64_t)value; }
    else if (s == "chunk_samples") { if (value < 256) return fail(ctx, VX_ERR_ARG, "chunk_samples too small"); ctx->chunk_samples = (int64_t)value; }
    else if (s == "pilot_sigmas") { if (!(value >= 0)) return fail(ctx, VX_ERR_ARG, "pilot_sigmas < 0"); ctx->pilot_sigmas = value; }
    else if (s == "use_pilot") ctx->use_pilot = value != 0;
    else if (s == "sort_samples") ctx->sort_samples = value != 0;
    else if (s == "fold_pad_smem") ctx->fold_pad_smem = (int64_t)value;
    else if (s == "exact_poisson") ctx->exact_poisson = value != 0;
    else if (s == "side_samples") ctx->side_samples = (int64_t)value;
    else if (s == "side_pad_smem") ctx->side_pad_smem = (int64_t)value;
    else if (s == "select_global") ctx->select_global = (int64_t)value;
    else if (s == "pipe_pilot") ctx->pipe_pilot = value != 0;
    else return fail(ctx, VX_ERR_ARG, "unknown option %s", name);
    return VX_OK;
}

// ---- host-only helpers -----------------------------------------------------------------
double vx_host_lerp(double a, double b, double t)
{
    // numpy/lib/_function_base_impl.py:_lerp
    const double diff = b - a;
    double r = a + diff * t;
    if (t >= 0.5) r = b - diff * (1 - t);
    return r;
}

void vx_host_quantile_index(int64_t n, double q, int64_t *prev, int64_t *next, double *gamma)
{
    // method="linear": virtual index (n-1)*q; _get_indexes + _get_gamma
    const double vi = (double)(n - 1) * q;
    long long pi, ni;
    quantile_ranks((long long)n, q, pi, ni);
    *prev = pi; *next = ni;
    *gamma = vi - std::floor(vi);
}

double vx_host_count_to_value(int series, int64_t count, int64_t N)
{
    const double c = (double)count, Nd = (double)N;
    switch (series) {
    case 0: return (c / Nd) * 100;   // fract_pop_vaccinated * 100          model.py:112,124
    case 1: return c * 1e6 / Nd;     // delta_n_vacc * 1e6 / N              model.py:125
    default: return c * 100 / Nd;    // cum received / stock * 100 / N      model.py:126-127
    }
}

int32_t vx_host_weeks(int32_t T) { return (T + 6) / 7; }
int64_t vx_host_rows(int32_t T) { return 2 * (int64_t)T + 2 * (int64_t)((T + 6) / 7); }

// ---- staged job --------------------------------------------------------------------------
static int job_begin_common(vx_ctx *ctx, const vx_config *cfg, int64_t shard_first,
                            int64_t shard_count, const double *h_params)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    free_job(ctx);
    rc = validate(ctx, cfg, h_params != nullptr, h_params);
    if (rc) return rc;
    if (shard_first < 0 || shard_count < 0 || shard_first + shard_count > cfg->n_samples)
        return fail(ctx, VX_ERR_ARG, "shard [%lld,+%lld) outside [0,%lld)", (long long)shard_first,
                    (long long)shard_count, (long long)cfg->n_samples);
    Job *j = new Job();
    ctx->job = j;
    j->cfg = *cfg;
    j->shard_first = shard_first;
    j->shard_count = shard_count;
    j->T = cfg->T;
    j->W = (cfg->T + 6) / 7;
    j->R = 2 * (int64_t)j->T + 2 * (int64_t)j->W;
    j->explicit_params = h_params != nullptr;
    j->bins64 = cfg->n_samples >= ((int64_t)1 << 32) ? 1 : 0;
    const int64_t dmax = cfg->direct_max > 0 ? cfg->direct_max : VX_DIRECT_MAX_DEFAULT;   // crossover of direct vs pilot+fold on B200 (USA shape)
    j->direct = (cfg->mode == VX_MODE_EXPECTED) || cfg->n_samples <= dmax;
    j->t_begin = std::chrono::steady_clock::now();
    double nvmax_hi = std::max(cfg->bounds[10], cfg->bounds[11]);
    if (h_params) {
        nvmax_hi = 0;
        for (int64_t i = 0; i < cfg->n_samples; ++i) nvmax_hi = std::max(nvmax_hi, h_params[6 * i + 5]);
        { void *p_ = nullptr; int rc_ = pool_get(ctx, SLOT_PARAMS, sizeof(double) * 6 * cfg->n_samples, &p_); if (rc_) return rc_; j->d_params = (double *)p_; }
        CK(cudaMemcpyAsync(j->d_params, h_params, sizeof(double) * 6 * cfg->n_samples, cudaMemcpyHostToDevice, ctx->stream));
        j->gstats[3] = (double)cfg->n_samples;
        j->stats_set = true;
    }
    const int64_t recv_max = (int64_t)j->W * ((int64_t)std::floor(nvmax_hi * (double)cfg->N) + 1);
    j->vmax[0] = cfg->N; j->vmax[1] = 2 * cfg->N; j->vmax[2] = recv_max; j->vmax[3] = recv_max;
    for (int i = 0; i < 6; ++i)
        if (!ctx->jev[i]) CK(cudaEventCreate(&ctx->jev[i]));
    j->ev0 = ctx->jev[0]; j->ev1 = ctx->jev[1]; j->evA = ctx->jev[2]; j->evB = ctx->jev[3];
    j->evM0 = ctx->jev[4]; j->evM1 = ctx->jev[5];
    CK(cudaEventRecord(j->ev0, ctx->stream));
    return VX_OK;
}

int vx_job_begin(vx_ctx *ctx, const vx_config *cfg, int64_t shard_first, int64_t shard_count)
{
    return job_begin_common(ctx, cfg, shard_first, shard_count, nullptr);
}

int vx_job_param_stats(vx_ctx *ctx, double stats[4])
{
    if (!ctx || !ctx->job || !stats) return fail(ctx, VX_ERR_STATE, "no job");
    Job *j = ctx->job;
    int rc = need_device(ctx);
    if (rc) return rc;
    Phase ph_(ctx, "param_stats");
    stats[0] = stats[1] = stats[2] = 0.0;
    stats[3] = (double)j->shard_count;
    struct Keep { Job *j; double *s; ~Keep() { memcpy(j->lstats, s, sizeof j->lstats); j->lstats_set = true; } } keep_{j, stats};
    if (j->explicit_params || j->shard_count == 0) return VX_OK;
    const double *b = j->cfg.bounds;
    if (std::min(b[0], b[1]) + std::min(b[2], b[3]) > 1.0) {
        // no acceptable pair exists: the reference burns 10*n_rep+1 rejections and aborts
        stats[0] = 10.0 * (double)j->cfg.n_samples + 1.0;
        return VX_OK;
    }
    rc = launch_param_stats(ctx, j);
    if (rc) return rc;
    double *d_out = j->d_stats3;
    CK(cudaMemcpyAsync(stats, d_out, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_job_set_param_stats(vx_ctx *ctx, const double stats[4])
{
    if (!ctx || !ctx->job || !stats) return fail(ctx, VX_ERR_STATE, "no job");
    Job *j = ctx->job;
    memcpy(j->gstats, stats, sizeof j->gstats);
    j->stats_set = true;
    // model.py:282-287: abort when the rejection count exceeds 10 * n_rep
    j->aborted = stats[0] > 10.0 * (double)j->cfg.n_samples;
    return j->aborted ? 1 : 0;
}

int vx_job_is_direct(vx_ctx *ctx) { return (ctx && ctx->job && ctx->job->direct) ? 1 : 0; }

// The pilot in two halves.  pilot_launch queues everything up to the copy-back of the pilot's
// order statistics (asynchronous; `on_side`: on the context's side stream, as the grid driver
// wants it -- then nothing else is stepped beside it); pilot_complete, once that stream has been
// synchronised, turns them into the window plan (host) and uploads it on the main stream.
static int pilot_launch(vx_ctx *ctx, bool on_side)
{
    if (!ctx || !ctx->job) return fail(ctx, VX_ERR_STATE, "no job");
    Job *j = ctx->job;
    if (j->aborted) return fail(ctx, VX_ERR_STATE, "job aborted by the rejection rule");
    int rc = need_device(ctx);
    if (rc) return rc;
    j->piloted = true;
    j->pilot_stage = 1;
    cudaStream_t ps = on_side ? ctx->side : ctx->stream;
    j->pilot_stream = ps;
    CK(cudaEventRecord(j->evA, ps));
    if (j->cfg.mode == VX_MODE_EXPECTED) return run_expected(ctx);

    const int64_t n = j->cfg.n_samples, R = j->R;
    // host-side preparation (result vectors, the certain bounds of every order statistic); done
    // AFTER the launches below are queued, while the GPU already runs the pilot
    auto host_prep = [&]() {
        alloc_results(j);
        j->targets.assign(R * 4, Target());
        for (int64_t r = 0; r < R; ++r)
            for (int k = 0; k < 4; ++k) { j->targets[r * 4 + k].lo = 0; j->targets[r * 4 + k].hi = j->vmax[series_of_row(j, r)]; }
    };

    if (!j->direct && !ctx->use_pilot) {
        host_prep();
        j->n_pilot = 0;
        j->pilot_stage = 3;      // nothing in flight: plan from the certain bounds
        return VX_OK;
    }
    // the default size balances the pilot's cost against the fold's, both PER RANK: with the range
    // sharded over `world` ranks (vx_run_bounds_comm) it is sized from n / world, a value every
    // rank computes alike, so the replicated pilot does not grow with the GPU count in weak scaling
    const int64_t np = j->direct ? n : std::min(n, j->cfg.pilot_samples > 0 ? j->cfg.pilot_samples
                                                     : default_pilot(j->cfg.q_lo, j->cfg.q_hi, std::max<int64_t>(n / j->world, 1)));
    j->n_pilot = np;
    if ((double)R * (double)np * 4.0 > 120e9) return fail(ctx, VX_ERR_NOMEM, "direct/pilot matrix too large");
    { void *p_ = nullptr; int rc_ = pool_get(ctx, SLOT_DUMP, sizeof(int32_t) * R * np, &p_); if (rc_) return rc_; j->d_dump = (int32_t *)p_; }
    rc = launch_dump(ctx, j, 0, np, j->d_params, j->d_dump, np, ps);
    if (rc) return rc;

    j->pilot_ranks.assign(4, 0);
    std::vector<int64_t> &ranks = j->pilot_ranks;
    if (j->direct) {
        double g;
        vx_host_quantile_index(n, j->cfg.q_lo, &ranks[0], &ranks[1], &g);
        vx_host_quantile_index(n, j->cfg.q_hi, &ranks[2], &ranks[3], &g);
    } else {
        // pilot order statistics bracketing each wanted quantile by +-k sigma of its rank
        const double qs[2] = {j->cfg.q_lo, j->cfg.q_hi};
        for (int q = 0; q < 2; ++q) {
            const double sd = std::sqrt(std::max(qs[q] * (1 - qs[q]), 0.0));
            const double delta = ctx->pilot_sigmas * sd * (1.0 / std::sqrt((double)np) + 1.0 / std::sqrt((double)n)) + 2.0 / (double)np;
            const double ra = std::floor((double)(np - 1) * (qs[q] - delta)) - 1.0;
            const double rb = std::ceil((double)(np - 1) * (qs[q] + delta)) + 1.0;
            ranks[2 * q] = (int64_t)std::min(std::max(ra, 0.0), (double)(np - 1));
            ranks[2 * q + 1] = (int64_t)std::min(std::max(rb, 0.0), (double)(np - 1));
        }
    }
    rc = select_launch(ctx, j->d_dump, R, np, np, ranks, ps);
    if (rc) return rc;
    CK(cudaEventRecord(j->evB, ps));
    // Everything below is queued BEHIND the pilot (host order): its few CTAs must be resident before
    // the side stream's launches fill the SMs.
    if (!j->direct && !(j->cfg.max_running_time > 0.0)) {
        // the part of the shard the stepper will fold (the pilot's own samples come from its matrix)
        int64_t lo = std::max(j->shard_first, np);
        const int64_t hi = j->shard_first + j->shard_count;
        // (a) optionally its first `side` samples are stepped NOW by a DUMP launch on the
        //     (low-priority) side stream and folded from their matrix later (vx_fold_matrix_kernel):
        //     the pilot stepper is one serial T-day chain on < 1 CTA per SM, followed by a select and a
        //     host round trip, ~0.5 ms during which the GPU is > 85 % idle.  Measured on B200 (1e6
        //     samples, ms per job for 0 / 32768 / 65536 / 98304 / 131072 side samples): 4.00 / 4.01 /
        //     4.01 / 4.05 / 4.10 -- every co-resident warp slows the pilot's chain and delays its
        //     select by as much as the fold saves, also when the pilot's CTAs are packed onto SMs of
        //     their own (512-thread CTAs holding all the shared memory: 4.09 at 118784).  So the
        //     default is none; "side_samples" keeps the path for other shapes.
        int64_t side = ctx->side_samples >= 0 ? ctx->side_samples : 0;
        if (on_side) side = 0;
        side = std::min(side, hi - lo);
        if (side * 4 > hi - lo) side = (hi - lo >= 4096) ? (hi - lo) / 4 : 0;   // never more than a quarter of the shard
        side &= ~(int64_t)127;
        if ((double)R * (double)side * 4.0 > 16e9) side = 0;
        // (b) the rest is counting-sorted by pressure bucket (regime sort), first thing on the side
        //     stream: the fold needs the order as soon as the windows exist
        if (ctx->sort_samples && !j->explicit_params && hi - (lo + side) >= 4096 && hi - (lo + side) <= ((int64_t)1 << 30)) {
            rc = launch_regime_sort(ctx, j, lo + side, hi - (lo + side), ctx->side, &j->d_perm);
            if (rc) return rc;
            j->perm_first = lo + side; j->perm_count = hi - (lo + side);
            if (!on_side) CK(cudaEventRecord(ctx->ev_sorted, ctx->side));
        }
        if (side > 0) {
            void *p_ = nullptr;
            rc = pool_get(ctx, SLOT_SIDE, sizeof(int32_t) * R * side, &p_);
            if (rc) return rc;
            j->d_side = (int32_t *)p_;
            j->side_first = lo; j->side_count = side;
            rc = launch_dump(ctx, j, lo, side, j->d_params ? j->d_params + 6 * lo : nullptr, j->d_side, side, ctx->side,
                             nullptr, (size_t)ctx->side_pad_smem);
            if (rc) return rc;
        }
    }
    if (on_side) CK(cudaEventRecord(ctx->side_done, ctx->side));
    host_prep();
    j->pilot_stage = 2;
    return VX_OK;
}

static int pilot_complete(vx_ctx *ctx)
{
    Job *j = ctx->job;
    if (j->pilot_stage == 1) return VX_OK;           // expected-value job: run_expected did everything
    if (j->pilot_stage == 3) { j->pilot_stage = 0; return plan_windows(ctx, false); }
    j->pilot_stage = 0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, j->evA, j->evB) == cudaSuccess) j->pilot_ms += ms;
    const int64_t n = j->cfg.n_samples, R = j->R, np = j->n_pilot;
    std::vector<int64_t> ord;
    std::vector<long long> sums;
    select_collect(ctx, R, 4, ord, sums);
    if (j->direct) {
        fill_from_counts(j, ord, sums, n);
        j->n_total = n;
        j->finished = true;
        return VX_OK;
    }
    j->pilot_stats = ord;
    // a bracket that ran off either end of the pilot sample is open on that side
    const std::vector<int64_t> &ranks = j->pilot_ranks;
    for (int64_t r = 0; r < R; ++r)
        for (int q = 0; q < 2; ++q) {
            if (ranks[2 * q] == 0) j->pilot_stats[r * 4 + 2 * q] = 0;
            if (ranks[2 * q + 1] == np - 1) j->pilot_stats[r * 4 + 2 * q + 1] = j->vmax[series_of_row(j, r)];
        }
    return plan_windows(ctx, true);
}

int vx_job_pilot(vx_ctx *ctx)
{
    Phase ph_(ctx, "pilot (total)");
    int rc = pilot_launch(ctx, false);
    if (rc) return rc;
    mark(ctx, "pilot queued");
    CK(cudaStreamSynchronize(ctx->stream));
    mark(ctx, "pilot waited");
    rc = pilot_complete(ctx);
    mark(ctx, "planned");
    return rc;
}

int vx_job_acc(vx_ctx *ctx, void **device_ptr, int64_t *n_words)
{
    if (!ctx || !ctx->job || !device_ptr || !n_words) return fail(ctx, VX_ERR_STATE, "no job");
    Job *j = ctx->job;
    *device_ptr = j->direct ? nullptr : (void *)j->d_acc;
    *n_words = j->direct ? 0 : j->acc_words;
    return VX_OK;
}

// One streaming pass over this rank's shard.  `sync_end`: the caller merges the slab itself
// (staged API + an external all-reduce), so the stream must be idle on return; the whole-job
// entry points pass false and the host does not wait before vx_job_finalize's own copy-back.
static int accumulate_impl(vx_ctx *ctx, bool sync_end)
{
    if (!ctx || !ctx->job) return fail(ctx, VX_ERR_STATE, "no job");
    Job *j = ctx->job;
    if (!j->piloted) return fail(ctx, VX_ERR_STATE, "vx_job_pilot must precede vx_job_accumulate");
    if (!j->stats_set && !j->lstats_set && !j->stats_on_device)
        return fail(ctx, VX_ERR_STATE, "vx_job_param_stats (or vx_job_set_param_stats) must precede vx_job_accumulate");
    if (j->aborted) return fail(ctx, VX_ERR_STATE, "job aborted by the rejection rule");
    if (j->direct || j->finished) return VX_OK;
    int rc = need_device(ctx);
    if (rc) return rc;
    Phase ph_(ctx, "accumulate (fold)");
    CK(cudaMemsetAsync(j->d_acc, 0, sizeof(unsigned long long) * j->acc_words, ctx->stream));
    if (!j->stats_set) {
        // no separate statistics collective: this shard's rejection count and fixed-point moments
        // of p_soft_no ride in header words 2..5 of the slab the ranks all-reduce anyway
        j->stats_in_slab = true;
        if (j->stats_on_device) {
            if (j->stats_on_side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stats, 0));
            vx_stats_to_header_kernel<<<1, 32, 0, ctx->stream>>>(j->d_stats3, (long long)j->shard_count, STATS_FIX, j->d_acc);
            CKL();
        } else {
            const unsigned long long hs[4] = {
                (unsigned long long)j->lstats[0], (unsigned long long)std::llround(j->lstats[1] * STATS_FIX),
                (unsigned long long)std::llround(j->lstats[2] * STATS_FIX), (unsigned long long)j->lstats[3]};
            CK(cudaMemcpyAsync(j->d_acc + 2, hs, sizeof hs, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    const bool first_pass = j->n_passes == 0;
    const int64_t todo = first_pass ? j->shard_count : j->shard_done;
    const bool timed = first_pass && j->cfg.max_running_time > 0.0;
    const int64_t chunk = timed ? ctx->chunk_samples : ((int64_t)1 << 30);
    // Samples that already exist as raw rows -- the pilot's share of this shard and the samples
    // stepped beside the pilot -- are folded from their matrices.  With a side stream of its own (a
    // whole job, not a grid sub-context whose side stream carries the pilot) that HBM-streaming work
    // runs BESIDE the issue-bound stepper below: the side stream has the lower priority, so its CTAs
    // fill the SMs the stepper's tail leaves idle.
    const bool matrix_on_side = j->pilot_stream != ctx->side && !timed;
    cudaStream_t ms = matrix_on_side ? ctx->side : ctx->stream;
    if (matrix_on_side) {
        CK(cudaEventRecord(ctx->ev_wins, ctx->stream));          // window table uploaded, slab cleared
        CK(cudaStreamWaitEvent(ctx->side, ctx->ev_wins, 0));
    } else if (j->d_side || j->d_perm) {
        CK(cudaEventRecord(ctx->side_done, ctx->side));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->side_done, 0));
    }
    const int64_t lo = j->shard_first, hi = j->shard_first + todo;
    const int64_t pil_hi = (j->d_dump && j->n_pilot > 0) ? std::min(hi, j->n_pilot) : lo;
    int64_t done = 0;
    auto fold_matrix = [&](const int32_t *d, int64_t ld, int64_t c0, int64_t c1) {
        const int64_t chunks = (c1 - c0 + FM_COLS - 1) / FM_COLS;
        const dim3 grid((unsigned)j->R, (unsigned)std::min<int64_t>(chunks, 4096));
        vx_fold_matrix_kernel<<<grid, FM_THREADS, 0, ms>>>(d, ld, c0, c1, j->d_wins, j->d_acc, (int)j->R, j->bins64);
    };
    if (pil_hi > lo) {
        fold_matrix(j->d_dump, j->n_pilot, lo, pil_hi);
        CKL();
        done = pil_hi - lo;
    }
    if (j->d_side && j->side_count > 0 && j->side_first == j->shard_first + done) {
        fold_matrix(j->d_side, j->side_count, 0, j->side_count);
        CKL();
        done += j->side_count;
    }
    if (matrix_on_side) {
        CK(cudaEventRecord(ctx->side_done, ctx->side));
        if (j->d_perm && first_pass) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_sorted, 0));
    }
    CK(cudaEventRecord(j->evA, ctx->stream));
    int64_t stepped = 0;
    auto out_of_time = [&]() {       // model.py:187-189, at chunk granularity; something is always folded first
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - j->t_begin).count() > j->cfg.max_running_time;
    };
    while (done < todo) {
        if (timed && done > 0 && out_of_time()) break;
        const int64_t cnt = std::min(chunk, todo - done);
        SimArgs a = make_args(j, j->shard_first + done, cnt,
                              j->d_params ? j->d_params + 6 * (j->shard_first + done) : nullptr);
        a.wins = j->d_wins;
        a.acc = j->d_acc;
        a.bins64 = j->bins64;
        if (j->d_perm && j->perm_first == a.first && j->perm_count == cnt) {
            a.perm = j->d_perm;
        } else if (ctx->sort_samples && !j->explicit_params && cnt >= 4096) {
            int32_t *d_perm = nullptr;
            rc = launch_regime_sort(ctx, j, a.first, cnt, ctx->stream, &d_perm);
            if (rc) return rc;
            a.perm = d_perm;
            j->d_perm = nullptr;      // the pool slot now holds this chunk's order
        }
        const int64_t grid = (cnt + SIM_THREADS - 1) / SIM_THREADS;
        if (ctx->exact_poisson) vx_sim_kernel<SIM_FOLD, 1><<<(unsigned)grid, SIM_THREADS, 0, ctx->stream>>>(a);
        else vx_sim_kernel<SIM_FOLD, 0><<<(unsigned)grid, SIM_THREADS, (size_t)ctx->fold_pad_smem, ctx->stream>>>(a);
        CKL();
        j->fold_launches++;
        done += cnt;
        stepped += cnt;
        if (timed) CK(cudaStreamSynchronize(ctx->stream));
    }
    if (first_pass) { j->shard_done = done; j->fold_samples = stepped; }
    j->n_passes++;
    CK(cudaEventRecord(j->evB, ctx->stream));
    j->fold_timing_pending = true;
    if (matrix_on_side) CK(cudaStreamWaitEvent(ctx->stream, ctx->side_done, 0));   // the matrices' share of the slab
    if (j->use_comm && ctx->comm) {
        // the path's one exchange step: all-reduce(sum) of the integer slab, issued on the stream the
        // fold ran on, so it starts the moment this rank's last fold CTA retires
        NcclApi &api = nccl_api();
        CK(cudaEventRecord(j->evM0, ctx->stream));
        const int nrc = api.AllReduce(j->d_acc, j->d_acc, (size_t)j->acc_words, NCCL_UINT64, NCCL_SUM, ctx->comm, ctx->stream);
        if (nrc != 0) return fail(ctx, VX_ERR_CUDA, "ncclAllReduce failed: %s", api.GetErrorString ? api.GetErrorString(nrc) : "?");
        CK(cudaEventRecord(j->evM1, ctx->stream));
        j->merge_timed = true;
    }
    if (sync_end) CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_job_accumulate(vx_ctx *ctx) { return accumulate_impl(ctx, true); }

// Finalisation in two halves: finalize_launch queues the rank extraction and the copy-back into
// pinned host memory (asynchronous, main stream); finalize_complete, after that stream has been
// synchronised, resolves the targets on the host (or re-plans the windows: VX_REFINE).
static int finalize_launch(vx_ctx *ctx)
{
    if (!ctx || !ctx->job) return fail(ctx, VX_ERR_STATE, "no job");
    Job *j = ctx->job;
    int rc = need_device(ctx);
    if (rc) return rc;
    if (!j->piloted) return fail(ctx, VX_ERR_STATE, "vx_job_pilot must precede vx_job_finalize");
    if (!j->stats_set && !j->lstats_set && !j->stats_on_device)
        return fail(ctx, VX_ERR_STATE, "vx_job_param_stats (or vx_job_set_param_stats) must precede vx_job_finalize");
    const int64_t R = j->R;
    void *h = nullptr;
    rc = hpin_get(ctx, HSLOT_FIN, sizeof(long long) * (R * 5 + ACC_HDR + 4), &h);
    if (rc) return rc;
    long long *hp = (long long *)h;           // [R*4] extract results | [R] sums | [ACC_HDR] header | [3] direct-job stats
    if (!j->finished) {
        const int64_t warps = 2 * R, grid = (warps * 32 + 255) / 256;
        vx_extract_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(j->d_acc, j->d_wins, j->d_tgt, j->d_res, (int)R, j->cfg.q_lo, j->cfg.q_hi, j->bins64);
        CKL();
        CK(cudaMemcpyAsync(hp, j->d_res, sizeof(long long) * R * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hp + R * 4, j->d_acc, sizeof(long long) * (ACC_HDR + R), cudaMemcpyDeviceToHost, ctx->stream));
    } else if (!j->stats_set && !j->lstats_set && j->stats_on_device) {
        if (j->stats_on_side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stats, 0));
        CK(cudaMemcpyAsync(hp + R * 5 + ACC_HDR, j->d_stats3, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
        j->stats_fetch_pending = true;
    }
    CK(cudaEventRecord(j->ev1, ctx->stream));
    return VX_OK;
}

static int finalize_complete(vx_ctx *ctx, vx_result *res)
{
    Job *j = ctx->job;
    int rc;
    const int64_t R = j->R;
    const long long *hp = (const long long *)ctx->hpin[HSLOT_FIN];
    if (j->stats_fetch_pending) {
        memcpy(j->lstats, hp + R * 5 + ACC_HDR, sizeof(double) * 3);
        j->lstats[3] = (double)j->shard_count;
        j->lstats_set = true;
        j->stats_fetch_pending = false;
    }
    if (!j->stats_set && j->direct) {       // a direct job without a statistics collective: local == global
        memcpy(j->gstats, j->lstats, sizeof j->gstats);
        j->stats_set = true;
        j->aborted = j->gstats[0] > 10.0 * (double)j->cfg.n_samples;
    }
    if (!j->finished) {
        const long long *hres = hp;
        const unsigned long long *hdr = (const unsigned long long *)(hp + R * 4);
        std::vector<long long> sums(hp + R * 4 + ACC_HDR, hp + R * 5 + ACC_HDR);
        if (j->fold_timing_pending) {
            float fms = 0.f;
            if (cudaEventElapsedTime(&fms, j->evA, j->evB) == cudaSuccess) j->fold_ms += fms;
            if (j->merge_timed && cudaEventElapsedTime(&fms, j->evM0, j->evM1) == cudaSuccess) j->merge_ms += fms;
            j->fold_timing_pending = j->merge_timed = false;
        }
        mark(ctx, "fc events");
        const int64_t n = (int64_t)hdr[0];
        if (n < 1) return fail(ctx, VX_ERR_STATE, "no samples were folded");
        if (j->stats_in_slab) {
            j->gstats[0] = (double)hdr[2]; j->gstats[1] = (double)hdr[3] / STATS_FIX;
            j->gstats[2] = (double)hdr[4] / STATS_FIX; j->gstats[3] = (double)hdr[5];
            j->aborted = j->gstats[0] > 10.0 * (double)j->cfg.n_samples;      // model.py:282-287
            if (j->aborted) {
                res->device_ms = res->fold_ms = res->pilot_ms = res->merge_ms = 0; res->fold_launches = 0;
                res->n_pilot = res->fold_samples = res->side_samples = 0;
                j->n_total = 0;
                fill_result_scalars(ctx, res);
                return VX_OK;
            }
        }
        if (!j->ranks_set) {
            j->n_total = n;
            int64_t p0, n0, p1, n1; double g;
            vx_host_quantile_index(n, j->cfg.q_lo, &p0, &n0, &g);      // what the kernel derived from acc[0]
            vx_host_quantile_index(n, j->cfg.q_hi, &p1, &n1, &g);
            for (int64_t r = 0; r < R; ++r) {
                Target *t = &j->targets[r * 4];
                t[0].rank = p0; t[1].rank = n0; t[2].rank = p1; t[3].rank = n1;
            }
            j->ranks_set = true;
        } else if (n != j->n_total) {
            return fail(ctx, VX_ERR_STATE, "pass folded %lld samples, first pass folded %lld", (long long)n, (long long)j->n_total);
        }
        bool all = true;
        for (int64_t i = 0; i < R * 2; ++i) {
            const WinPlan &wp = j->plan[i];
            for (int k = 0; k < 2; ++k) {
                if (wp.tgt[k] < 0) continue;
                Target &t = j->targets[(i / 2) * 4 + wp.tgt[k]];
                const long long rr = hres[i * 2 + k];
                const int status = (int)(rr >> 32);
                const int64_t bin = (int64_t)(rr & 0xffffffffll);
                const int64_t a = wp.w.a, len = wp.w.len;
                if (status == 1) t.hi = std::min(t.hi, a - 1);
                else if (status == 2) t.lo = std::max(t.lo, a + len);
                else {
                    t.lo = std::max(t.lo, a + (bin << wp.w.shift));
                    t.hi = std::min(t.hi, a + ((bin + 1) << wp.w.shift) - 1);
                }
                if (t.lo > t.hi) return fail(ctx, VX_ERR_STATE, "window bookkeeping lost rank %lld of row %lld", (long long)t.rank, (long long)(i / 2));
                if (t.lo == t.hi) { t.resolved = true; t.value = t.lo; }
            }
        }
        mark(ctx, "fc targets");
        for (const Target &t : j->targets) all = all && t.resolved;
        if (!all) {
            rc = plan_windows(ctx, false);
            if (rc) return rc;
            return VX_REFINE;
        }
        std::vector<int64_t> ord(R * 4);
        for (int64_t i = 0; i < R * 4; ++i) ord[i] = j->targets[i].value;
        fill_from_counts(j, ord, sums, n);
        j->finished = true;
        mark(ctx, "fc fill");
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, j->ev0, j->ev1);
    res->device_ms = ms;
    res->fold_ms = j->fold_ms;
    res->pilot_ms = j->pilot_ms;
    res->fold_launches = j->fold_launches;
    res->merge_ms = j->merge_ms;
    res->n_pilot = j->n_pilot;
    res->fold_samples = j->fold_samples;
    res->side_samples = j->side_count;
    fill_result_scalars(ctx, res);
    return copy_results(ctx, res);
}

int vx_job_finalize(vx_ctx *ctx, vx_result *res)
{
    if (!res) return fail(ctx, VX_ERR_STATE, "no job");
    Phase ph_(ctx, "finalize");
    int rc = finalize_launch(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    mark(ctx, "fold waited");
    return finalize_complete(ctx, res);
}

int vx_job_end(vx_ctx *ctx)
{
    if (!ctx) return VX_ERR_ARG;
    free_job(ctx);
    return VX_OK;
}

// ---- whole jobs ------------------------------------------------------------------------
static int run_job(vx_ctx *ctx, vx_result *res, int32_t *counts_out)
{
    Job *j = ctx->job;
    int rc;
    bool stats_async = false;
    if (!j->explicit_params) {
        // Rejection statistics of the parameter draw.  When most of the (p_pro, p_anti) box is
        // acceptable the 10*n_rep abort cannot trigger and nothing downstream waits for the
        // numbers: the kernels run asynchronously and their result travels in the slab header
        // (streamed job) or is read with the final copy-back (direct job).  Otherwise (a box that is
        // mostly or wholly infeasible) they are read first, as the reference does, so that a job that
        // is going to abort does not run its pilot with millions of rejection attempts per sample.
        if (acceptance_probability(j->cfg.bounds) >= 0.25 && j->shard_count > 0) {
            stats_async = true;           // queued behind the pilot below
        } else {
            double st[4];
            rc = vx_job_param_stats(ctx, st);
            if (rc) return rc;
            // a sharded job decides on the MERGED count (slab header) unless no pair is acceptable at
            // all, which every rank sees alike; a single-GPU job knows the global count already
            const bool certain = st[0] > 10.0 * (double)j->cfg.n_samples &&
                                 (!j->use_comm || acceptance_probability(j->cfg.bounds) == 0.0);
            if ((!j->use_comm || j->direct) ? vx_job_set_param_stats(ctx, st) == 1 : certain) {
                j->aborted = true;
                memcpy(j->gstats, st, sizeof st);
                res->device_ms = 0; res->fold_ms = 0; res->pilot_ms = 0; res->merge_ms = 0; res->fold_launches = 0;
                res->n_pilot = res->fold_samples = res->side_samples = 0;
                fill_result_scalars(ctx, res);
                res->n_finished = 0;
                return VX_OK;
            }
        }
    }
    {
        Phase ph_(ctx, "pilot (total)");
        rc = pilot_launch(ctx, false);
        if (rc) return rc;
        if (stats_async) {
            // on the side stream and behind the pilot in host order: nothing on the way to the
            // windows waits for these kernels
            rc = launch_param_stats(ctx, j, ctx->side);
            if (rc) return rc;
            CK(cudaEventRecord(ctx->ev_stats, ctx->side));
            j->stats_on_device = true;
            j->stats_on_side = true;
        }
        mark(ctx, "pilot queued");
        CK(cudaStreamSynchronize(ctx->stream));
        mark(ctx, "pilot waited");
        rc = pilot_complete(ctx);
        mark(ctx, "planned");
        if (rc) return rc;
    }
    if (counts_out) {
        if (!j->direct || !j->d_dump) return fail(ctx, VX_ERR_ARG, "counts_out needs a direct job (n_samples <= direct_max)");
        CK(cudaMemcpyAsync(counts_out, j->d_dump, sizeof(int32_t) * j->R * j->cfg.n_samples, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (int it = 0; it < 64; ++it) {
        rc = accumulate_impl(ctx, false);
        if (rc) return rc;
        mark(ctx, "fold queued");
        rc = vx_job_finalize(ctx, res);
        mark(ctx, "finalized");
        if (rc != VX_REFINE) return rc;
    }
    return fail(ctx, VX_ERR_STATE, "window refinement did not converge");
}

int vx_run_bounds(vx_ctx *ctx, const vx_config *cfg, vx_result *res)
{
    if (!ctx || !res) return VX_ERR_ARG;
    mark(ctx, "enter");
    int rc = job_begin_common(ctx, cfg, 0, cfg ? cfg->n_samples : 0, nullptr);
    mark(ctx, "begin");
    if (rc == VX_OK) rc = run_job(ctx, res, nullptr);
    const std::string keep = ctx->err;
    free_job(ctx);
    ctx->err = keep;
    mark(ctx, "free");
    marks_flush(ctx);
    return rc;
}

// ---- a batch of independent scenarios on one GPU (BASELINE configs[4]) -------------------------
// Every box is a whole run_sampling job of its own (model.py:139-228 over its own parameter box).
// Run one after the other each would pay its latency-bound pilot (~0.4 ms with the GPU > 85 %
// idle) and three host round trips; here the boxes are software-pipelined over sub-contexts
// (own streams, workspace and job state): while the boxes of one wave fold -- the GPU-saturating
// part -- the next wave's pilots (stepper + select) run on high-priority streams in their shadow,
// and the host only ever waits for work that was queued two waves ago.
namespace {

int grid_stage_a(vx_ctx *sub, const vx_config *cfg, vx_result *res, bool *skip)
{
    *skip = false;
    int rc = job_begin_common(sub, cfg, 0, cfg->n_samples, nullptr);
    if (rc) return rc;
    Job *j = sub->job;
    if (acceptance_probability(cfg->bounds) >= 0.25) {
        rc = launch_param_stats(sub, j, sub->side);
        if (rc) return rc;
        j->stats_on_device = true;
    } else {                                   // a mostly infeasible box: read the count first (rare)
        double st[4];
        rc = vx_job_param_stats(sub, st);
        if (rc) return rc;
        if (vx_job_set_param_stats(sub, st) == 1) {
            res->device_ms = res->fold_ms = res->pilot_ms = res->merge_ms = 0; res->fold_launches = 0;
            res->n_pilot = res->fold_samples = res->side_samples = 0;
            fill_result_scalars(sub, res);
            res->n_finished = 0;
            *skip = true;
            return VX_OK;
        }
    }
    return pilot_launch(sub, true);
}

int grid_stage_b(vx_ctx *sub)
{
    vx_ctx *ctx = sub;
    CK(cudaEventSynchronize(sub->side_done));
    int rc = pilot_complete(sub);
    if (rc) return rc;
    rc = accumulate_impl(sub, false);
    if (rc) return rc;
    return finalize_launch(sub);
}

int grid_stage_c(vx_ctx *sub, vx_result *res)
{
    vx_ctx *ctx = sub;
    CK(cudaStreamSynchronize(sub->stream));
    int rc = finalize_complete(sub, res);
    for (int it = 0; rc == VX_REFINE && it < 64; ++it) {      // rare: a rank fell outside its window
        rc = accumulate_impl(sub, false);
        if (rc) return rc;
        rc = vx_job_finalize(sub, res);
    }
    return rc == VX_REFINE ? fail(sub, VX_ERR_STATE, "window refinement did not converge") : rc;
}

} // namespace

int vx_run_grid(vx_ctx *ctx, const vx_config *cfgs, int32_t n_boxes, vx_result *results, int32_t wave)
{
    if (!ctx || !cfgs || !results || n_boxes < 0) return fail(ctx, VX_ERR_ARG, "bad grid arguments");
    if (ctx->job) return fail(ctx, VX_ERR_STATE, "vx_run_grid while a job is open");
    int rc = need_device(ctx);
    if (rc) return rc;
    const int G = wave > 0 ? std::min(wave, 64) : 8;
    while ((int)ctx->subs.size() < 2 * G) {
        vx_ctx *sub = nullptr;
        rc = create_ctx(&sub, ctx->device, true);
        if (rc) return fail(ctx, rc, "vx_run_grid: %s", g_create_error.c_str());
        ctx->subs.push_back(sub);
    }
    for (vx_ctx *sub : ctx->subs) {
        sub->max_bins = ctx->max_bins; sub->chunk_samples = ctx->chunk_samples; sub->pilot_sigmas = ctx->pilot_sigmas;
        sub->use_pilot = ctx->use_pilot; sub->sort_samples = ctx->sort_samples; sub->exact_poisson = ctx->exact_poisson;
    }
    const int n_waves = (n_boxes + G - 1) / G;
    std::vector<char> skip(n_boxes, 0);
    auto sub_of = [&](int b) { return ctx->subs[((b / G) & 1) * G + (b % G)]; };
    auto fail_from = [&](vx_ctx *sub, int code) {
        ctx->err = sub->err;
        for (vx_ctx *x : ctx->subs) { cudaStreamSynchronize(x->stream); cudaStreamSynchronize(x->side); free_job(x); }
        return code;
    };
    auto stage_a = [&](int w) {
        for (int b = w * G; b < std::min(n_boxes, (w + 1) * G); ++b) {
            bool sk = false;
            const int r = grid_stage_a(sub_of(b), &cfgs[b], &results[b], &sk);
            if (r) return fail_from(sub_of(b), r);
            skip[b] = sk;
            if (sk) free_job(sub_of(b));
        }
        return (int)VX_OK;
    };
    auto stage_b = [&](int w) {
        for (int b = w * G; b < std::min(n_boxes, (w + 1) * G); ++b) {
            if (skip[b]) continue;
            const int r = grid_stage_b(sub_of(b));
            if (r) return fail_from(sub_of(b), r);
        }
        return (int)VX_OK;
    };
    auto stage_c = [&](int w) {
        for (int b = w * G; b < std::min(n_boxes, (w + 1) * G); ++b) {
            if (skip[b]) continue;
            vx_ctx *sub = sub_of(b);
            const int r = grid_stage_c(sub, &results[b]);
            if (r) return fail_from(sub, r);
            free_job(sub);
        }
        return (int)VX_OK;
    };
    const int64_t l0 = [&] { int64_t t = 0; for (vx_ctx *x : ctx->subs) t += x->launches; return t; }();
    if (n_waves > 0 && (rc = stage_a(0))) return rc;
    for (int w = 0; w < n_waves; ++w) {
        if ((rc = stage_b(w))) return rc;                        // wave w: pilots done -> folds queued
        if (w >= 1 && (rc = stage_c(w - 1))) return rc;          // wave w-1: results (its sub-contexts become free)
        if (w + 1 < n_waves && (rc = stage_a(w + 1))) return rc; // wave w+1: pilots, in the shadow of wave w's folds
    }
    if (n_waves > 0 && (rc = stage_c(n_waves - 1))) return rc;
    int64_t l1 = 0;
    for (vx_ctx *x : ctx->subs) l1 += x->launches;
    ctx->launches += l1 - l0;
    return VX_OK;
}

// ---- multi-GPU: one rank of a job sharded over the ranks of the context's communicator -------
int vx_comm_unique_id(vx_nccl_id *out)
{
    if (!out) return VX_ERR_ARG;
    NcclApi &api = nccl_api();
    if (!api.err.empty()) return fail(nullptr, VX_ERR_STATE, "NCCL unavailable: %s", api.err.c_str());
    const int nrc = api.GetUniqueId(out);
    if (nrc != 0) return fail(nullptr, VX_ERR_CUDA, "ncclGetUniqueId failed: %s", api.GetErrorString(nrc));
    return VX_OK;
}

int vx_comm_init(vx_ctx *ctx, const vx_nccl_id *id, int rank, int world)
{
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, VX_ERR_ARG, "bad communicator arguments");
    int rc = need_device(ctx);
    if (rc) return rc;
    if (ctx->comm) return fail(ctx, VX_ERR_STATE, "the context already has a communicator");
    NcclApi &api = nccl_api();
    if (!api.err.empty()) return fail(ctx, VX_ERR_STATE, "NCCL unavailable: %s", api.err.c_str());
    void *comm = nullptr;
    const int nrc = api.CommInitRank(&comm, world, *id, rank);
    if (nrc != 0) return fail(ctx, VX_ERR_CUDA, "ncclCommInitRank failed: %s", api.GetErrorString(nrc));
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
    return VX_OK;
}

int vx_comm_destroy(vx_ctx *ctx)
{
    if (!ctx) return VX_ERR_ARG;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl_api().CommDestroy(ctx->comm);
    }
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_world = 1;
    return VX_OK;
}

int vx_comm_info(vx_ctx *ctx, int *rank, int *world, int *nccl_version)
{
    if (!ctx) return VX_ERR_ARG;
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->comm ? ctx->comm_world : 1;
    if (nccl_version) { *nccl_version = 0; NcclApi &api = nccl_api(); if (api.err.empty() && api.GetVersion) api.GetVersion(nccl_version); }
    return VX_OK;
}

void vx_host_shard_range(int64_t n, int rank, int world, int64_t *first, int64_t *count)
{
    const int64_t base = n / world, rem = n % world;
    if (first) *first = (int64_t)rank * base + std::min<int64_t>(rank, rem);
    if (count) *count = base + (rank < rem ? 1 : 0);
}

int vx_run_bounds_comm(vx_ctx *ctx, const vx_config *cfg, vx_result *res)
{
    if (!ctx || !res || !cfg) return VX_ERR_ARG;
    if (!ctx->comm || ctx->comm_world == 1) return vx_run_bounds(ctx, cfg, res);
    const int64_t dmax = cfg->direct_max > 0 ? cfg->direct_max : VX_DIRECT_MAX_DEFAULT;
    const bool direct = cfg->mode == VX_MODE_EXPECTED || cfg->n_samples <= dmax;
    int64_t first = 0, count = cfg->n_samples;
    if (!direct) vx_host_shard_range(cfg->n_samples, ctx->comm_rank, ctx->comm_world, &first, &count);
    int rc = job_begin_common(ctx, cfg, first, count, nullptr);
    if (rc == VX_OK) {
        ctx->job->use_comm = !direct;     // a direct job is computed whole on every rank: no collective
        ctx->job->world = ctx->comm_world;
        rc = run_job(ctx, res, nullptr);
    }
    const std::string keep = ctx->err;
    free_job(ctx);
    ctx->err = keep;
    return rc;
}

int vx_run_params(vx_ctx *ctx, const vx_config *cfg, const double *params, vx_result *res,
                  int32_t *counts_out)
{
    if (!ctx || !res) return VX_ERR_ARG;
    if (!params) return fail(ctx, VX_ERR_ARG, "params is NULL");
    int rc = job_begin_common(ctx, cfg, 0, cfg ? cfg->n_samples : 0, params);
    if (rc == VX_OK && counts_out && cfg->mode != VX_MODE_STOCHASTIC)
        rc = fail(ctx, VX_ERR_ARG, "counts_out is only available in stochastic mode");
    if (rc == VX_OK) rc = run_job(ctx, res, counts_out);
    const std::string keep = ctx->err;
    free_job(ctx);
    ctx->err = keep;
    return rc;
}

// ---- pieces ----------------------------------------------------------------------------
int vx_sample_params(vx_ctx *ctx, const vx_config *cfg, int64_t first, int64_t count,
                     double *params_out, double *p_soft_no_out)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    rc = validate(ctx, cfg, false, nullptr);
    if (rc) return rc;
    if (count < 0 || !params_out) return fail(ctx, VX_ERR_ARG, "bad output");
    if (count == 0) return VX_OK;
    double *d_p = nullptr, *d_s = nullptr;
    { void *p_ = nullptr; int rc_;
      rc_ = pool_get(ctx, SLOT_TMP_A, sizeof(double) * 6 * count, &p_); if (rc_) return rc_; d_p = (double *)p_;
      if (p_soft_no_out) { rc_ = pool_get(ctx, SLOT_TMP_B, sizeof(double) * count, &p_); if (rc_) return rc_; d_s = (double *)p_; } }
    Bounds bd;
    memcpy(bd.b, cfg->bounds, sizeof bd.b);
    vx_sample_params_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(bd, cfg->seed, first, count, d_p, d_s);
    CKL();
    CK(cudaMemcpyAsync(params_out, d_p, sizeof(double) * 6 * count, cudaMemcpyDeviceToHost, ctx->stream));
    if (p_soft_no_out) CK(cudaMemcpyAsync(p_soft_no_out, d_s, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_dump_counts(vx_ctx *ctx, const vx_config *cfg, int64_t first, int64_t count,
                   int32_t *counts_out)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    rc = validate(ctx, cfg, false, nullptr);
    if (rc) return rc;
    if (count < 1 || !counts_out) return fail(ctx, VX_ERR_ARG, "bad output");
    Job tmp;
    tmp.cfg = *cfg; tmp.T = cfg->T; tmp.W = (cfg->T + 6) / 7; tmp.R = vx_host_rows(cfg->T);
    int32_t *d = nullptr;
    { void *p_ = nullptr; rc = pool_get(ctx, SLOT_DUMP, sizeof(int32_t) * tmp.R * count, &p_); if (rc) return rc; d = (int32_t *)p_; }
    rc = launch_dump(ctx, &tmp, first, count, nullptr, d, count);
    if (rc) return rc;
    CK(cudaMemcpyAsync(counts_out, d, sizeof(int32_t) * tmp.R * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_expected_series(vx_ctx *ctx, const vx_config *cfg, const double *params, int64_t n,
                       double *series_out)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    if (!cfg || !params || !series_out || n < 1) return fail(ctx, VX_ERR_ARG, "bad argument");
    vx_config c = *cfg;
    c.n_samples = n;
    c.mode = VX_MODE_EXPECTED;
    rc = validate(ctx, &c, true, params);
    if (rc) return rc;
    Job tmp;
    tmp.cfg = c; tmp.T = c.T; tmp.W = (c.T + 6) / 7;
    const int64_t RE = 4 * (int64_t)c.T + tmp.W, T = c.T;
    double *d_p = nullptr, *d_o = nullptr;
    { void *p_ = nullptr; int rc_;
      rc_ = pool_get(ctx, SLOT_PARAMS, sizeof(double) * 6 * n, &p_); if (rc_) return rc_; d_p = (double *)p_;
      rc_ = pool_get(ctx, SLOT_EXP, sizeof(double) * RE * n, &p_); if (rc_) return rc_; d_o = (double *)p_; }
    CK(cudaMemcpyAsync(d_p, params, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, ctx->stream));
    SimArgs a = make_args(&tmp, 0, n, d_p);
    vx_expected_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(a, d_o, n);
    CKL();
    std::vector<double> h(4 * T * n);
    CK(cudaMemcpyAsync(h.data(), d_o, sizeof(double) * 4 * T * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < 4; ++s)
        for (int64_t t = 0; t < T; ++t)
            for (int64_t i = 0; i < n; ++i)
                series_out[((int64_t)s * n + i) * T + t] = h[((int64_t)s * T + t) * n + i];
    return VX_OK;
}

// ---- self tests --------------------------------------------------------------------------
int vx_test_philox(vx_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    int rc = need_device(ctx);
    if (rc) return rc;
    uint32_t *d = nullptr;
    { void *p_ = nullptr; rc = pool_get(ctx, SLOT_TMP_A, 16, &p_); if (rc) return rc; d = (uint32_t *)p_; }
    vx_test_philox_kernel<<<1, 1, 0, ctx->stream>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
    CKL();
    CK(cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_test_poisson(vx_ctx *ctx, double lam, int64_t n, uint64_t seed, int32_t *out_host)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    if (n < 1 || !out_host) return fail(ctx, VX_ERR_ARG, "bad argument");
    int32_t *d = nullptr;
    { void *p_ = nullptr; rc = pool_get(ctx, SLOT_TMP_A, sizeof(int32_t) * n, &p_); if (rc) return rc; d = (int32_t *)p_; }
    vx_test_poisson_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((float)lam, n, seed, d, (int)ctx->exact_poisson);
    CKL();
    CK(cudaMemcpyAsync(out_host, d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_test_normal(vx_ctx *ctx, int64_t n, uint64_t seed, float *out_host)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    if (n < 1 || !out_host) return fail(ctx, VX_ERR_ARG, "bad argument");
    float *d = nullptr;
    { void *p_ = nullptr; rc = pool_get(ctx, SLOT_TMP_A, sizeof(float) * n, &p_); if (rc) return rc; d = (float *)p_; }
    const int64_t pairs = (n + 1) / 2;
    vx_test_normal_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, ctx->stream>>>(n, seed, d);
    CKL();
    CK(cudaMemcpyAsync(out_host, d, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_test_select(vx_ctx *ctx, const int32_t *data, int64_t rows, int64_t n, const int64_t *ranks,
                   int32_t n_ranks, int32_t *out, int64_t *sums)
{
    int rc = need_device(ctx);
    if (rc) return rc;
    if (!data || rows < 1 || n < 1 || !ranks || n_ranks < 1 || n_ranks > SEL_MAXR || !out)
        return fail(ctx, VX_ERR_ARG, "bad argument");
    for (int k = 0; k < n_ranks; ++k)
        if (ranks[k] < 0 || ranks[k] >= n) return fail(ctx, VX_ERR_ARG, "rank out of range");
    int32_t *d = nullptr;
    { void *p_ = nullptr; rc = pool_get(ctx, SLOT_DUMP, sizeof(int32_t) * rows * n, &p_); if (rc) return rc; d = (int32_t *)p_; }
    CK(cudaMemcpyAsync(d, data, sizeof(int32_t) * rows * n, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int64_t> rk(ranks, ranks + n_ranks), ord;
    std::vector<long long> sm;
    rc = run_select_i32(ctx, d, rows, n, n, rk, ord, sm);
    if (rc) return rc;
    for (int64_t i = 0; i < rows * n_ranks; ++i) out[i] = (int32_t)ord[i];
    if (sums) for (int64_t i = 0; i < rows; ++i) sums[i] = sm[i];
    return VX_OK;
}

} // extern "C"
