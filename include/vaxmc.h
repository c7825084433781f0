/*
 * include/vaxmc.h -- C ABI of the B200-native Monte Carlo engine (libvaxmc.so).
 *
 * Drop-in boundary.  The reference (dfcastellanos/COVID19-Vaccination-Model) has no
 * FFI of its own; the narrowest seam shared by both of its callers (model.py:519-533
 * `main()` and app.py:458-473 `update_figures`) is
 *
 *     run_model(p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds,
 *               nv_max_bounds, CI, n_rep, N, date_range, max_running_time=None)
 *                                                              (model.py:306-368)
 *
 * and everything below it is the hot path this library replaces:
 *
 *     sample_param_combinations   model.py:231-303   -> vx_param_stats / vx_sample_params
 *     run_single_realization      model.py:29-136    -> the stepper kernels
 *     run_sampling                model.py:139-228   -> vx_run_bounds / vx_run_params
 *                                                       (stack, weekly series, quantile, mean)
 *
 * A binding (ctypes / cgo / JNI) needs nothing but this header: plain pointers and
 * sizes, caller-owned host buffers, no torch types.  The Python shim that restores the
 * reference's `run_model` signature lives in covid19-vaccination-model_b200/model.py;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns VX_OK (0), a positive status (VX_REFINE) or a negative error;
 *     vx_last_error(ctx) gives the text.  The library never frees caller memory.
 *   - one context per (process, GPU); calls on one context are serialised by the caller.
 *   - bounds are in MODEL units (fractions): the /100 of model.py:329-334 and the CI/100
 *     of model.py:358 are the caller's job (the Python shim does both).
 *   - series order everywhere is the reference's dict order (model.py:77-80):
 *       0 people_vaccinated_per_hundred        length T   (T = number of days)
 *       1 daily_vaccinations_per_million       length W   (W = ceil(T/7), model.py:196-214)
 *       2 cum_number_vac_received_per_hundred  length T
 *       3 vaccines_in_stock_per_hundred        length T
 *   - count rows ("raw rows", used by vx_run_params / vx_dump_counts): R = 2T + 2W rows
 *       [0,T)            n_vaccinated            per day
 *       [T,T+W)          delta_n_vacc[7k] + delta_n_vacc[7k-30]   per week
 *       [T+W,T+2W)       cum_number_vac_received per WEEK (constant within a week,
 *                        model.py:85-92: arrivals only on day%7==0)
 *       [T+2W,2T+2W)     vaccines_stock          per day
 */
#ifndef VAXMC_H
#define VAXMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_SERIES 4
#define VX_ABI_VERSION 2 /* bumped whenever vx_config / vx_result / a signature changes */

enum {
    VX_OK = 0,
    VX_REFINE = 1,       /* vx_job_finalize: windows were narrowed; accumulate again   */
    VX_ERR_ARG = -1,     /* invalid argument / unsupported size                        */
    VX_ERR_CUDA = -2,    /* CUDA runtime error (text in vx_last_error)                 */
    VX_ERR_STATE = -3,   /* call order violated                                        */
    VX_ERR_NOMEM = -4,   /* device memory                                              */
    VX_ERR_NODEVICE = -5 /* no CUDA device: this library has NO CPU fallback           */
};

enum { VX_MODE_STOCHASTIC = 0, /* model.py as written                                  */
       VX_MODE_EXPECTED = 1    /* np.random.poisson := identity (north_star check 1)   */ };

typedef struct vx_ctx vx_ctx;

/* Where materialise + select and pilot + fold cost the same on B200 (USA shape, measured:
 * 1.7 ms; a direct job is linear in n, 0.73 ms at 65536 samples). */
#define VX_DIRECT_MAX_DEFAULT 196608

typedef struct vx_config {
    double bounds[12];       /* pro lo,hi; anti lo,hi; pressure lo,hi; tau lo,hi; nv_0 lo,hi;
                                nv_max lo,hi  (model.py:231-239 argument order), fractions  */
    int64_t N;               /* population size (model.py:512: 50000)                       */
    int32_t T;               /* number of days = len(pd.date_range(start,end,"1d"))          */
    int32_t mode;            /* VX_MODE_*                                                    */
    int64_t n_samples;       /* n_rep: GLOBAL number of Monte Carlo samples                  */
    uint64_t seed;           /* Philox4x32-10 key                                            */
    double q_lo, q_hi;       /* quantile fractions (1-CI)/2 and (1+CI)/2  (model.py:220)     */
    int64_t direct_max;      /* n_samples <= this: materialise + exact select; above: pilot +
                                streaming fold.  0 = VX_DIRECT_MAX_DEFAULT                   */
    int64_t pilot_samples;   /* size of the pilot used to place the windows. 0 = default     */
    double max_running_time; /* seconds, <= 0: none (model.py:187-189).  Checked between chunks of
                                "chunk_samples" of the streaming fold (the first chunk always runs,
                                as the reference's first sample does); a direct job (n_samples <=
                                direct_max, < 2 ms of GPU time) and an expected-value job are ONE
                                chunk and therefore always complete.                             */
} vx_config;

typedef struct vx_result {
    double *mean[VX_SERIES]; /* caller-allocated, lengths T, W, T, T                          */
    double *lower[VX_SERIES];
    double *upper[VX_SERIES];
    int64_t n_finished;      /* model.py:226 number_finished_samples                          */
    int64_t n_rejected;      /* rejections in sample_param_combinations (model.py:280-281)    */
    double soft_no_mean;     /* mean and population std of p_soft_no = 1-(p_pro+p_anti),      */
    double soft_no_std;      /*   as fractions (model.py:339-342 multiplies by 100)           */
    int32_t aborted;         /* 1: the reference would return (None, None) (model.py:283-287) */
    int32_t n_passes;        /* streaming passes over the shard that were needed (>= 0)       */
    double device_ms;        /* device time of the job measured with CUDA events              */
    double fold_ms;          /* of which: streaming-fold stepper launches (all passes)        */
    double pilot_ms;         /* of which: pilot/direct stepper + exact select                 */
    int64_t fold_launches;   /* number of streaming-fold launches behind fold_ms              */
    double merge_ms;         /* of which: the NCCL all-reduce of the slab (vx_run_bounds_comm) */
    int64_t n_pilot;         /* samples of the replicated pilot (0: direct job)               */
    int64_t fold_samples;    /* samples this rank's fold launches stepped in the first pass   */
    int64_t side_samples;    /* samples stepped beside the pilot and folded from their matrix */
} vx_result;

/* ncclUniqueId, byte for byte (nccl.h: 128 bytes).  Made by ONE rank (vx_comm_unique_id), shipped
 * to the others by whatever rendezvous the host program has (torch.distributed, MPI, a file). */
typedef struct vx_nccl_id { char internal[128]; } vx_nccl_id;

/* ---- lifetime ------------------------------------------------------------------- */
int vx_create(vx_ctx **out, int device);
void vx_destroy(vx_ctx *ctx);
const char *vx_last_error(vx_ctx *ctx); /* ctx may be NULL: last error of vx_create  */
const char *vx_version(void);
/* ABI stamp for bindings that mirror the structs by hand (ctypes): VX_ABI_VERSION of the built
 * library and sizeof(vx_config) [0], sizeof(vx_result) [1], sizeof(vx_nccl_id) [2].          */
int vx_abi_version(void);
int64_t vx_sizeof(int which);
int vx_device_count(void);              /* 0 when no CUDA device is usable            */
/* tuning knobs (mainly for tests): "max_bins" (bins per window, default 262144),
 * "pilot_sigmas" (half-width of the pilot bracket in rank sigmas, default 6),
 * "use_pilot" (0: start from the coarse full-range windows), "sort_samples" (0: fold the
 * shard in id order instead of pressure-bucket order; same integers either way), "chunk_samples"
 * (granularity of the max_running_time check, default 2^20), "exact_poisson" (1: validation
 * mode, numpy's regimes on the Philox stream -- inversion below 10, PTRS above; 11-18x slower),
 * "side_samples" (samples stepped on a side stream while the pilot runs and folded from their
 * matrix; default 0 = none -- measured neutral on B200; same integers either way),
 * "select_global" (0: histogram select, rows staged on chip when they fit; 1: the 8-bit radix
 * select; 2: the histogram select streamed from global memory whatever the row length -- same
 * integers), "pipe_pilot" (0: the plain stepper also for launches that cannot fill the GPU).   */
int vx_set_option(vx_ctx *ctx, const char *name, double value);
/* The context keeps its device workspace (pilot matrix, accumulators) between jobs;
 * vx_trim returns it to the driver. */
int vx_trim(vx_ctx *ctx);

/* ---- whole job on one GPU (run_sampling, model.py:139-228, + parameter draw 231-303) */
int vx_run_bounds(vx_ctx *ctx, const vx_config *cfg, vx_result *res);

/* ---- the same job sharded over the GPUs of one box (run_sampling's sample loop, model.py:176-189,
 * split by global sample id; the per-date reduction of model.py:216-228 after the merge) ------
 * One process per GPU.  vx_comm_unique_id on one rank, the 128 bytes to every rank, then
 * vx_comm_init(ctx, id, rank, world) on each (collective: ncclCommInitRank).  vx_run_bounds_comm
 * is then called by EVERY rank with the same cfg: rank r steps the contiguous shard
 * vx_host_shard_range(n, r, world), the integer slab is merged by ONE ncclAllReduce(sum) per pass
 * issued by the library on its own stream right behind the fold, and every rank returns the same
 * (bit-identical, world-independent) result.  Jobs with n_samples <= direct_max run whole on each
 * rank without a collective.  libnccl.so.2 is dlopen'ed on first use ($VAXMC_NCCL_LIB overrides).
 * Without a communicator (or world == 1) vx_run_bounds_comm == vx_run_bounds.                  */
int vx_comm_unique_id(vx_nccl_id *out);
int vx_comm_init(vx_ctx *ctx, const vx_nccl_id *id, int rank, int world);
int vx_comm_destroy(vx_ctx *ctx);
int vx_comm_info(vx_ctx *ctx, int *rank, int *world, int *nccl_version);
int vx_run_bounds_comm(vx_ctx *ctx, const vx_config *cfg, vx_result *res);
void vx_host_shard_range(int64_t n, int rank, int world, int64_t *first, int64_t *count);

/* ---- a batch of independent scenarios on one GPU (BASELINE.json configs[4]) -------------------
 * cfgs[0..n_boxes): one whole job each (own bounds / seed / sizes), results[b] with caller-allocated
 * buffers like vx_run_bounds.  Equivalent to n_boxes calls of vx_run_bounds (same bytes out), but
 * pipelined: the boxes' latency-bound pilots run on high-priority streams in the shadow of other
 * boxes' folds, and the host never waits for a kernel it has just queued.  `wave` = boxes in
 * flight per pipeline stage (0: default 8; 2 x wave sub-contexts are kept by ctx until vx_trim).
 * Under torchrun the caller partitions the boxes over ranks; no collective is involved.         */
int vx_run_grid(vx_ctx *ctx, const vx_config *cfgs, int32_t n_boxes, vx_result *results, int32_t wave);

/* Same reduction over EXPLICIT parameter tuples, params = host [n][6] row-major in the
 * reference's tuple order (model.py:296-300).  cfg->bounds is ignored, cfg->n_samples = n.
 * counts_out (optional, may be NULL; stochastic mode only): host int32 [R][n], raw rows. */
int vx_run_params(vx_ctx *ctx, const vx_config *cfg, const double *params, vx_result *res,
                  int32_t *counts_out);

/* ---- pieces of the path, for tests and for sample_param_combinations ------------- */
/* Parameter tuples of samples [first, first+count) as drawn on the device from cfg->bounds
 * (model.py:275-300); params_out host [count][6]; p_soft_no_out host [count] or NULL.     */
int vx_sample_params(vx_ctx *ctx, const vx_config *cfg, int64_t first, int64_t count,
                     double *params_out, double *p_soft_no_out);
/* Raw integer rows of samples [first, first+count) in bounds mode: host int32 [R][count] --
 * the integers behind what run_single_realization records (model.py:124-127) and run_sampling
 * stacks (model.py:193), before the float scaling.                                          */
int vx_dump_counts(vx_ctx *ctx, const vx_config *cfg, int64_t first, int64_t count,
                   int32_t *counts_out);
/* Expected-value trajectories of explicit parameter tuples, host double [4][n][T] in the
 * reference's units (what run_single_realization returns with poisson := identity).       */
int vx_expected_series(vx_ctx *ctx, const vx_config *cfg, const double *params, int64_t n,
                       double *series_out);

/* ---- staged job: one shard of the global sample range per GPU (multi-GPU path) ----
 * The sample loop of run_sampling (model.py:176-189) split over ranks by global sample id;
 * the per-date reduction (model.py:216-228) happens after the integer slab is merged.
 *   vx_job_begin(cfg, shard_first, shard_count)
 *   vx_job_param_stats(stats[4])        this shard's rejection count / p_soft_no moments
 *   [vx_job_set_param_stats(stats[4])]  OPTIONAL: hand back stats summed over ranks by the
 *                                       caller (returns 1 = aborted, then skip to vx_job_end).
 *                                       Without it the shard's statistics travel in the header
 *                                       of the slab (fixed point) and are merged by the same
 *                                       all-reduce; vx_job_finalize then reports `aborted`.
 *   vx_job_pilot()                      replicated on every rank (same global sample ids)
 *   loop: vx_job_accumulate()           streaming fold of this rank's shard
 *         vx_job_acc(&dev_ptr,&n_words) -> all-reduce(sum) this int64 device slab (NCCL)
 *         vx_job_finalize(res)          VX_OK: done;  VX_REFINE: loop again
 *   vx_job_end()
 * When the job is "direct" (n_samples <= direct_max) the pilot already is the whole job,
 * vx_job_acc returns n_words = 0 and the loop body runs zero collectives.                 */
int vx_job_begin(vx_ctx *ctx, const vx_config *cfg, int64_t shard_first, int64_t shard_count);
int vx_job_param_stats(vx_ctx *ctx, double stats[4]); /* n_rej, sum s, sum s^2, n (shard) */
int vx_job_set_param_stats(vx_ctx *ctx, const double stats[4]);
int vx_job_pilot(vx_ctx *ctx);
int vx_job_is_direct(vx_ctx *ctx);
int vx_job_accumulate(vx_ctx *ctx);
int vx_job_acc(vx_ctx *ctx, void **device_ptr, int64_t *n_words);
int vx_job_finalize(vx_ctx *ctx, vx_result *res);
int vx_job_end(vx_ctx *ctx);

/* ---- host-only helpers (no CUDA calls; usable on a box without a GPU) ------------- */
/* np.quantile(..., method="linear") for one date given the two neighbouring order
 * statistics already converted to the reference's float units (model.py:220).           */
double vx_host_lerp(double x_prev, double x_next, double gamma);
/* virtual index (n-1)*q split into the previous index and gamma, as numpy does.         */
void vx_host_quantile_index(int64_t n, double q, int64_t *prev, int64_t *next, double *gamma);
/* integer count -> the float the reference records (model.py:124-127), series 0..3.     */
double vx_host_count_to_value(int series, int64_t count, int64_t N);
/* Number of weekly points and raw rows for T days. */
int32_t vx_host_weeks(int32_t T);
int64_t vx_host_rows(int32_t T);

/* ---- device self-tests of the building blocks (used by tests/, tiny launches) -----
 * vx_test_poisson draws with the stepper's own sampler (the np.random.poisson calls of
 * model.py:104,116); vx_test_select is the order-statistics step of np.quantile (model.py:220). */
int vx_test_philox(vx_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int vx_test_poisson(vx_ctx *ctx, double lam, int64_t n, uint64_t seed, int32_t *out_host);
int vx_test_normal(vx_ctx *ctx, int64_t n, uint64_t seed, float *out_host);
/* exact multi-rank select of each row of host int32 data [rows][n] (kernel per "select_global") */
int vx_test_select(vx_ctx *ctx, const int32_t *data, int64_t rows, int64_t n,
                   const int64_t *ranks, int32_t n_ranks, int32_t *out /*[rows][n_ranks]*/,
                   int64_t *sums /*[rows]*/);
/* number of kernels this context has launched since creation (bench.py gpu_launches)   */
int64_t vx_launch_count(vx_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* VAXMC_H */
