/*
 * mtraffic.h -- C ABI of libmtraffic, the B200 (sm_100a) engine for the
 * batched finite-volume step of the macroscopic traffic models.
 *
 * Drop-in boundary.  The reference exposes the step as Python methods of
 * its Model classes; each entry point below names the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   mt_create / mt_destroy   <- Model.__init__            mbrl_traffic/models/base.py:26-48
 *                               LWRModel.__init__         mbrl_traffic/models/lwr.py:56-186
 *                               ARZModel.__init__         mbrl_traffic/models/arz.py:60-199
 *   mt_set_params            <- the attribute writes the reference performs
 *                               between steps: fitness_fn lwr.py:161-164,
 *                               arz.py:169-173; update() lwr.py:322-325,
 *                               arz.py:498-502; load() lwr.py:371-379,
 *                               arz.py:549-558
 *   mt_step                  <- Model.get_next_obs(obs, action)
 *                               base.py:58-74, lwr.py:192-204, arz.py:205-221
 *   mt_rollout               <- the chained-step loop of
 *                               experiments/evaluate_model.py:398-403 and the
 *                               K-shoot rollout shape policies/kshoot.py:65-85
 *   mt_step_host             <- the same get_next_obs call on host (NumPy)
 *                               buffers, copies included
 *   mt_density_sqerr         <- Model.compute_loss(...)   lwr.py:329-343,
 *                               arz.py:506-520 (per-row squared density error;
 *                               the caller takes the mean)
 *
 * All pointers are plain addresses, sizes are plain integers; there are no
 * torch or CUDA types in the signatures (a stream is passed as `void*`
 * holding a cudaStream_t, NULL = the legacy default stream).
 *
 * Data layout.  An observation batch is a dense row-major (n_envs, 2*n_cells)
 * array of the engine dtype; row b is [rho_0..rho_{N-1} | v_0..v_{N-1}],
 * exactly the reference's per-env observation (lwr.py:197,204;
 * arz.py:210-211,221).  Input and output batches must not overlap.
 *
 * Every function returns MT_OK (0) or a negative mt_status; the message of
 * the last failure on the calling thread is returned by mt_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with MT_ECUDA.
 */
#ifndef MTRAFFIC_H_
#define MTRAFFIC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_VERSION 200 /* 0.2.0: mt_params grew (scheme, alpha), mt_aggregate's scratch
                          * holds one more word, new entry points (see below)       */

typedef struct mt_engine mt_engine;

typedef enum mt_status {
  MT_OK = 0,
  MT_EINVAL = -1,  /* bad argument (unknown model/bc/dtype, bad shape, NULL) */
  MT_ECUDA = -2,   /* CUDA runtime error or no device                        */
  MT_ENOMEM = -3,  /* host or device allocation failed                       */
  MT_ESTATE = -4   /* call order violated (e.g. step before set_params)      */
} mt_status;

typedef enum mt_model {
  MT_MODEL_LWR = 0,      /* lwr.py  */
  MT_MODEL_ARZ = 1,      /* arz.py  */
  MT_MODEL_NONLOCAL = 2  /* paper cited at lwr.py:209-211; no reference code */
} mt_model;

typedef enum mt_dtype { MT_F32 = 0, MT_F64 = 1 } mt_dtype;

/* Boundary rule of one road end.  The reference's three modes
 * (lwr.py:237-264, arz.py:316-364) are LOOP/LOOP, EXTEND/EXTEND and
 * CONSTANT/CONSTANT; mixed ends (e.g. CONSTANT inflow + EXTEND outflow) and
 * HALO (the cell is a ghost owned by a neighbouring rank and is left as is)
 * are additions used by the highway and domain-decomposed configurations. */
typedef enum mt_bc {
  MT_BC_LOOP = 0,
  MT_BC_EXTEND = 1,
  MT_BC_CONSTANT = 2,
  MT_BC_HALO = 3
} mt_bc;

/* Numerical scheme of the non-local model (the reference only names the
 * model, utils/train.py:685-689; lwr.py:209-211 cites the first paper):
 *   GODUNOV  mean downstream VELOCITY, upwind flux rho_j W_j
 *            (Friedrich, Kolb, Goettlich, NHM 13, 2018)
 *   LXF      speed of the mean downstream DENSITY, Lax-Friedrichs flux
 *            (Blandin, Goatin, Numer. Math. 132, 2016)                      */
typedef enum { MT_SCHEME_GODUNOV = 0, MT_SCHEME_LXF = 1 } mt_scheme;

enum {
  /* mt_params.flags */
  MT_FLAG_FAST_MATH = 1,  /* allow FMA contraction, approximate division and
                             x**lam as exp(lam log x) for a general lam
                             (not bit-comparable; throughput mode)         */
  MT_FLAG_NO_PIPELINE = 2,/* use the direct (one CTA per road tile) kernel
                             instead of the persistent TMA-pipelined one;
                             same results, for A/B measurements and tests  */
  MT_FLAG_INDEPENDENT_INPUT = 4,
                          /* the caller guarantees that the INPUT batch of a
                             step is not written by the work enqueued right
                             before it on the stream (independent batches, a
                             pool of buffers).  The step kernel then starts
                             fetching its input while the previous launch
                             still drains; its stores still wait for it.
                             Same results.  Do NOT set it for chained steps
                             (input = the previous step's output).  Used only
                             when this launch and the library's previous one
                             on the stream both fill the GPU (anything older
                             is then known to have finished), with scalar
                             parameters and without speed limits / rewards.
                             WITHOUT this flag the library decides by itself:
                             it remembers, per stream, the bytes its last
                             launch writes and lets the next step's loads
                             start early when its input lies outside them;
                             a chained step (input = previous output) waits
                             as before.                                    */
  MT_FLAG_DEPENDENT_INPUT = 8,
                          /* every step waits for the launch before it
                             (turns the automatic check above off; for A/B
                             measurements, or when another library enqueues
                             programmatic-dependent kernels that write this
                             engine's inputs between its steps)            */
  MT_FLAG_FSOLVE_ABORT = 16
                          /* ARZ: mimic the reference on roads whose
                             relaxation system it cannot solve.  With an
                             exact-zero density rhs is NaN (arz.py:411,
                             460), scipy's fsolve returns its initial guess
                             (arz.py:462-463) and y stays frozen for the
                             WHOLE road that step.  Off: the closed-form
                             root is used everywhere (NaN at the affected
                             cells only).  On: one more kernel per step
                             re-reads the input, finds such roads and
                             rewrites their speeds; multi-step launches are
                             split into single steps.                      */
};

/* Fixed for the life of an engine. */
typedef struct mt_config {
  int32_t model;        /* mt_model                                          */
  int32_t dtype;        /* mt_dtype                                          */
  int64_t n_envs;       /* B: independent roads in the batch                 */
  int64_t n_cells;      /* N: cells per road (observation row = 2N values)   */
  int32_t device;       /* CUDA device ordinal                               */
  int32_t max_n_eta;    /* non-local: largest look-ahead window in cells     */
  int32_t host_staging; /* !=0: allocate device staging for mt_step_host     */
  int32_t reserved;
} mt_config;

/* Model parameters; the names follow the reference attributes
 * (lwr.py:128-138, arz.py:134-145, arz.py:199).  A per-env pointer, when not
 * NULL, is a DEVICE array of n_envs values of the engine dtype and overrides
 * the scalar of the same name. */
typedef struct mt_params {
  double dx, dt;
  double rho_max, v_max, tau, lam;
  double rho_crit;            /* ARZ: frozen at construction (arz.py:199);
                                 <0 means rho_max/2.  Ignored by LWR, which
                                 recomputes rho_max/2 every call (lwr.py:282) */
  const void *rho_max_env, *v_max_env, *tau_env, *lam_env, *rho_crit_env;
  int32_t bc_left, bc_right;  /* mt_bc                                       */
  double bc_left_val[2];      /* CONSTANT: rho (LWR, non-local) or (rho, y)  */
  double bc_right_val[2];     /*           for ARZ (arz.py:347-360)          */
  const void *gamma;          /* non-local: DEVICE array of n_eta weights;
                                 the sums take them in ascending k, one
                                 fused multiply-add per term             */
  int32_t n_eta;
  int32_t flags;
  int32_t scheme;             /* non-local: mt_scheme                        */
  int32_t reserved;
  double alpha;               /* non-local, MT_SCHEME_LXF: numerical viscosity
                                 (>= max |V|); < 0 means each road's v_max   */
} mt_params;

int mt_version(void);
const char *mt_last_error(void);

int mt_create(const mt_config *cfg, mt_engine **out);
void mt_destroy(mt_engine *e);

/* Derive the per-env constant table on the device (asynchronous on stream).
 * With per-road arrays the call then waits for that kernel and reads 8 bytes
 * back (do all roads share an exponent kind? is every rho_max 1?) to pick a
 * step kernel without the run-time dispatch; not while the stream is being
 * captured, where the general kernel is kept. */
int mt_set_params(mt_engine *e, const mt_params *p, void *stream);

/* One step for the whole batch.  obs_in/obs_out: DEVICE (B, 2N).  action:
 * DEVICE [B] speed limits overriding v_max for this step, or NULL (the
 * reference discards the action, lwr.py:194, arz.py:207).  reward_out: DEVICE
 * [B] density-weighted mean speed of the new state, or NULL.  Asynchronous;
 * no allocation; no host synchronisation.
 * Launches of ONE engine must be ordered with respect to each other (same
 * stream, or events between streams): the engine owns device-side scheduling
 * state (the tile counter of large batches).  Use one engine per stream for
 * concurrent batches. */
int mt_step(mt_engine *e, const void *obs_in, const void *action,
            void *obs_out, void *reward_out, void *stream);

/* mt_step that also leaves a bfloat16 copy of the new state in obs_out_bf16
 * (DEVICE, (B, 2N) row-major, 8-byte aligned): what a policy network reads
 * next, written by the step kernel while the values are still in registers
 * instead of a separate conversion pass over obs_out.  (The rollout loop the
 * reference intends, policies/kshoot.py:65-85 / algorithm.py:564-620: policy
 * -> model step -> reward, every step.) */
int mt_step_obs16(mt_engine *e, const void *obs_in, const void *action,
                  void *obs_out, void *reward_out, void *obs_out_bf16,
                  void *stream);

/* Output layer of a feed-forward speed-limit policy fused with its squashing
 * (the actor shape of policies/sac.py:303-322 with the action range of
 * envs/vsl.py:92-100):  action[b] = scale * (tanh(hidden[b,:] . weight + bias)
 * + 1), hidden (B, n_hidden) / weight (n_hidden) / bias (1) bfloat16 DEVICE
 * arrays, action in `dtype` (mt_dtype) -- the buffer mt_step reads its speed
 * limits from.  reward_prev / return_sum (both or neither; DEVICE [B] in
 * `dtype`): return_sum[b] += reward_prev[b], the accumulation of a rollout's
 * return (algorithm.py:564-620 sums rewards per episode) riding on this launch
 * -- pass the reward buffer of the PREVIOUS step.  Asynchronous on stream. */
int mt_actor_head(int32_t device, int32_t dtype, const void *hidden_bf16,
                  const void *weight_bf16, const void *bias_bf16,
                  void *action_out, int64_t n_envs, int32_t n_hidden,
                  double scale, const void *reward_prev, void *return_sum,
                  void *stream);

/* n_steps chained steps ping-ponging between two caller-owned DEVICE batches,
 * starting from obs_a.  actions: DEVICE [n_steps, B] or NULL.  rewards:
 * DEVICE [n_steps, B] or NULL.  *result_in_b is set to 1 when the final state
 * is in obs_b, else 0; the other batch is scratch.  Batches whose roads fit
 * one tile run as ONE launch (state in registers between steps, one HBM read
 * and write per road); longer roads (without rewards, without the LOOP rule)
 * advance as overlapping windows, 16 steps per launch; results are
 * bit-identical to n_steps mt_step calls. */
int mt_rollout(mt_engine *e, void *obs_a, void *obs_b, const void *actions,
               void *rewards, int32_t n_steps, void *stream,
               int32_t *result_in_b);

/* get_next_obs on HOST buffers (pinned for full speed): copies obs (and the
 * action) to the device, runs n_steps chained steps, copies the final
 * observation (and the last step's rewards) back and synchronises.  Needs
 * cfg.host_staging. */
int mt_step_host(mt_engine *e, const void *obs_in_host,
                 const void *action_host, void *obs_out_host,
                 void *reward_out_host, int32_t n_steps);

/* out[b] = sum over the density half of (pred - target)^2, DEVICE pointers. */
int mt_density_sqerr(mt_engine *e, const void *obs_pred,
                     const void *obs_target, void *out_per_env, void *stream);

/* flags_out[b] = 1 if any density or speed of road b is inf or NaN, else 0
 * (DEVICE int32 [B]).  The reference has no such check: a diverged road just
 * propagates NaN through get_next_obs (arz.py:205-221); rollout loops use this
 * to reset or drop roads an erratic controller has driven unstable. */
int mt_nonfinite(mt_engine *e, const void *obs, int32_t *flags_out, void *stream);

/* Microscopic -> macroscopic aggregation (the reference's notebook loop,
 * experiments/plotting_results.ipynb cell 8): records (time, position, speed)
 * -> per (time row, section) mean speed and density.  All pointers are DEVICE
 * float64; records must be grouped by non-decreasing time row (MT_EINVAL
 * otherwise); scratch_idx holds 2*n_records + 1 int32.  Bit-equal to the
 * reference's sequential accumulation.  Synchronises the stream once (to
 * read the sortedness flag). */
int mt_aggregate(const void *time, const void *position, const void *speed,
                 int64_t n_records, double dt, double dx, int32_t n_rows,
                 int32_t n_sections, double empty_speed, void *scratch_idx,
                 void *speed_out, void *density_out, int32_t device,
                 void *stream);

/* ---- ghost-cell exchange of a domain-decomposed road over peer memory ----
 * (one road split over the GPUs of a node; the
 * reference steps a road as one NumPy vector, arz.py:205-221, and has no
 * multi-device path).  Each rank owns a mailbox in its own device memory,
 * mapped into its neighbours through CUDA IPC; mt_halo_push stores this rank's
 * first / last `halo` own cells into the neighbours' mailboxes and publishes
 * the exchange number, mt_halo_pull waits for both neighbours' numbers and
 * fills this rank's ghost cells.  local: DEVICE [2 * n_local] block
 * [rho | v] with `halo` ghost cells on every internal edge.  Both calls are
 * asynchronous on `stream`; every rank must issue the same sequence of
 * push / pull pairs. */
typedef struct mt_halo mt_halo;
enum { MT_IPC_HANDLE_BYTES = 64 };

int mt_halo_create(int32_t device, int32_t dtype, int32_t halo, mt_halo **out);
void mt_halo_destroy(mt_halo *h);
/* The IPC handle of this rank's mailbox, to be sent to the neighbours. */
int mt_halo_handle(const mt_halo *h, unsigned char handle_out[MT_IPC_HANDLE_BYTES]);
/* Map the neighbours' mailboxes (NULL: no neighbour on that side). */
int mt_halo_connect(mt_halo *h, const unsigned char *left_handle,
                    const unsigned char *right_handle);
/* Same-process variant (tests, several ranks emulated by one process). */
int mt_halo_connect_local(mt_halo *h, mt_halo *left, mt_halo *right);
int mt_halo_push(mt_halo *h, void *local, int64_t n_local, void *stream);
int mt_halo_pull(mt_halo *h, void *local, int64_t n_local, void *stream);
/* n_steps steps of a rank's block with the ghost-cell exchange folded into the
 * step launches: every `halo` steps are ONE launch that fetches the ghost
 * cells from this rank's mailbox (after waiting for the neighbours' flags,
 * unless ghosts_fresh says the block's ghost cells are current, as right after
 * a scatter or an explicit exchange), advances overlapping windows of the block
 * on chip, and stores the block's first / last `halo` own cells straight into
 * the neighbours' mailboxes with the epoch flag.  Equivalent to n_steps
 * mt_step calls with mt_halo_push / mt_halo_pull every `halo` steps; on return
 * the ghost cells are stale and one push is pending (the next call pulls it).
 * The engine holds one road block (n_envs == 1, MT_BC_HALO on every edge that
 * has a neighbour) too long for the one-tile kernels; halo * sizeof(dtype)
 * must be a multiple of 16.  Every rank issues the same sequence of calls. */
int mt_halo_advance(mt_engine *e, mt_halo *h, void *obs_a, void *obs_b,
                    int32_t n_steps, int32_t ghosts_fresh, void *stream,
                    int32_t *result_in_b);
/* Synchronises the device; *timed_out = 1 if a pull gave up waiting. */
int mt_halo_status(mt_halo *h, int32_t *timed_out);

/* Number of kernels this engine has launched so far. */
int mt_launch_count(const mt_engine *e, int64_t *out);

/* Page-locked host memory for mt_step_host callers. */
int mt_alloc_pinned(uint64_t bytes, void **out);
int mt_free_pinned(void *p);

#ifdef __cplusplus
}
#endif
#endif /* MTRAFFIC_H_ */
