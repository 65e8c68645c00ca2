/*
 * pgmorl_b200.h -- C ABI of the B200-native PG-MORL hot path.
 *
 * The reference (albo437/PGMORL) is pure Python and has no FFI layer; the
 * "binding" a maintainer adds is a ctypes stub. Each entry
 * point below names the reference code whose arithmetic it replaces
 * (paths relative to the reference root; a2c = externals/pytorch-a2c-ppo-acktr-gail/a2c_ppo_acktr).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *     the caller owns all buffers, kernels never allocate;
 *   - arrays are dense, row-major, in the shapes written beside them;
 *   - P = tasks in the population shard, T = num_steps, N = envs per task,
 *     O/A/M = obs / action / objective dims, S = T*N samples per task,
 *     n_par = pgm_n_par(O,A,M) parameters per task in the reference's
 *     named_parameters() order (pgmorl_b200/layout.py);
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); calls are
 *     asynchronous w.r.t. the host unless stated;
 *   - return value 0 = success, otherwise an error code; pgm_last_error()
 *     gives the message of the last failure on the calling thread.
 *   - hidden width is fixed at 64 (a2c/model.py:202 default, warm_up.py:34-38).
 */
#ifndef PGMORL_B200_H
#define PGMORL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGM_ABI_VERSION 1
#define PGM_HIDDEN 64

enum {
    PGM_OK = 0,
    PGM_ERR_ARG = 1,      /* bad shape / null pointer / unsupported size */
    PGM_ERR_CUDA = 2,     /* a CUDA runtime call failed */
    PGM_ERR_WORKSPACE = 3 /* workspace too small */
};

/* action mode of pgm_policy_forward_f32 */
enum {
    PGM_ACT_SAMPLE = 0,        /* action = mean + exp(logstd)*eps   (a2c/model.py:64-65)  */
    PGM_ACT_DETERMINISTIC = 1, /* action = mean                      (a2c/model.py:61-62)  */
    PGM_ACT_EVALUATE = 2       /* action given, only log-prob        (a2c/model.py:75-82)  */
};

int pgm_abi_version(void);
const char *pgm_last_error(void);

/* number of parameters of one policy (actor 64-64 + critic 64-64 + heads + logstd) */
int pgm_n_par(int obs_dim, int act_dim, int obj_num);

/* ---------------------------------------------------------------------------
 * K1  population-batched actor-critic forward.
 * Replaces Policy.act / get_value / evaluate_actions (a2c/model.py:57-82),
 * MLPBase/MOMLPBase.forward (a2c/model.py:237-256), DiagGaussian + FixedNormal
 * (a2c/distributions.py:30-40,71-90) as called 2048x per iteration from
 * morl/mopg.py:103-135.
 *
 *   params  [P, n_par]
 *   obs     [P, rows_v, O]          value is produced for all rows_v rows
 *   eps     [P or 1, rows_a, A]     N(0,1) draws (SAMPLE mode); eps_shared!=0 => one copy for all tasks
 *   action  [P, rows_a, A]          out (SAMPLE/DETERMINISTIC) or in (EVALUATE); first rows_a rows of obs
 *   value   [P, rows_v, M]  out ;   logp [P, rows_a] out
 * Trajectory mode: rows_v=(T+1)*N, rows_a=T*N.  Per-step mode: rows_v=rows_a=N.
 * get_value: rows_a=0.
 */
int pgm_policy_forward_f32(const float *params, const float *obs, const float *eps, int eps_shared,
                           float *action, float *value, float *logp, int mode, int P, int rows_v,
                           int rows_a, int O, int A, int M, void *stream);

/* ---------------------------------------------------------------------------
 * K2  vector-reward GAE + scalarised, normalised advantage.
 * Replaces RolloutStorage.compute_returns, branch use_gae && use_proper_time_limits
 * (a2c/storage.py:83-94) and the PPO.update preamble (a2c/algo/ppo.py:43-56) with
 * WeightedSumScalarization.evaluate (morl/scalarization_methods.py:28-29).
 *
 *   rewards [P,T,N,M]  value [P,T+1,N,M]  masks,bad_masks [P,T+1,N]
 *   weights [P,M]  obj_var [P,M] or NULL (NULL => no un-normalisation)
 *   returns [P,T,N,M] out     adv [P,T,N] out (skipped when weights==NULL or adv==NULL)
 */
int pgm_gae_adv_f32(const float *rewards, const float *value, const float *masks,
                    const float *bad_masks, const float *weights, const float *obj_var,
                    float gamma, float lam, float *returns, float *adv, int P, int T, int N, int M,
                    void *stream);

/* ---------------------------------------------------------------------------
 * K3  PPO update: E epochs x B minibatches of clipped surrogate + clipped vector
 * value loss (- ecoef*entropy), backward, global grad-norm clip, Adam -- all tasks
 * in one launch, each task's chain of E*B dependent steps run by one thread-block
 * cluster. Replaces PPO.update (a2c/algo/ppo.py:62-115), feed_forward_generator
 * (a2c/storage.py:118-154), nn.utils.clip_grad_norm_ and torch.optim.Adam.step.
 *
 *   params, adam_m, adam_v [P, n_par]  in/out     adam_step [P] int32 in/out
 *   lr  [P] double (per-task learning rate of this iteration, a2c/utils.py:46-50)
 *   obs [P,S(+N),O] (first S rows used)   action [P,S,A]   logp_old [P,S]
 *   value_old [P,S(+N),M] (first S rows; task stride given)   returns [P,S,M]   adv [P,S]
 *   perm [P or 1, E, S] int32 : minibatch b of epoch e = perm[e, b*mb:(b+1)*mb], mb = S / B
 *   losses [P,3] out: mean over E*B updates of (value_loss, action_loss, entropy)
 *   workspace: >= pgm_ppo_workspace_bytes(...) bytes, 256-byte aligned
 *   cluster: CTAs per task (1,2,4,8,16) or 0 = choose from P and the SM count
 */
typedef struct {
    double clip_param;      /* 0.2  a2c/algo/ppo.py:83-84  */
    double value_loss_coef; /* 0.5  */
    double entropy_coef;    /* 0.0  */
    double max_grad_norm;   /* 0.5  */
    double beta1, beta2;    /* 0.9, 0.999 (torch.optim.Adam defaults) */
    double adam_eps;        /* 1e-5 morl/sample.py:31, warm_up.py:50 */
} pgm_ppo_hyper;

size_t pgm_ppo_workspace_bytes(int P, int S, int O, int A, int M, int cluster);

int pgm_ppo_update_f32(float *params, float *adam_m, float *adam_v, int32_t *adam_step,
                       const double *lr, const float *obs, size_t obs_task_stride,
                       const float *action, const float *logp_old, const float *value_old,
                       size_t value_task_stride, const float *returns, const float *adv,
                       const int32_t *perm, int perm_shared, int E, int B,
                       const pgm_ppo_hyper *hyper_host, float *losses, void *workspace,
                       size_t workspace_bytes, int cluster, int P, int S, int O, int A, int M,
                       void *stream);

/* Debug/verification aid: gradient of one minibatch (rows idx[0..mb)) of every task,
 * before clipping; grad [P, n_par], losses [P,3]. Same device code as pgm_ppo_update_f32. */
int pgm_ppo_grad_f32(const float *params, const float *obs, size_t obs_task_stride,
                     const float *action, const float *logp_old, const float *value_old,
                     size_t value_task_stride, const float *returns, const float *adv,
                     const int32_t *idx, int mb, const pgm_ppo_hyper *hyper_host, float *grad,
                     float *losses, void *workspace, size_t workspace_bytes, int cluster, int P,
                     int S, int O, int A, int M, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PGMORL_B200_H */
