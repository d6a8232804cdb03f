/*
 * openrl_b200.h — C-ABI of libopenrl_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for OpenRL's rollout-collection + PPO/MAPPO-update hot path.
 * The reference has no FFI: its seams are Python constructor-injection points
 * (SURVEY.md §8b).  Every entry point below replaces the inner loop of one
 * reference function; the Python host package `openrl_b200` binds them with ctypes
 * (openrl_b200/lib.py) and INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *     name ends in _host; the caller owns all memory (torch tensors in practice);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, otherwise a cudaError_t or ORL_ERR_*; the message is
 *     available from orl_last_error();
 *   - all floating-point buffers are float32, row-major, laid out like the
 *     reference's ReplayData arrays with the (env, agent) axes flattened:
 *     element (t, n, a, k) of a (T[+1], N, A, K) array is at ((t*N + n)*A + a)*K + k;
 *     B = N*A is the number of "rows" (columns of the time scan).
 */
#ifndef OPENRL_B200_H
#define OPENRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORL_ABI_VERSION 1

#define ORL_ERR_BAD_ARG 10001
#define ORL_ERR_UNSUPPORTED 10002

/* library / device --------------------------------------------------------------- */
int orl_abi_version(void);
const char* orl_last_error(void);
/* number of SMs of the current device (grids are sized from it) */
int orl_device_sm_count(int* sm_count_out);

/* ---- GAE / returns ---------------------------------------------------------------
 * Replaces ReplayData.compute_returns (openrl/buffers/replay_data.py:320-423, all 8
 * branches) and, fused behind it, the advantage construction of
 * PPOAlgorithm.train_ppo (openrl/algorithms/ppo.py:384-399) plus the moments needed
 * by its normalisation (ppo.py:402-409) and by ValueNorm.update on a full-buffer
 * minibatch (openrl/modules/utils/valuenorm.py:59-76).
 *
 * flags: */
#define ORL_GAE_USE_GAE 1            /* cfg.use_gae */
#define ORL_GAE_PROPER_TIME_LIMITS 2 /* cfg.use_proper_time_limits (reads bad_masks) */
#define ORL_GAE_DENORM 4             /* (use_popart|use_valuenorm) and normalizer given */
/*
 * rewards      (T,   B)   in
 * value_preds  (T+1, B)   in/out: row T is overwritten with next_value when USE_GAE
 * masks        (T+1, B)   in
 * bad_masks    (T+1, B)   in   (may be NULL unless PROPER_TIME_LIMITS)
 * active_masks (T+1, B)   in   (may be NULL: treated as all ones; only used for stats)
 * next_value   (B)        in   bootstrap value of slot T
 * vn_state     (3)        in   ValueNorm {running_mean, running_mean_sq, debiasing_term}
 *                              (valuenorm.py:27-35); required iff ORL_GAE_DENORM
 * returns      (T+1, B)   out  (row T = next_value when !USE_GAE, untouched otherwise)
 * advantages   (T,   B)   out  returns[:-1] - denorm(value_preds[:-1]); may be NULL
 * stats        (ORL_GAE_NSTATS doubles) out, may be NULL; ZEROED by the call, then
 *              accumulated: see ORL_GS_* indices.
 * gamma, gae_lambda are passed as double because the reference multiplies the two
 * Python floats in double before the product meets the float32 arrays.
 * Bit-exact with the reference's numpy float32 evaluation order (no FMA contraction).
 */
#define ORL_GAE_NSTATS 8
#define ORL_GS_ADV_SUM 0      /* sum adv            over all (t<T, b)            */
#define ORL_GS_ADV_SQSUM 1    /* sum adv^2                                        */
#define ORL_GS_COUNT 2        /* T*B                                              */
#define ORL_GS_ADV_ACT_SUM 3  /* sum adv   where active_masks[t] != 0             */
#define ORL_GS_ADV_ACT_SQSUM 4
#define ORL_GS_ACT_COUNT 5    /* number of active elements                        */
#define ORL_GS_RET_SUM 6      /* sum returns[t<T]                                 */
#define ORL_GS_RET_SQSUM 7    /* sum returns[t<T]^2                               */
int orl_gae(const float* rewards, float* value_preds, const float* masks,
            const float* bad_masks, const float* active_masks, const float* next_value,
            const float* vn_state, float* returns, float* advantages, double* stats,
            int T, int B, double gamma, double gae_lambda, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPENRL_B200_H */
