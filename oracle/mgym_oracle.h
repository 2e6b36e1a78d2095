/*
 * mgym_oracle.h -- CPU oracle for the classic-control step/reset hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under modurl_gym_b200/ may include, link
 * or load this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the CPU baseline, never as the product path.
 *
 * What it restates (all citations relative to /root/reference):
 *   CartPoleV1     src/classic_control/cartpole.rs:45-56 (constants),
 *                  :238-249 (reset), :251-348 (step)
 *   MountainCarV0  src/classic_control/mountain_car.rs:35-40 (constants),
 *                  :279-291 (reset), :293-330 (step)
 *   f32::sin/cos   -> platform libm.  Third-party, not under /root/reference:
 *                  glibc 2.39 sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,
 *                  sincosf.h,s_sincosf_data.c} (x86_64 multiarch FMA variant).
 *                  Restated here as oracle_sinf/oracle_cosf and checked
 *                  bit-for-bit against this image's libm (tests/test_oracle_trig.py,
 *                  oracle/exhaustive_trig.c).
 *   Tensor::rand   -> candle-core 0.9.1 (git rev d205fb4), not on disk and not
 *                  seed-reproducible here: PARITY UNPINNED for the random
 *                  stream.  Replaced by Philox4x32-10 (Salmon et al., SC'11),
 *                  sampled as f64 then cast to f32 like cartpole.rs:240-241.
 *
 * MountainCarContinuous-v0, Pendulum-v1 and Acrobot-v1 do NOT exist in the
 * reference (src/classic_control.rs:1-2).  Their oracle is this repo's own
 * f32 restatement of the Gymnasium equations: PARITY UNPINNED for those three.
 *
 * Pinning: tests/test_oracle_golden.py replays the reference's own
 * python_tests/{cartpole,mountain_car} fixtures (copied to tests/golden/) with
 * the teacher-forced protocol of src/testing.rs:65-134.
 *
 * Build: gcc -O2 -ffp-contract=off (mandatory: Rust never contracts a*b+c).
 */
#ifndef MGYM_ORACLE_H
#define MGYM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORACLE_CARTPOLE_V1 = 0,
  ORACLE_MOUNTAIN_CAR_V0 = 1,
  ORACLE_MOUNTAIN_CAR_CONTINUOUS_V0 = 2,
  ORACLE_PENDULUM_V1 = 3,
  ORACLE_ACROBOT_V1 = 4,
  ORACLE_NUM_KINDS = 5
};

/* flags byte: bit0 = StepInfo.done (terminated), bit1 = StepInfo.truncated */
#define ORACLE_FLAG_TERMINATED 1u
#define ORACLE_FLAG_TRUNCATED 2u

/* steps_beyond_terminated: Option<usize> encoded as 0 = None, k+1 = Some(k) */
#define ORACLE_SBT_NONE 0u

typedef struct oracle_config {
  int32_t auto_reset;          /* 1: same-step auto-reset (caller loop of cartpole.rs:468-470 folded in) */
  int32_t max_episode_steps;   /* 0 = none.  CartPole ignores it (500 is hard-coded, cartpole.rs:297) */
  int32_t sutton_barto_reward; /* cartpole.rs:39 */
  int32_t is_euler;            /* cartpole.rs:40 */
  float goal_velocity;         /* mountain_car.rs:33 */
  uint64_t seed;               /* Philox key */
  uint64_t env_index_base;     /* global index of env 0 of this slice */
} oracle_config;

void oracle_config_default(int kind, oracle_config *cfg);
int oracle_state_dim(int kind);
int oracle_obs_dim(int kind);
int oracle_action_is_continuous(int kind);

/* glibc 2.39 sinf/cosf, restated */
float oracle_sinf(float x);
float oracle_cosf(float x);

/* digest of (x, sin x, cos x) over magnitudes first, first+stride, ... and both signs (exhaustive device check) */
uint64_t oracle_trig_checksum(uint32_t first_bits, uint64_t count, uint32_t stride, int n_threads);

/* Philox4x32-10 */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* reset state of global env g for (tag, t); tag 0 = auto-reset at step t, 1 = explicit reset number t */
void oracle_reset_state(int kind, uint64_t seed, uint64_t g, uint64_t t, uint32_t tag, float *state);
/* device-policy action of global env g at step t (tag 2): discrete -> *a_u8, continuous -> *a_f32 */
void oracle_sample_action(int kind, uint64_t seed, uint64_t g, uint64_t t, uint8_t *a_u8, float *a_f32);

/*
 * One scalar env step, reference-faithful ("manual" semantics: no reset).
 *   state[state_dim] in/out, *steps and *sbt in/out, action: u8 value or f32.
 * Returns flags; *reward out.
 */
uint32_t oracle_env_step(int kind, const oracle_config *cfg, float *state, uint32_t *steps,
                         uint32_t *sbt, uint32_t action_u, float action_f, float *reward);
void oracle_env_obs(int kind, const float *state, float *obs);

/*
 * Batched driver with exactly the semantics of mgym_step (include/mgym.h):
 * SoA state[state_dim][ld], per-env counters, optional outputs may be NULL.
 * t = index of this step since creation.  An env that finishes is reset from the Philox block with counter
 * (global env index, t + 1 - episode length) = (g, step index at which the finished episode began).
 * reset_pool (optional, SoA [state_dim][pool_len]): injected reset states,
 * entry (g + t) % pool_len, replaces Philox.
 */
typedef struct oracle_stats {
  uint64_t episodes, terminated, truncated, length_sum;
  double return_sum;
} oracle_stats;

void oracle_vec_step(int kind, const oracle_config *cfg, uint64_t n, uint64_t ld, uint64_t t,
                     float *state, uint32_t *steps, uint32_t *sbt, float *ep_return,
                     const void *actions, const float *reset_pool, uint64_t pool_len,
                     float *obs_out, float *reward_out, uint8_t *flags_out, float *final_obs_out,
                     oracle_stats *stats);

/* K fused steps; actions NULL -> device policy (oracle_sample_action).  Trajectories are
 * time-major: obs[k][c][ld], reward[k][ld], flags[k][ld]. */
void oracle_vec_rollout(int kind, const oracle_config *cfg, uint64_t n, uint64_t ld, uint64_t t0,
                        uint32_t K, float *state, uint32_t *steps, uint32_t *sbt, float *ep_return,
                        const void *actions, const float *reset_pool, uint64_t pool_len,
                        float *obs_traj, float *reward_traj, uint8_t *flags_traj,
                        uint64_t *done_count, oracle_stats *stats);

void oracle_vec_reset(int kind, const oracle_config *cfg, uint64_t n, uint64_t ld, uint64_t reset_index,
                      const uint8_t *mask, float *state, uint32_t *steps, uint32_t *sbt,
                      float *ep_return, const float *reset_pool, uint64_t pool_len, float *obs_out);

/*
 * CPU baseline: the reference's caller loop (cartpole.rs:460-471): one env per
 * thread, step with a pseudo-random action, reset() on done (and on truncated).
 * Returns env-steps executed in total; *seconds = wall time of the slowest thread.
 */
uint64_t oracle_baseline_loop(int kind, const oracle_config *cfg, uint64_t steps_per_thread,
                              int n_threads, double *seconds, double *checksum);

#ifdef __cplusplus
}
#endif
#endif
