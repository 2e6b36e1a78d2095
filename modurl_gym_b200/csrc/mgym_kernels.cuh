// mgym_kernels.cuh -- the two kernel modes of the hot path.
//
//   step_kernel     one Gym::step per env per launch: SoA state rows are read and rewritten with
//                   128-bit accesses (one thread = V consecutive envs), actions / reward / flags /
//                   step counters are packed vector accesses.  HBM-bound by construction.
//   rollout_kernel  K fused steps: state and counters stay in registers, only the trajectory
//                   (obs / reward / flags) is written, actions are read or drawn from Philox.
//
// Both are persistent grid-stride kernels (grid = multiple of the SM count) so the episode
// statistics reduce to one set of atomics per CTA.
#pragma once

#include "mgym_device.cuh"

namespace mgym {

enum CounterMode : int { CNT_NONE = 0, CNT_U16 = 1, CNT_U32 = 2 };

struct KernelParams {
  // resident env state (owned by the handle)
  float* state;        // [SD][n]
  void* steps;         // uint16_t[n] or uint32_t[n] or null, see CounterMode
  uint32_t* sbt;       // manual CartPole only: steps_beyond_terminated (0 = None, k+1 = Some(k))
  float* ep_return;    // per-env running return (kinds without an analytic return), or null
  // caller buffers
  const void* actions;  // step: [n]; rollout: [K][n] or null (device policy)
  float* obs_out;       // step: [OD][n]; rollout: [K][OD][n]
  float* reward_out;    // step: [n];     rollout: [K][n]
  uint8_t* flags_out;   // step: [n];     rollout: [K][n]
  float* final_obs_out; // step only: pre-reset observation [OD][n]
  // reset source
  const float* reset_pool;  // [SD][pool_len] or null
  uint64_t pool_len;
  // statistics: u64 {episodes, terminated, truncated, length_sum}, then double return_sum
  unsigned long long* stats;
  unsigned long long* done_count;  // rollout: finished env-steps
  uint32_t* bad_action;            // validate_actions: set to 1 on an out-of-range discrete action
  uint64_t n, seed, env_base, t;
  uint32_t K;
  EnvConsts k;
};

// ---- vector load/store of V consecutive elements --------------------------------------
template <typename T, int V>
struct Vec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ldv(const T* p) {
  Vec<T, V> r;
  if constexpr (V == 1) {
    r.v[0] = *p;
  } else if constexpr (sizeof(T) * V == 16) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    *reinterpret_cast<uint4*>(&r) = u;
  } else if constexpr (sizeof(T) * V == 8) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    *reinterpret_cast<uint2*>(&r) = u;
  } else {
    static_assert(sizeof(T) * V == 4, "unsupported vector width");
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    *reinterpret_cast<uint32_t*>(&r) = u;
  }
  return r;
}

template <typename T, int V>
__device__ __forceinline__ void stv(T* p, const Vec<T, V>& r) {
  if constexpr (V == 1) {
    *p = r.v[0];
  } else if constexpr (sizeof(T) * V == 16) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&r);
  } else if constexpr (sizeof(T) * V == 8) {
    *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(&r);
  } else {
    static_assert(sizeof(T) * V == 4, "unsupported vector width");
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&r);
  }
}

// ---- per-thread episode statistics, reduced once per CTA ---------------------------------
struct StatAcc {
  uint32_t episodes = 0, terminated = 0, truncated = 0;
  unsigned long long length_sum = 0;
  double return_sum = 0.0;
  unsigned long long done_steps = 0;
};

__device__ __forceinline__ void stats_flush(const StatAcc& a, const KernelParams& p) {
  __shared__ unsigned long long sh_u[5];
  __shared__ double sh_d;
  if (threadIdx.x == 0) {
    sh_u[0] = sh_u[1] = sh_u[2] = sh_u[3] = sh_u[4] = 0ull;
    sh_d = 0.0;
  }
  __syncthreads();
  unsigned long long u0 = a.episodes, u1 = a.terminated, u2 = a.truncated, u3 = a.length_sum, u4 = a.done_steps;
  double d = a.return_sum;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    u0 += __shfl_xor_sync(0xffffffffu, u0, o);
    u1 += __shfl_xor_sync(0xffffffffu, u1, o);
    u2 += __shfl_xor_sync(0xffffffffu, u2, o);
    u3 += __shfl_xor_sync(0xffffffffu, u3, o);
    u4 += __shfl_xor_sync(0xffffffffu, u4, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if ((threadIdx.x & 31) == 0 && (u0 | u4)) {
    atomicAdd(&sh_u[0], u0);
    atomicAdd(&sh_u[1], u1);
    atomicAdd(&sh_u[2], u2);
    atomicAdd(&sh_u[3], u3);
    atomicAdd(&sh_u[4], u4);
    atomicAdd(&sh_d, d);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p.stats && sh_u[0]) {
      atomicAdd(&p.stats[0], sh_u[0]);
      atomicAdd(&p.stats[1], sh_u[1]);
      atomicAdd(&p.stats[2], sh_u[2]);
      atomicAdd(&p.stats[3], sh_u[3]);
      atomicAdd(reinterpret_cast<double*>(&p.stats[4]), sh_d);
    }
    if (p.done_count && sh_u[4]) atomicAdd(p.done_count, sh_u[4]);
  }
}

// ---- one env: step, then (AUTO) same-step reset --------------------------------------------
// Returns flags.  st / steps / sbt / ret are updated in place; obs holds the observation the
// caller sees (post-reset when the episode ended), fin the pre-reset observation.
template <int KIND, bool AUTO, bool WANT_FINAL>
__device__ __forceinline__ uint32_t env_transition(const KernelParams& p, bool count, uint64_t local_index, uint64_t t,
                                                   typename Env<KIND>::act_t action, float (&st)[Env<KIND>::SD],
                                                   uint32_t& steps, uint32_t& sbt, float& ret, bool track_ret,
                                                   float& reward, float (&obs)[Env<KIND>::OD],
                                                   float (&fin)[Env<KIND>::OD], StatAcc& acc) {
  using E = Env<KIND>;
  if constexpr (AUTO) sbt = SBT_NONE;  // auto-reset presumes reset() precedes every episode
  const uint32_t flags = E::step(st, action, steps, sbt, p.k, reward);
  if (track_ret) ret = fadd(ret, reward);
  E::obs(st, obs);
  if constexpr (WANT_FINAL) {
#pragma unroll
    for (int c = 0; c < E::OD; ++c) fin[c] = obs[c];
  }
  if constexpr (AUTO) {
    if (flags) {
      if (count) {
        acc.episodes += 1;
        acc.terminated += (flags & FLAG_TERMINATED) ? 1u : 0u;
        acc.truncated += (flags & FLAG_TRUNCATED) ? 1u : 0u;
        acc.length_sum += steps;
        acc.return_sum += (double)(E::ANALYTIC_RETURN ? E::episode_return(p.k, steps, flags) : ret);
      }
      const uint64_t g = p.env_base + local_index;
      if (p.reset_pool) {
        const uint64_t j = (g + t) % p.pool_len;
#pragma unroll
        for (int c = 0; c < E::SD; ++c) st[c] = p.reset_pool[(uint64_t)c * p.pool_len + j];
      } else {
        E::reset(philox_env(p.seed, g, t, TAG_AUTO_RESET), st);
      }
      steps = 0;
      ret = 0.0f;
      E::obs(st, obs);
    }
  }
  return flags;
}

template <int MODE>
struct CounterType {
  using type = uint32_t;
};
template <>
struct CounterType<CNT_U16> {
  using type = uint16_t;
};

// =============================================================================================
// Mode 1: per-call step kernel
// =============================================================================================
template <int KIND, int V, bool AUTO, int CNT>
__global__ void __launch_bounds__(256) step_kernel(const __grid_constant__ KernelParams p) {
  using E = Env<KIND>;
  using act_t = typename E::act_t;
  using cnt_t = typename CounterType<CNT>::type;
  constexpr int SD = E::SD, OD = E::OD;
  StatAcc acc;
  const uint64_t groups = p.n / V;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const bool track_ret = p.ep_return != nullptr;
  const bool want_final = p.final_obs_out != nullptr;

  for (uint64_t grp = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; grp < groups; grp += stride) {
    const uint64_t base = grp * V;
    Vec<float, V> s[SD];
#pragma unroll
    for (int c = 0; c < SD; ++c) s[c] = ldv<float, V>(p.state + (uint64_t)c * p.n + base);
    const Vec<act_t, V> a = ldv<act_t, V>(reinterpret_cast<const act_t*>(p.actions) + base);
    Vec<cnt_t, V> cnt;
    if constexpr (CNT != CNT_NONE) cnt = ldv<cnt_t, V>(reinterpret_cast<const cnt_t*>(p.steps) + base);
    Vec<uint32_t, V> sb;
    if constexpr (!AUTO && KIND == 0) sb = ldv<uint32_t, V>(p.sbt + base);
    Vec<float, V> er;
    if (track_ret) er = ldv<float, V>(p.ep_return + base);

    Vec<float, V> o[OD], f[OD], rw;
    Vec<uint8_t, V> fl;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float st[SD], obs[OD], fin[OD], reward;
#pragma unroll
      for (int c = 0; c < SD; ++c) st[c] = s[c].v[v];
      uint32_t steps = 0, sbt = SBT_NONE;
      if constexpr (CNT != CNT_NONE) steps = cnt.v[v];
      if constexpr (!AUTO && KIND == 0) sbt = sb.v[v];
      float ret = track_ret ? er.v[v] : 0.0f;
      if constexpr (!E::CONTINUOUS) {
        if (p.bad_action && a.v[v] >= E::NUM_ACTIONS) *p.bad_action = 1u;
      }
      uint32_t flags;
      if (want_final)
        flags = env_transition<KIND, AUTO, true>(p, true, base + v, p.t, a.v[v], st, steps, sbt, ret, track_ret, reward,
                                                 obs, fin, acc);
      else
        flags = env_transition<KIND, AUTO, false>(p, true, base + v, p.t, a.v[v], st, steps, sbt, ret, track_ret, reward,
                                                  obs, fin, acc);
#pragma unroll
      for (int c = 0; c < SD; ++c) s[c].v[v] = st[c];
#pragma unroll
      for (int c = 0; c < OD; ++c) {
        o[c].v[v] = obs[c];
        f[c].v[v] = fin[c];
      }
      rw.v[v] = reward;
      fl.v[v] = (uint8_t)flags;
      if constexpr (CNT != CNT_NONE) cnt.v[v] = (cnt_t)steps;
      if constexpr (!AUTO && KIND == 0) sb.v[v] = sbt;
      if (track_ret) er.v[v] = ret;
    }

#pragma unroll
    for (int c = 0; c < SD; ++c) stv<float, V>(p.state + (uint64_t)c * p.n + base, s[c]);
    if constexpr (CNT != CNT_NONE) stv<cnt_t, V>(reinterpret_cast<cnt_t*>(p.steps) + base, cnt);
    if constexpr (!AUTO && KIND == 0) stv<uint32_t, V>(p.sbt + base, sb);
    if (track_ret) stv<float, V>(p.ep_return + base, er);
    if (p.obs_out) {
#pragma unroll
      for (int c = 0; c < OD; ++c) stv<float, V>(p.obs_out + (uint64_t)c * p.n + base, o[c]);
    }
    if (want_final) {
#pragma unroll
      for (int c = 0; c < OD; ++c) stv<float, V>(p.final_obs_out + (uint64_t)c * p.n + base, f[c]);
    }
    if (p.reward_out) stv<float, V>(p.reward_out + base, rw);
    if (p.flags_out) stv<uint8_t, V>(p.flags_out + base, fl);
  }
  if constexpr (AUTO) stats_flush(acc, p);
}

// =============================================================================================
// Mode 2: fused K-step rollout kernel
// =============================================================================================
template <int KIND, int V, bool AUTO, int CNT>
__global__ void __launch_bounds__(256) rollout_kernel(const __grid_constant__ KernelParams p) {
  using E = Env<KIND>;
  using act_t = typename E::act_t;
  using cnt_t = typename CounterType<CNT>::type;
  constexpr int SD = E::SD, OD = E::OD;
  StatAcc acc;
  uint32_t warp_dones = 0;  // identical in every lane of the warp (ballot + popc)
  const uint64_t groups = p.n / V;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const bool track_ret = p.ep_return != nullptr;
  const bool policy = p.actions == nullptr;
  const act_t* actions = reinterpret_cast<const act_t*>(p.actions);
  const uint32_t lane = threadIdx.x & 31;

  // warp-uniform trip count so every lane reaches the ballots
  for (uint64_t grp0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); grp0 < groups; grp0 += stride) {
    const uint64_t grp = grp0 + lane;
    const bool active = grp < groups;
    const uint64_t base = active ? grp * V : 0;
    float st[V][SD];
    uint32_t steps[V], sbt[V];
    float ret[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      steps[v] = 0;
      sbt[v] = SBT_NONE;
      ret[v] = 0.0f;
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < SD; ++c) {
        const Vec<float, V> s = ldv<float, V>(p.state + (uint64_t)c * p.n + base);
#pragma unroll
        for (int v = 0; v < V; ++v) st[v][c] = s.v[v];
      }
      if constexpr (CNT != CNT_NONE) {
        const Vec<cnt_t, V> cnt = ldv<cnt_t, V>(reinterpret_cast<const cnt_t*>(p.steps) + base);
#pragma unroll
        for (int v = 0; v < V; ++v) steps[v] = cnt.v[v];
      }
      if constexpr (!AUTO && KIND == 0) {
        const Vec<uint32_t, V> sb = ldv<uint32_t, V>(p.sbt + base);
#pragma unroll
        for (int v = 0; v < V; ++v) sbt[v] = sb.v[v];
      }
      if (track_ret) {
        const Vec<float, V> er = ldv<float, V>(p.ep_return + base);
#pragma unroll
        for (int v = 0; v < V; ++v) ret[v] = er.v[v];
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < SD; ++c) st[v][c] = 0.0f;
    }

    Vec<act_t, V> a_next;
#pragma unroll
    for (int v = 0; v < V; ++v) a_next.v[v] = act_t(0);
    if (!policy && active) a_next = ldv<act_t, V>(actions + base);

    for (uint32_t kk = 0; kk < p.K; ++kk) {
      const uint64_t t = p.t + kk;
      Vec<act_t, V> a = a_next;
      if (policy) {
        // Space::sample: one Philox block serves 4 consecutive envs (global group g >> 2)
        if constexpr (V == 4) {
          const uint4 w = philox_env(p.seed, (p.env_base + base) >> 2, t, TAG_ACTION);
          a.v[0] = action_from_word<KIND>(w.x);
          a.v[1] = action_from_word<KIND>(w.y);
          a.v[2] = action_from_word<KIND>(w.z);
          a.v[3] = action_from_word<KIND>(w.w);
        } else {
          const uint64_t g = p.env_base + base;
          const uint4 w = philox_env(p.seed, g >> 2, t, TAG_ACTION);
          const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
          a.v[0] = action_from_word<KIND>(ws[g & 3]);
        }
      } else if (active && kk + 1 < p.K) {
        a_next = ldv<act_t, V>(actions + (uint64_t)(kk + 1) * p.n + base);  // prefetch next step
      }

      Vec<float, V> o[OD], rw;
      Vec<uint8_t, V> fl;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float obs[OD], fin[OD], reward;
        if constexpr (!E::CONTINUOUS) {
          if (p.bad_action && a.v[v] >= E::NUM_ACTIONS) *p.bad_action = 1u;
        }
        const uint32_t flags = env_transition<KIND, AUTO, false>(p, active, base + v, t, a.v[v], st[v], steps[v], sbt[v],
                                                                 ret[v], track_ret, reward, obs, fin, acc);
#pragma unroll
        for (int c = 0; c < OD; ++c) o[c].v[v] = obs[c];
        rw.v[v] = reward;
        fl.v[v] = (uint8_t)flags;
        warp_dones += __popc(__ballot_sync(0xffffffffu, active && flags != 0));
      }
      if (active) {
        if (p.obs_out) {
          float* ob = p.obs_out + (uint64_t)kk * OD * p.n + base;
#pragma unroll
          for (int c = 0; c < OD; ++c) stv<float, V>(ob + (uint64_t)c * p.n, o[c]);
        }
        if (p.reward_out) stv<float, V>(p.reward_out + (uint64_t)kk * p.n + base, rw);
        if (p.flags_out) stv<uint8_t, V>(p.flags_out + (uint64_t)kk * p.n + base, fl);
      }
    }

    if (active) {
#pragma unroll
      for (int c = 0; c < SD; ++c) {
        Vec<float, V> s;
#pragma unroll
        for (int v = 0; v < V; ++v) s.v[v] = st[v][c];
        stv<float, V>(p.state + (uint64_t)c * p.n + base, s);
      }
      if constexpr (CNT != CNT_NONE) {
        Vec<cnt_t, V> cnt;
#pragma unroll
        for (int v = 0; v < V; ++v) cnt.v[v] = (cnt_t)steps[v];
        stv<cnt_t, V>(reinterpret_cast<cnt_t*>(p.steps) + base, cnt);
      }
      if constexpr (!AUTO && KIND == 0) {
        Vec<uint32_t, V> sb;
#pragma unroll
        for (int v = 0; v < V; ++v) sb.v[v] = sbt[v];
        stv<uint32_t, V>(p.sbt + base, sb);
      }
      if (track_ret) {
        Vec<float, V> er;
#pragma unroll
        for (int v = 0; v < V; ++v) er.v[v] = ret[v];
        stv<float, V>(p.ep_return + base, er);
      }
    }
  }
  if (lane == 0) acc.done_steps = warp_dones;
  stats_flush(acc, p);
}

// =============================================================================================
// cold kernels: reset, observation, action sampling, counter conversion
// =============================================================================================
template <int KIND, int CNT>
__global__ void reset_kernel(const KernelParams p, const uint8_t* mask, uint64_t reset_index) {
  using E = Env<KIND>;
  using cnt_t = typename CounterType<CNT>::type;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float st[E::SD], obs[E::OD];
  if (mask && !mask[i]) {
    if (p.obs_out) {
#pragma unroll
      for (int c = 0; c < E::SD; ++c) st[c] = p.state[(uint64_t)c * p.n + i];
      E::obs(st, obs);
#pragma unroll
      for (int c = 0; c < E::OD; ++c) p.obs_out[(uint64_t)c * p.n + i] = obs[c];
    }
    return;
  }
  const uint64_t g = p.env_base + i;
  if (p.reset_pool) {
    const uint64_t j = (g + reset_index) % p.pool_len;
#pragma unroll
    for (int c = 0; c < E::SD; ++c) st[c] = p.reset_pool[(uint64_t)c * p.pool_len + j];
  } else {
    E::reset(philox_env(p.seed, g, reset_index, TAG_RESET), st);
  }
#pragma unroll
  for (int c = 0; c < E::SD; ++c) p.state[(uint64_t)c * p.n + i] = st[c];
  if constexpr (CNT != CNT_NONE) reinterpret_cast<cnt_t*>(p.steps)[i] = 0;  // cartpole.rs:243
  if (p.sbt) p.sbt[i] = SBT_NONE;                                            // cartpole.rs:239
  if (p.ep_return) p.ep_return[i] = 0.0f;
  if (p.obs_out) {
    E::obs(st, obs);
#pragma unroll
    for (int c = 0; c < E::OD; ++c) p.obs_out[(uint64_t)c * p.n + i] = obs[c];
  }
}

template <int KIND>
__global__ void obs_kernel(const float* state, float* obs_out, uint64_t n) {
  using E = Env<KIND>;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float st[E::SD], obs[E::OD];
#pragma unroll
  for (int c = 0; c < E::SD; ++c) st[c] = state[(uint64_t)c * n + i];
  E::obs(st, obs);
#pragma unroll
  for (int c = 0; c < E::OD; ++c) obs_out[(uint64_t)c * n + i] = obs[c];
}

template <int KIND>
__global__ void sample_actions_kernel(typename Env<KIND>::act_t* out, uint64_t n, uint64_t seed, uint64_t env_base,
                                      uint64_t t) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t g = env_base + i;
  const uint4 w = philox_env(seed, g >> 2, t, TAG_ACTION);
  const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
  out[i] = action_from_word<KIND>(ws[g & 3]);
}

template <typename From, typename To>
__global__ void convert_kernel(const From* in, To* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (To)in[i];
}

__global__ void fill_u32_kernel(uint32_t* out, uint32_t value, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = value;
}

// stats counters -> 5 doubles (the vector a host layer all-reduces)
__global__ void stats_export_kernel(const unsigned long long* stats, double* out) {
  if (threadIdx.x < 4) out[threadIdx.x] = (double)stats[threadIdx.x];
  if (threadIdx.x == 4) out[4] = *reinterpret_cast<const double*>(&stats[4]);
}

// exposes sin_ref / cos_ref for the parity tests (tests/test_gpu_trig.py)
__global__ void trig_probe_kernel(const float* x, float* s, float* c, float* s_only, float* c_only, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float sv, cv;
  sincos_ref(x[i], sv, cv);
  s[i] = sv;
  c[i] = cv;
  s_only[i] = sin_ref(x[i]);
  c_only[i] = cos_ref(x[i]);
}

__global__ void philox_probe_kernel(const uint32_t* ctr_key, uint32_t* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* ck = ctr_key + i * 6;
  const uint4 r = philox4x32_10(make_uint4(ck[0], ck[1], ck[2], ck[3]), ck[4], ck[5]);
  out[i * 4 + 0] = r.x, out[i * 4 + 1] = r.y, out[i * 4 + 2] = r.z, out[i * 4 + 3] = r.w;
}

}  // namespace mgym
