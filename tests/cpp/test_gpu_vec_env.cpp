// C++ parity test of the host layer (include/mgym.hpp) over the C ABI, written to read like the
// reference's own tests:
//   test_gym_against_python  src/testing.rs:34-146 (teacher-forced replay of python_tests/<env>/{inputs,output}.json)
//   test_cartpole / test_mountain_car / reward_is_one_when_not_terminated / *_invalid_action
//                            src/classic_control/cartpole.rs:365-434, mountain_car.rs:347-400
//   batched step vs the CPU oracle (oracle/mgym_oracle.h: test infrastructure), bit for bit.
// Fixtures come from tests/golden/<env>_gymnasium.txt (written by tests/golden/make_golden.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mgym.hpp"
#include "../../oracle/mgym_oracle.h"

static int g_failed = 0;
#define CHECK(cond, ...)                                            \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__);     \
      std::fprintf(stderr, __VA_ARGS__);                            \
      std::fprintf(stderr, "\n");                                   \
      ++g_failed;                                                   \
      return;                                                       \
    }                                                               \
  } while (0)

struct Expected {
  unsigned action;
  std::vector<float> observation;
  float reward;
  bool done, truncated;
};

static std::vector<Expected> load_fixture(const std::string& dir, const std::string& folder, int obs_dim) {
  std::ifstream in(dir + "/" + folder + "_gymnasium.txt");
  std::vector<Expected> v;
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    Expected e;
    int done, trunc;
    ss >> e.action;
    e.observation.resize(obs_dim);
    for (auto& o : e.observation) ss >> o;
    ss >> e.reward >> done >> trunc;
    e.done = done, e.truncated = trunc;
    v.push_back(e);
  }
  return v;
}

// src/testing.rs:34-146
static void test_gym_against_python(const std::string& dir, const std::string& folder, int kind) {
  const float reward_tol = 1e-4f, obs_tol = 1e-4f;  // testing.rs:42-45
  auto env = mgym::GpuVecEnv::builder(kind, 1).auto_reset(false).build();
  const int sd = env.state_dim();
  auto expected = load_fixture(dir, folder, sd);
  CHECK(expected.size() == 100, "%s: fixture has %zu steps", folder.c_str(), expected.size());
  auto reset_deterministic = [&] {  // cartpole.rs:437-442: reset() then the zero state
    env.reset();
    env.set_state(std::vector<float>(sd, 0.0f));
  };
  reset_deterministic();  // :65
  for (size_t i = 0; i < expected.size(); ++i) {
    if (i == 0 || expected[i - 1].done) {  // :73-75
      reset_deterministic();
    } else {  // :77-86 teacher forcing; set_state only replaces the state tensor (cartpole.rs:444-446)
      std::vector<float> st;
      std::vector<uint32_t> steps, sbt;
      env.get_state(st, steps, sbt);
      env.set_state(expected[i - 1].observation, &steps, &sbt);
    }
    auto info = env.step_host(std::vector<uint8_t>{(uint8_t)expected[i].action});  // :89
    CHECK(std::fabs(info.reward[0] - expected[i].reward) <= reward_tol, "%s step %zu reward %f", folder.c_str(), i,
          info.reward[0]);                                                                      // :99
    CHECK((bool)info.done[0] == expected[i].done, "%s: done mismatch at step %zu", folder.c_str(), i + 1);        // :106
    CHECK((bool)info.truncated[0] == expected[i].truncated, "%s: truncated mismatch at step %zu", folder.c_str(), i + 1);  // :114
    for (int j = 0; j < sd; ++j)                                                                // :124-133
      CHECK(std::fabs(info.state[j] - expected[i].observation[j]) < obs_tol && std::fabs(info.state[j] - expected[i].observation[j]) < 2e-7f,
            "%s: step %zu obs[%d] expected %g got %g", folder.c_str(), i, j, expected[i].observation[j], info.state[j]);
  }
  std::printf("ok   test_%s_against_python (100 teacher-forced steps, |err| < 2e-7)\n", folder.c_str());
}

// cartpole.rs:365-390, :405-434
static void test_cartpole() {
  auto env = mgym::GpuVecEnv::builder(MGYM_CARTPOLE_V1, 1).auto_reset(false).build();
  auto state = env.reset_host();
  CHECK(state.size() == 4, "reset dim");
  auto info = env.step_host(std::vector<uint8_t>{0});
  CHECK(info.state.size() == 4 && info.reward[0] == 1.0f && !info.done[0], "first step");
  bool done = false;
  for (int i = 0; i < 50 && !done; ++i) done = env.step_host(std::vector<uint8_t>{1}).done[0];
  CHECK(done, "pushing right must terminate within 51 steps");
  std::printf("ok   test_cartpole / reward_is_one_when_not_terminated\n");
}

// mountain_car.rs:347-372, :387-400
static void test_mountain_car() {
  auto env = mgym::GpuVecEnv::builder(MGYM_MOUNTAIN_CAR_V0, 1).auto_reset(false).build();
  auto state = env.reset_host();
  CHECK(state.size() == 2 && state[0] >= -0.6f && state[0] <= -0.4f && state[1] == 0.0f, "reset");
  auto info = env.step_host(std::vector<uint8_t>{0});
  CHECK(info.reward[0] == -1.0f && !info.done[0] && !info.truncated[0], "first step");
  std::printf("ok   test_mountain_car / reward_is_negative_one_when_not_terminated\n");
}

// cartpole.rs:392-403, mountain_car.rs:374-385: #[should_panic] becomes an exception
static void test_invalid_action() {
  for (auto [kind, bad] : {std::pair<int, int>{MGYM_CARTPOLE_V1, 2}, {MGYM_MOUNTAIN_CAR_V0, 3}}) {
    auto env = mgym::GpuVecEnv::builder(kind, 4).validate_actions(true).build();
    env.reset();
    std::vector<float> before, after;
    std::vector<uint32_t> steps0, steps1, sbt;
    env.get_state(before, steps0, sbt);
    bool threw = false;
    try {
      env.step_host(std::vector<uint8_t>{0, 1, (uint8_t)bad, 0});
    } catch (const mgym::InvalidAction&) {
      threw = true;
    }
    CHECK(threw, "kind %d: action %d must be rejected", kind, bad);
    // the reference asserts before any mutation (cartpole.rs:252): nothing was stepped
    env.get_state(after, steps1, sbt);
    CHECK(before == after && steps0 == steps1, "kind %d: the rejected batch must leave the handle untouched", kind);
    CHECK(!env.action_space_discrete().contains(bad) && env.action_space_discrete().contains(bad - 1), "Discrete::contains");
  }
  std::printf("ok   test_*_invalid_action\n");
}

// batched auto-reset steps, every kind, against the oracle's batched driver: bit equality
static void test_batched_against_oracle() {
  const uint64_t n = 2048;
  std::mt19937 rng(7);
  for (int kind = 0; kind < MGYM_NUM_KINDS; ++kind) {
    const uint64_t seed = 0x5EED + kind;
    auto env = mgym::GpuVecEnv::builder(kind, n).seed(seed).build();
    const int sd = env.state_dim(), od = env.obs_dim();
    oracle_config cfg;
    oracle_config_default(kind, &cfg);
    cfg.auto_reset = 1, cfg.seed = seed;
    std::vector<float> st(sd * n), ret(n, 0.0f), obs(od * n), rew(n);
    std::vector<uint32_t> steps(n, 0), sbt(n, 0);
    std::vector<uint8_t> flg(n);
    oracle_stats stats{};
    oracle_vec_reset(kind, &cfg, n, n, 0, nullptr, st.data(), steps.data(), sbt.data(), ret.data(), nullptr, 0, obs.data());
    auto got0 = env.reset_host();
    CHECK(std::memcmp(got0.data(), obs.data(), sizeof(float) * od * n) == 0, "kind %d: reset obs differ", kind);
    const int T = kind == 4 ? 60 : 150;
    for (int t = 0; t < T; ++t) {
      mgym::HostStepInfo info;
      if (env.continuous()) {
        std::vector<float> a(n);
        for (auto& x : a) x = std::uniform_real_distribution<float>(-2.2f, 2.2f)(rng);
        oracle_vec_step(kind, &cfg, n, n, t, st.data(), steps.data(), sbt.data(), ret.data(), a.data(), nullptr, 0,
                        obs.data(), rew.data(), flg.data(), nullptr, &stats);
        info = env.step_host(a);
      } else {
        std::vector<uint8_t> a(n);
        for (auto& x : a) x = (uint8_t)(rng() % mgym_num_actions(kind));
        oracle_vec_step(kind, &cfg, n, n, t, st.data(), steps.data(), sbt.data(), ret.data(), a.data(), nullptr, 0,
                        obs.data(), rew.data(), flg.data(), nullptr, &stats);
        info = env.step_host(a);
      }
      CHECK(std::memcmp(info.state.data(), obs.data(), sizeof(float) * od * n) == 0, "kind %d step %d: obs differ", kind, t);
      CHECK(std::memcmp(info.reward.data(), rew.data(), sizeof(float) * n) == 0, "kind %d step %d: reward differ", kind, t);
      for (uint64_t i = 0; i < n; ++i)
        CHECK((info.done[i] | (info.truncated[i] << 1)) == flg[i], "kind %d step %d env %llu: flags differ", kind, t,
              (unsigned long long)i);
    }
    auto s = env.stats();
    CHECK(s.episodes == stats.episodes && s.length_sum == stats.length_sum, "kind %d: statistics differ", kind);
    std::printf("ok   batched %-26s %d steps x %llu envs bit-identical to the oracle (%llu episodes)\n", mgym_kind_name(kind), T,
                (unsigned long long)n, (unsigned long long)s.episodes);
  }
}

// Gym::step on device buffers: the zero-copy observation of obs-is-state kinds equals mgym_get_obs
static void test_device_step_zero_copy() {
  const uint64_t n = 4096;
  auto env = mgym::GpuVecEnv::builder(MGYM_CARTPOLE_V1, n).seed(3).build();
  env.reset();
  unsigned char* acts = nullptr;
  float* obs = nullptr;
  cudaMalloc(reinterpret_cast<void**>(&acts), n);
  cudaMalloc(reinterpret_cast<void**>(&obs), sizeof(float) * 4 * n);
  for (int t = 0; t < 20; ++t) {
    env.sample_actions(acts);
    mgym::StepInfo info = env.step(acts);
    CHECK(info.state == mgym_state_ptr(env.handle()), "CartPole observation must alias the state rows");
    mgym::check(mgym_get_obs(env.handle(), obs, nullptr));
    std::vector<float> a(4 * n), b(4 * n);
    cudaMemcpy(a.data(), info.state, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), obs, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost);
    CHECK(std::memcmp(a.data(), b.data(), sizeof(float) * 4 * n) == 0, "step %d: zero-copy obs differs from get_obs", t);
  }
  cudaFree(acts);
  cudaFree(obs);
  std::printf("ok   Gym::step on device buffers (zero-copy observation)\n");
}

int main(int argc, char** argv) {
  const std::string golden = argc > 1 ? argv[1] : "tests/golden";
  try {
    test_gym_against_python(golden, "cartpole", MGYM_CARTPOLE_V1);
    test_gym_against_python(golden, "mountain_car", MGYM_MOUNTAIN_CAR_V0);
    test_cartpole();
    test_mountain_car();
    test_invalid_action();
    test_device_step_zero_copy();
    test_batched_against_oracle();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "FAIL exception: %s\n", e.what());
    return 2;
  }
  if (g_failed) {
    std::fprintf(stderr, "%d check(s) failed\n", g_failed);
    return 1;
  }
  std::printf("ALL OK\n");
  return 0;
}
