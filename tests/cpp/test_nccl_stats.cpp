// Native path of the only collective: mgym_stats_allreduce(handle, ncclComm_t, ...) on two GPUs
// driven from one process (ncclCommInitAll + ncclGroupStart/End).  Each GPU owns half of a CartPole population;
// after K steps the all-reduced statistics on both GPUs must equal the statistics of ONE handle holding the whole
// population (Philox streams are keyed by the global env index).  Skips (exit 77) with fewer than 2 GPUs.
#include <cuda_runtime_api.h>
#include <nccl.h>

#include <cstdio>
#include <vector>

#include "../../include/mgym.h"

#define CK(x)                                                                  \
  do {                                                                         \
    int rc_ = (x);                                                             \
    if (rc_ != 0) {                                                            \
      std::fprintf(stderr, "FAIL %s -> %d (%s)\n", #x, rc_, mgym_last_error()); \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main() {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev < 2) {
    std::printf("SKIP: needs 2 GPUs, found %d\n", ndev);
    return 77;
  }
  const uint64_t n = 1 << 16, half = n / 2;
  const uint32_t K = 64;
  const uint64_t seed = 0xA11CE;
  ncclComm_t comms[2];
  int devs[2] = {0, 1};
  CK(ncclCommInitAll(comms, 2, devs));

  mgym_env* part[2];
  double* vec[2];
  cudaStream_t st[2];
  for (int d = 0; d < 2; ++d) {
    cudaSetDevice(d);
    cudaStreamCreate(&st[d]);
    mgym_config cfg;
    CK(mgym_config_default(MGYM_CARTPOLE_V1, &cfg));
    cfg.env_index_base = d * half;
    CK(mgym_create(MGYM_CARTPOLE_V1, half, d, seed, &cfg, &part[d]));
    CK(mgym_reset(part[d], nullptr, st[d]));
    CK(mgym_rollout(part[d], K / 2, nullptr, nullptr, nullptr, nullptr, nullptr, st[d]));  // device policy
    // per-call steps (the TMA-staged kernel, configured per device) with the same sampled actions
    unsigned char* acts = nullptr;
    cudaMalloc(reinterpret_cast<void**>(&acts), half);
    for (uint32_t k = 0; k < K / 2; ++k) {
      CK(mgym_sample_actions(part[d], acts, st[d]));
      CK(mgym_step(part[d], acts, nullptr, nullptr, nullptr, nullptr, st[d]));
    }
    cudaMalloc(reinterpret_cast<void**>(&vec[d]), 5 * sizeof(double));
  }
  ncclGroupStart();
  for (int d = 0; d < 2; ++d) {
    cudaSetDevice(d);
    CK(mgym_stats_allreduce(part[d], comms[d], vec[d], st[d]));
  }
  ncclGroupEnd();

  cudaSetDevice(0);
  mgym_env* whole;
  CK(mgym_create(MGYM_CARTPOLE_V1, n, 0, seed, nullptr, &whole));
  CK(mgym_reset(whole, nullptr, nullptr));
  CK(mgym_rollout(whole, K, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
  mgym_stats_t want;
  CK(mgym_stats_get(whole, &want, nullptr));

  int bad = 0;
  for (int d = 0; d < 2; ++d) {
    cudaSetDevice(d);
    cudaStreamSynchronize(st[d]);
    double got[5];
    cudaMemcpy(got, vec[d], sizeof(got), cudaMemcpyDeviceToHost);
    std::printf("gpu %d: episodes %.0f terminated %.0f truncated %.0f length_sum %.0f return_sum %.1f\n", d, got[0], got[1],
                got[2], got[3], got[4]);
    bad += got[0] != (double)want.episodes || got[1] != (double)want.terminated || got[2] != (double)want.truncated ||
           got[3] != (double)want.length_sum || got[4] != want.return_sum;
  }
  std::printf("one handle: episodes %llu length_sum %llu return_sum %.1f\n", (unsigned long long)want.episodes,
              (unsigned long long)want.length_sum, want.return_sum);
  if (bad || want.episodes == 0) {
    std::printf("FAIL\n");
    return 1;
  }
  std::printf("ALL OK\n");
  return 0;
}
