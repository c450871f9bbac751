// Synthetic environment step shared by the training rollout (env.cu) and the evaluation rollout (eval.cu).
#pragma once
#include "common.cuh"

constexpr int MAX_OBS_PER_LANE = 8;  // obs <= 256

struct EnvStepOut {
  float reward;
  bool early, time_limit;
};

// One warp advances environment e: x' = tanh(x M + clamp(a) N + c), reward = x' . w_r - 1e-3 |clamp(a)|^2, writes the
// observation (+ absorbing bit 0, environments.py:39) to ns and bumps the episode step counter. All lanes return the same values.
__device__ __forceinline__ EnvStepOut env_step_core(const il_env& env, int e, int lane, const float* __restrict__ a, float* __restrict__ ns) {
  const int obs = env.obs, act = env.act;
  float* x = env.x + (int64_t)e * obs;
  float nx[MAX_OBS_PER_LANE];
#pragma unroll
  for (int q = 0; q < MAX_OBS_PER_LANE; ++q) {
    const int j = lane + 32 * q;
    float acc = 0.f;
    if (j < obs) {
      for (int i = 0; i < obs; ++i) acc = fmaf(x[i], __ldg(env.M + (int64_t)i * obs + j), acc);
      for (int k = 0; k < act; ++k) acc = fmaf(fminf(fmaxf(a[k], -1.f), 1.f), __ldg(env.N + (int64_t)k * obs + j), acc);  // environments.py:36 clamp
      acc = tanhf(acc + __ldg(env.c + j));
    }
    nx[q] = acc;
  }
  __syncwarp();
  float rew = 0.f;
#pragma unroll
  for (int q = 0; q < MAX_OBS_PER_LANE; ++q) {
    const int j = lane + 32 * q;
    if (j < obs) {
      x[j] = nx[q];
      rew = fmaf(nx[q], __ldg(env.w_r + j), rew);
    }
  }
  float a2 = 0.f;
  for (int k = lane; k < act; k += 32) {
    const float ak = fminf(fmaxf(a[k], -1.f), 1.f);
    a2 = fmaf(ak, ak, a2);
  }
  rew = warp_sum(rew) - 1e-3f * warp_sum(a2);
  const float x0 = __shfl_sync(0xffffffffu, nx[0], 0);
#pragma unroll
  for (int q = 0; q < MAX_OBS_PER_LANE; ++q) {
    const int j = lane + 32 * q;
    if (j < obs) ns[j] = nx[q];
  }
  int t = 0;
  if (lane == 0) {
    if (env.absorbing) ns[obs] = 0.f;  // environments.py:39
    t = env.t[e] + 1;
    env.t[e] = t;
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  EnvStepOut o;
  o.reward = rew;
  o.time_limit = t >= env.max_episode_steps;
  o.early = env.early_termination && fabsf(x0) > env.term_threshold;
  return o;
}
