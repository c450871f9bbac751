// Synthetic batched locomotion-shaped environment (stands in for gym/MuJoCo behind
// environments.py:29-40) + batched-evaluation bookkeeping (evaluation.py:11-35).
//   x' = tanh(x M + a N + c),  reward = x' . w_r - 1e-3 |a|^2,  early termination |x'_0| > threshold.
#include "env.cuh"

namespace {

// one warp per environment
__global__ void env_step_kernel(il_env env, int n_envs, const float* __restrict__ action, float* __restrict__ next_state, float* __restrict__ reward, int32_t* __restrict__ done,
                                int32_t* __restrict__ timeout, float* __restrict__ terminal_f, float* __restrict__ timeout_f, const int32_t* __restrict__ frozen) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
  if (e >= n_envs) return;
  if (frozen && frozen[e]) return;
  const int S = env.obs + (env.absorbing ? 1 : 0);
  const EnvStepOut o = env_step_core(env, e, lane, action + (int64_t)e * env.act, next_state + (int64_t)e * S);
  if (lane == 0) {
    const bool early = o.early, tl = o.time_limit;
    reward[e] = o.reward;
    if (done) done[e] = (early || tl) ? 1 : 0;
    if (timeout) timeout[e] = tl ? 1 : 0;
    if (terminal_f) terminal_f[e] = (early && !tl) ? 1.f : 0.f;  // train.py:157: terminal and t != max_episode_steps
    if (timeout_f) timeout_f[e] = tl ? 1.f : 0.f;                // train.py:157: t == max_episode_steps
  }
}

__global__ void env_reset_kernel(il_env env, int n_envs, const float* __restrict__ u, const int32_t* __restrict__ mask, float* __restrict__ state,
                                 const float* __restrict__ else_state) {
  const int S = env.obs + (env.absorbing ? 1 : 0);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_envs * S) return;
  const int e = (int)(i / S), j = (int)(i % S);
  if (mask && !mask[e]) {
    if (else_state) state[i] = else_state[i];
    return;
  }
  if (j < env.obs) {
    const float v = (__fmul_rn(u[(int64_t)e * env.obs + j], 2.f) - 1.f) * 0.1f;
    env.x[(int64_t)e * env.obs + j] = v;
    state[i] = v;
  } else {
    state[i] = 0.f;  // environments.py:32
  }
  if (j == 0) env.t[e] = 0;
}

__global__ void eval_accumulate_kernel(int n, const float* __restrict__ reward, const int32_t* __restrict__ done, float* __restrict__ returns, int32_t* __restrict__ finished,
                                       int32_t* __restrict__ n_unfinished) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int running = 0;
  if (i < n && !finished[i]) {
    returns[i] += reward[i];  // evaluation.py:24,28
    if (done[i]) finished[i] = 1;
    else running = 1;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, running);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_unfinished, __popc(ballot));
}

__global__ void rollout_bookkeep_kernel(int n, const float* __restrict__ reward, const int32_t* __restrict__ done, float* __restrict__ running, float* __restrict__ last_return,
                                        float* __restrict__ return_sum, int32_t* __restrict__ episodes, float* __restrict__ step_f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float run = running[i] + reward[i];  // train.py:155
  if (done[i]) {                         // train.py:160-168
    if (last_return) last_return[i] = run;
    if (return_sum) return_sum[i] += run;
    if (episodes) episodes[i] += 1;
    run = 0.f;
  }
  running[i] = run;
  if (step_f) step_f[i] += 1.f;
}

__global__ void zero_int_kernel(int32_t* p) { *p = 0; }

// single block: deterministic (sum, sum of squares, count)
__global__ void return_stats_kernel(const float* __restrict__ returns, int64_t n, float* __restrict__ out3) {
  __shared__ float red[32];
  float s = 0.f, s2 = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = returns[i];
    s += v;
    s2 += v * v;
  }
  s = block_sum(s, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    out3[0] = s;
    out3[1] = s2;
    out3[2] = (float)n;
  }
}

int check_env(const il_env* e, const char* what) {
  IL_CHECK(e && e->M && e->N && e->c && e->w_r && e->x && e->t, "%s: null env field", what);
  IL_CHECK(e->obs > 0 && e->obs <= 32 * MAX_OBS_PER_LANE && e->act > 0, "%s: obs=%d act=%d unsupported", what, e->obs, e->act);
  return 0;
}

}  // namespace

extern "C" int il_env_reset(il_handle* h, const il_env* env, int n_envs, const float* u, const int32_t* mask, float* state, const float* else_state, void* stream) {
  IL_CHECK(h && u && state && n_envs > 0, "il_env_reset: bad argument");
  IL_TRY(check_env(env, "il_env_reset"));
  const int S = env->obs + (env->absorbing ? 1 : 0);
  IL_LAUNCH(h, env_reset_kernel, (unsigned)(((int64_t)n_envs * S + 255) / 256), 256, 0, (cudaStream_t)stream, *env, n_envs, u, mask, state, else_state);
  return 0;
}

extern "C" int il_env_step(il_handle* h, const il_env* env, int n_envs, const float* action, float* next_state, float* reward, int32_t* done, int32_t* timeout,
                           float* terminal_f, float* timeout_f, const int32_t* frozen, void* stream) {
  IL_CHECK(h && action && next_state && reward && n_envs > 0, "il_env_step: bad argument");
  IL_TRY(check_env(env, "il_env_step"));
  IL_LAUNCH(h, env_step_kernel, (unsigned)(((int64_t)n_envs * 32 + 127) / 128), 128, 0, (cudaStream_t)stream, *env, n_envs, action, next_state, reward, done, timeout, terminal_f,
            timeout_f, frozen);
  return 0;
}

extern "C" int il_rollout_bookkeep(il_handle* h, int n_envs, const float* reward, const int32_t* done, float* running, float* last_return, float* return_sum,
                                   int32_t* episodes, float* step_f, void* stream) {
  IL_CHECK(h && reward && done && running && n_envs > 0, "il_rollout_bookkeep: bad argument");
  IL_LAUNCH(h, rollout_bookkeep_kernel, (unsigned)((n_envs + 255) / 256), 256, 0, (cudaStream_t)stream, n_envs, reward, done, running, last_return, return_sum, episodes, step_f);
  return 0;
}

extern "C" int il_eval_accumulate(il_handle* h, int n_envs, const float* reward, const int32_t* done, float* returns, int32_t* finished, int32_t* n_unfinished, void* stream) {
  IL_CHECK(h && reward && done && returns && finished && n_unfinished && n_envs > 0, "il_eval_accumulate: bad argument");
  IL_LAUNCH(h, zero_int_kernel, 1, 1, 0, (cudaStream_t)stream, n_unfinished);
  IL_LAUNCH(h, eval_accumulate_kernel, (unsigned)((n_envs + 255) / 256), 256, 0, (cudaStream_t)stream, n_envs, reward, done, returns, finished, n_unfinished);
  return 0;
}

extern "C" int il_return_stats(il_handle* h, const float* returns, int64_t n, float* out3, void* stream) {
  IL_CHECK(h && returns && out3 && n > 0, "il_return_stats: bad argument");
  IL_LAUNCH(h, return_stats_kernel, 1, 1024, 0, (cudaStream_t)stream, returns, n, out3);
  return 0;
}
