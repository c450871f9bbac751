"""TEST INFRASTRUCTURE ONLY — the reference's training loop (train.py:146-227) restated around the oracle port
(oracle/port.py) on the CPU twin of the synthetic environment. One instance = one reference process (one env, one
agent, one update per env step). Used (a) by the k-step loop parity tests with injected noise and (b) as the CPU
baseline / `bench.py --impl reference` arm (train.py, environments.py, utils.py of the reference need hydra / gym /
d4rl / matplotlib, which are not installed, so the loop itself cannot be imported — SURVEY.md §8c)."""
from __future__ import annotations

import math
import time
from typing import Dict, Optional

import numpy as np
import torch

from . import port


class NoiseSource:
  """Default noise: the reference's own RNG calls (torch global RNG for policy / GP noise, numpy global RNG for
  replay indices, a per-env generator for resets). Tests replace the methods to inject recorded noise."""

  uses_reference_rng_streams = True  # also reproduces RNG-stream side effects that do not touch the math (see OracleLoop.run_step)

  def __init__(self, seed: int, obs: int):
    self.gen = torch.Generator().manual_seed(seed)
    self.obs = obs

  def reset_u(self): return torch.rand(self.obs, generator=self.gen)
  def act_eps(self, A): return torch.randn(1, A)
  def policy_indices(self, mem, n): return mem.draw_indices(n)
  def expert_indices(self, mem, n): return mem.draw_indices(n)
  def gp_eps(self, B): return torch.rand(B)
  def mixup_eps(self, B, alpha): return torch.distributions.Beta(torch.full((B, ), float(alpha)), torch.full((B, ), float(alpha))).sample()  # training.py:106
  def sac_eps(self, B, A): return torch.randn(B, A), torch.randn(B, A)


class OracleLoop:
  def __init__(self, algorithm: str = 'GAIL', env_name: str = 'hopper', seed: int = 0, batch_size: int = 256, start: int = 1000, memory_size: int = 100000,
               hidden_size: int = 256, depth: int = 2, lr: float = 3e-4, weight_decay: float = 0.0, trajectories: int = 5, subsample: int = 1, absorbing: bool = True,
               mix_expert_data: str = 'none', imitation: Optional[dict] = None, discount: Optional[float] = None, polyak: Optional[float] = None,
               target_temperature: Optional[float] = None, max_episode_steps: int = 1000, expert_raw: Optional[Dict[str, torch.Tensor]] = None, init: Optional[dict] = None,
               bc_aux_loss: bool = False, build_expert_memory: bool = False):
    gail = algorithm == 'GAIL'
    self.algorithm, self.B, self.start, self.absorbing, self.mix = algorithm, batch_size, start, absorbing, mix_expert_data
    self.bc_aux_loss = bc_aux_loss  # train.py:201
    adril = algorithm == 'AdRIL'
    self.discount = (0.97 if gail else (0.98 if adril else 0.99)) if discount is None else discount  # GAIL.yaml:5 / AdRIL.yaml:5 / train_config.yaml:36
    self.polyak = (0.99 if gail else (0.98 if adril else 0.995)) if polyak is None else polyak  # GAIL.yaml:7 / AdRIL.yaml:6 / train_config.yaml:38
    tt = (-0.5 if gail else -1.0) if target_temperature is None else target_temperature  # GAIL.yaml:6 / train_config.yaml:37
    self.im = dict(hidden_size=64, learning_rate=3e-5, weight_decay=10.0, grad_penalty=1.0, spectral_norm=True, entropy_bonus=0.0, loss_function='BCE', reward_function='AIRL',
                   mixup_alpha=1.0, pos_class_prior=0.7, nonnegative_margin=float('inf'), reward_scale=5.0, reward_bandwidth_scale=5.0,
                   depth=1, activation='relu', reward_shaping=False, subtract_log_policy=False, state_only=False, update_freq=1250, balanced=True,
                   input_dropout=0.0, dropout=0.0, pretraining_iterations=0, quantile_cutoff=0.98, red_sigma=None)  # GAIL.yaml:8-27, PWIL.yaml:4-6, AdRIL.yaml:9-10, RED.yaml / DRIL.yaml, train_config.yaml:45
    self.im.update(imitation or {})
    np.random.seed(seed)
    torch.manual_seed(seed)  # train.py:51-52
    self.env = port.SyntheticEnv(env_name, absorbing, max_episode_steps)
    self.noise = NoiseSource(seed, self.env.obs)
    S, A = self.env.state_size, self.env.act
    self.S, self.A = S, A
    # train.py:60 loads the expert dataset for EVERY algorithm, SAC included, and train.py:173 samples it every update:
    # SAC never uses that batch, but its index draws advance the global numpy stream. With the default NoiseSource
    # (the reference's own RNG calls) the loop reproduces that; injected noise sources skip it (see run_step).
    self.expert_memory = None
    self._expert_args = (expert_raw, env_name, absorbing, trajectories, subsample, max_episode_steps, S, A)
    if algorithm != 'SAC' or build_expert_memory: self._build_expert_memory()
    # train.py:64-66 (construction order fixes the RNG stream: actor, critic_1, critic_2, then the discriminator)
    sizes_a, sizes_c = [S] + [hidden_size] * depth + [2 * A], [S + A] + [hidden_size] * depth + [1]
    if init is None:
      actor, twin = port.init_mlp(sizes_a), [port.init_mlp(sizes_c), port.init_mlp(sizes_c)]
    else:
      actor, twin = init['actor'], init['twin']
    self.agent = port.SacAgent(actor, twin, lr=lr, weight_decay=weight_decay)
    self.entropy_target = tt * A  # train.py:65
    self.memory = port.Replay(memory_size, S, A, absorbing)
    self.disc = None
    if gail:
      im = self.im
      H, din = im['hidden_size'], (S if im['state_only'] else S + A)

      def fcnn(sizes):  # models.py:48-69 with spectral_norm: per layer Linear() draws, orthogonal_ draws, then the u / v draws of the parametrization
        params, bufs = [], []
        for l in range(len(sizes) - 1):
          last = l == len(sizes) - 2
          p = port.init_mlp([sizes[l], sizes[l + 1]], final_gain=1.0 if last else torch.nn.init.calculate_gain(im['activation']))
          params += p
          if im['spectral_norm']: bufs.append(port.spectral_norm_init(p[0]))
        return params, bufs

      h = h_sn = None
      if init is None or 'g' not in init:
        if im['reward_shaping']:  # models.py:157-160: g is a plain nn.Linear (default init), h the MLP
          lin = torch.nn.Linear(din, 1)
          g, sn = [lin.weight.detach().clone(), lin.bias.detach().clone()], []
          if im['spectral_norm']: sn.append(port.spectral_norm_init(g[0]))
          h, h_sn = fcnn([S] + [H] * im['depth'] + [1])
        else:
          g, sn = fcnn([din] + [H] * im['depth'] + [1])  # models.py:162
      else:
        g, sn = init['g'], init.get('sn')
      self.disc = port.GailDiscriminator(g, sn if im['spectral_norm'] else None, self.discount, activation=im['activation'], reward_function=im['reward_function'],
                                         state_only=im['state_only'], subtract_log_policy=im['subtract_log_policy'], h=h, h_sn=h_sn if im['spectral_norm'] else None)
      self.disc_opt = torch.optim.AdamW(self.disc.parameters(), lr=self.im['learning_rate'], weight_decay=self.im['weight_decay'])  # train.py:84
    elif algorithm in ('DRIL', 'RED'):
      im = self.im
      din = S if im['state_only'] else S + A
      if algorithm == 'DRIL':  # train.py:74: SoftActor(state_size, action_size, cfg.imitation.discriminator) — a dropout policy
        self.disc = [torch.nn.Parameter(p) for p in port.init_mlp([S] + [im['hidden_size']] * im['depth'] + [2 * A], im['activation'])]
        params = self.disc
      else:  # train.py:82: predictor then target EmbeddingNetwork (models.py:265-266)
        sizes = [din] + [im['hidden_size']] * im['depth'] + [din]
        self.disc = port.RedDiscriminator(port.init_mlp(sizes, im['activation']), port.init_mlp(sizes, im['activation']), im['state_only'], im['activation'], im['input_dropout'], im['dropout'],
                                          im['red_sigma'])
        params = self.disc.parameters()
      self.disc_opt = torch.optim.AdamW(params, lr=im['learning_rate'], weight_decay=im['weight_decay'])  # train.py:84
    elif algorithm == 'AdRIL':
      assert mix_expert_data == 'mixed_batch'  # train.py:36
      self.disc = port.RewardRelabeller(self.im['update_freq'], self.im['balanced'])  # train.py:72
    elif algorithm == 'GMMIL':
      self.disc = port.GmmilDiscriminator()
    elif algorithm == 'PWIL':
      self.disc = port.PwilDiscriminator(self.expert_memory.data['states'], self.expert_memory.data['actions'], max_episode_steps, self.im['reward_scale'],
                                         self.im['reward_bandwidth_scale'])
    if algorithm == 'PWIL' and mix_expert_data != 'none':  # train.py:136-140: greedy PWIL rewards for the expert's own transitions
      with torch.inference_mode():
        d = self.expert_memory.data
        for i in range(self.expert_memory.size):
          d['rewards'][i] = self.disc.compute_reward(d['states'][i].unsqueeze(0), d['actions'][i].unsqueeze(0))
          if d['terminals'][i] or d['timeouts'][i]: self.disc.reset()
    if algorithm in ('PWIL', 'GMMIL') and mix_expert_data == 'prefill_memory': self.memory.transfer_transitions(self.expert_memory)  # train.py:141,143
    self._pretrained = algorithm not in ('DRIL', 'RED')  # DRIL / RED: call pretrain_discriminator() before the first run_step (train.py:117-133)
    self.t, self.train_return, self.step = 0, 0.0, 0
    self.state = self.env.reset(self.noise.reset_u())  # train.py:146
    self.episode_returns = []
    self.last = {}

  def _build_expert_memory(self):
    expert_raw, env_name, absorbing, trajectories, subsample, max_episode_steps, S, A = self._expert_args
    raw = expert_raw if expert_raw is not None else synthesize_raw_dataset(env_name, absorbing, max(trajectories, 5), max_episode_steps)
    tr = port.build_expert_transitions(raw, trajectories, subsample, absorbing)
    self.expert_memory = port.Replay(tr['states'].size(0), S, A, absorbing, transitions=tr)

  def run_step(self):
    """One iteration of train.py:149-211."""
    assert self._pretrained, 'DRIL / RED: call pretrain_discriminator() first (train.py:117-133)'
    self.step += 1
    step, env, agent = self.step, self.env, self.agent
    with torch.inference_mode():  # train.py:151-158
      action, _ = port.actor_sample(agent.actor, self.state, self.noise.act_eps(self.A))
      next_state, reward, terminal = env.step(action)
      self.t += 1
      self.train_return += reward
      if self.algorithm == 'PWIL': reward = self.disc.compute_reward(self.state, action)
      self.memory.append(step, self.state[0], action[0], reward, next_state[0], terminal and self.t != env.max_episode_steps, self.t == env.max_episode_steps)
      self.state = next_state
    if terminal:  # train.py:161-168
      if self.absorbing and self.t != env.max_episode_steps: self.memory.wrap_for_absorbing_states()
      if self.algorithm == 'PWIL': self.disc.reset()
      self.episode_returns.append(self.train_return)
      self.t, self.state, self.train_return = 0, env.reset(self.noise.reset_u()), 0.0
    if step >= self.start:  # train.py:171 (interval 1)
      B = self.B
      transitions = self.memory.gather(self.noise.policy_indices(self.memory, B))
      if self.algorithm != 'SAC':
        expert = self.expert_memory.gather(self.noise.expert_indices(self.expert_memory, B))  # train.py:173
      else:
        expert = None
        if getattr(self.noise, 'uses_reference_rng_streams', False):  # the unused expert batch of train.py:173
          if self.expert_memory is None: self._build_expert_memory()
          self.noise.expert_indices(self.expert_memory, B)
      if self.algorithm == 'GAIL':  # train.py:177-180
        eps_mix = self.noise.mixup_eps(B, self.im['mixup_alpha']) if self.im['loss_function'] == 'Mixup' else None  # drawn before the GP noise (training.py:106 then :118)
        eps_gp = self.noise.gp_eps(B) if self.im['grad_penalty'] > 0 else None
        self.last['gail'] = port.gail_update(self.disc, self.disc_opt, transitions, expert, eps_gp, loss_function=self.im['loss_function'],
                                             grad_penalty=self.im['grad_penalty'], entropy_bonus=self.im['entropy_bonus'], pos_class_prior=self.im['pos_class_prior'],
                                             nonnegative_margin=self.im['nonnegative_margin'], eps_mixup=eps_mix, actor=agent.actor)
      if self.algorithm in ('GAIL', 'GMMIL'):
        if self.mix == 'mixed_batch': port.mix_expert_agent_transitions(transitions, expert)  # train.py:183
        with torch.inference_mode():
          if self.algorithm == 'GAIL':  # train.py:194 via make_gail_input (models.py:145-149)
            lp = port.actor_log_prob(agent.actor, transitions['states'], transitions['actions']) if self.disc.subtract_log_policy else None
            rewards = self.disc.predict_reward(transitions['states'], transitions['actions'], transitions['next_states'], transitions['terminals'], lp)
          else: rewards = self.disc.predict_reward(transitions['states'], transitions['actions'], expert['states'], expert['actions'], transitions['weights'], expert['weights'])
        transitions['rewards'] = rewards.clone()
      if self.algorithm in ('DRIL', 'RED'):  # train.py:183,190-191,196-197
        if self.mix == 'mixed_batch': port.mix_expert_agent_transitions(transitions, expert)
        im = self.im
        with torch.inference_mode():
          if self.algorithm == 'DRIL': rewards = port.dril_predict_reward(self.disc, self.dril_q, transitions['states'], transitions['actions'], im['activation'], im['input_dropout'], im['dropout'])
          else: rewards = self.disc.predict_reward(transitions['states'], transitions['actions'])
        transitions['rewards'] = rewards.clone()
      if self.algorithm == 'AdRIL':  # train.py:188-189 (mix_expert_agent_transitions of :183 is skipped for AdRIL)
        with torch.inference_mode():
          self.disc.resample_and_relabel(transitions, expert, step, self.memory.num_trajectories, self.expert_memory.num_trajectories)
      if self.bc_aux_loss: port.behavioural_cloning_update(agent.actor, agent.opt_actor, expert)  # train.py:201 (the SAC actor optimiser)
      e1, e2 = self.noise.sac_eps(B, self.A)
      self.last['sac'] = port.sac_update(agent, transitions, e1, e2, self.discount, self.entropy_target, self.polyak)  # train.py:203
      self.last['rewards'] = transitions['rewards']

  def bc_pretrain(self, iterations: int, learning_rate: float, weight_decay: float):
    """train.py:95-101: behavioural cloning on shuffled expert minibatches (DataLoader(shuffle=True, drop_last=True), cycled)
    with a separate AdamW. Per epoch the global torch stream gives one draw to the DataLoader iterator (base seed) and one
    to the RandomSampler, which seeds the generator of that epoch's permutation."""
    opt = torch.optim.AdamW(self.agent.actor, lr=learning_rate, weight_decay=weight_decay)
    n, B, done = self.expert_memory.size, self.B, 0
    while done < iterations:
      torch.empty((), dtype=torch.int64).random_()  # the DataLoader iterator's base seed (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__)
      seed = int(torch.empty((), dtype=torch.int64).random_().item())  # then torch/utils/data/sampler.py RandomSampler.__iter__
      perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed)).tolist()
      for b in range(n // B):
        if done == iterations: break
        idx = perm[b * B:(b + 1) * B]
        batch = {k: self.expert_memory.data[k][idx] for k in port.FIELDS}
        port.behavioural_cloning_update(self.agent.actor, opt, batch)
        done += 1

  def pretrain_discriminator(self, batches=None):
    """train.py:117-133: DRIL's dropout policy ensemble is behaviour-cloned / RED's predictor is regressed onto its random target on shuffled expert minibatches
    (same DataLoader stream as bc_pretrain), then the uncertainty threshold / kernel bandwidth is fixed and (optionally) the expert data pre-fills the replay.
    `batches` (tests): the minibatch index lists, instead of the DataLoader's shuffles."""
    im, n, B, done = self.im, self.expert_memory.size, self.B, 0
    drop = dict(input_dropout=im['input_dropout'], dropout=im['dropout'], training=True)
    while done < im['pretraining_iterations']:
      if batches is None:
        torch.empty((), dtype=torch.int64).random_()
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed)).tolist()
      for b in range(n // B if batches is None else len(batches)):
        if done == im['pretraining_iterations']: break
        idx = perm[b * B:(b + 1) * B] if batches is None else list(batches[b])
        batch = {k: self.expert_memory.data[k][idx] for k in port.FIELDS}
        if self.algorithm == 'DRIL': port.behavioural_cloning_update(self.disc, self.disc_opt, batch, im['activation'], **drop)
        else: port.target_estimation_update(self.disc, self.disc_opt, batch)
        done += 1
    d = self.expert_memory.data
    with torch.inference_mode():
      if self.algorithm == 'DRIL': self.dril_q = port.dril_uncertainty_threshold(self.disc, d['states'], d['actions'], im['quantile_cutoff'], im['activation'], im['input_dropout'], im['dropout'])
      else: self.disc.set_sigma(d['states'][:B], d['actions'][:B])
    if self.mix == 'prefill_memory': self.memory.transfer_transitions(self.expert_memory)  # train.py:133
    if self.algorithm == 'RED': self.disc.training = False  # train.py:147
    self._pretrained = True

  def prefill(self, n: int):
    """Runs the warm-up phase of train.py:171 (`training.start` env steps without updates)."""
    start, self.start = self.start, 1 << 60
    for _ in range(n): self.run_step()
    self.start = start


def synthesize_raw_dataset(env_name: str, absorbing: bool, episodes: int, max_episode_steps: int = 1000) -> Dict[str, torch.Tensor]:
  """CPU twin of the product's expert-data synthesis: `episodes` greedy rollouts of the fixed tanh-MLP expert."""
  env = port.SyntheticEnv(env_name, absorbing, max_episode_steps)
  expert = port.expert_policy_params(env_name, absorbing)
  g = torch.Generator().manual_seed(977 + port.ENVS.index(env_name))
  u = torch.rand(episodes, env.obs, generator=g)
  obs, act, nobs, rew, term, tout = [], [], [], [], [], []
  with torch.inference_mode():
    for e in range(episodes):
      s, done = env.reset(u[e]), False
      while not done:
        a = port.actor_greedy_action(expert, s, 'tanh')
        ns, r, done = env.step(a)
        obs.append(s[0, :env.obs]); act.append(a[0]); nobs.append(ns[0, :env.obs]); rew.append(r)
        term.append(float(done and env.t != env.max_episode_steps)); tout.append(float(env.t == env.max_episode_steps))
        s = ns
  return dict(observations=torch.stack(obs), actions=torch.stack(act), next_observations=torch.stack(nobs), rewards=torch.tensor(rew), terminals=torch.tensor(term),
              timeouts=torch.tensor(tout))


def measure_steps_per_second(algorithm: str = 'GAIL', env_name: str = 'hopper', steps: int = 200, warmup: int = 10, seed: int = 0, batch_size: int = 256, prefill: int = 300,
                             threads: Optional[int] = None, max_episode_steps: int = 1000) -> Dict[str, float]:
  """Times `steps` full loop iterations (rollout + discriminator update + relabel + SAC update) of ONE reference
  process after `prefill` update-free steps; returns steps/s (= env-steps/s = grad-updates/s)."""
  if threads is not None: torch.set_num_threads(threads)
  loop = OracleLoop(algorithm, env_name, seed=seed, batch_size=batch_size, start=prefill + 1, max_episode_steps=max_episode_steps)
  loop.prefill(prefill)
  for _ in range(warmup): loop.run_step()
  t0 = time.perf_counter()
  for _ in range(steps): loop.run_step()
  dt = time.perf_counter() - t0
  return dict(steps_per_s=steps / dt, seconds=dt, steps=steps, threads=torch.get_num_threads())


def _worker(args):
  algorithm, env_name, steps, warmup, seed, batch_size, prefill = args
  torch.set_num_threads(1)
  return measure_steps_per_second(algorithm, env_name, steps, warmup, seed, batch_size, prefill, threads=1)


def measure_multiprocess(n_procs: int, algorithm: str = 'GAIL', env_name: str = 'hopper', steps: int = 200, warmup: int = 10, batch_size: int = 256, prefill: int = 300) -> Dict[str, float]:
  """The reference's own scaling model (train_all.py:26: one process per run): `n_procs` independent single-thread
  processes; aggregate steps/s = n_procs * steps / slowest wall time."""
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  with ctx.Pool(n_procs) as pool:
    res = pool.map(_worker, [(algorithm, env_name, steps, warmup, s, batch_size, prefill) for s in range(n_procs)])
  slowest = max(r['seconds'] for r in res)
  return dict(steps_per_s=n_procs * steps / slowest, seconds=slowest, steps=steps, procs=n_procs)
