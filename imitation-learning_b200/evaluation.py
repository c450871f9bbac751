"""evaluate_agent (reference evaluation.py:11-35): the reference runs `num_episodes` greedy episodes one after
another with batch size 1; here all replicas x episodes run in parallel on the device with finished episodes
frozen, which is the same computation because episodes are independent given their initial states."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Union

import torch
from torch import Tensor

from . import _lib
from .environments import D4RLEnv
from .models import SoftActor


def evaluate_agent(actor: SoftActor, env: D4RLEnv, num_episodes: int, return_trajectories: bool = False, render: bool = False, reset_noise: Optional[Tensor] = None,
                   check_every: int = 50) -> Union[List[float], Tensor]:
  """Returns the list of episode returns (R == 1, like the reference) or a [R, num_episodes] tensor.
  `reset_noise` ([R * num_episodes, obs] U[0,1) draws) injects the initial states; default: the env's seeded stream."""
  if return_trajectories: raise NotImplementedError('return_trajectories (evaluation.py:30-33, save_trajectories) is outside the accelerated path')
  R, E, dev = actor.replicas, num_episodes, actor.device
  assert env.replicas == R
  eb = env.eval_batch(E)
  n, S, A = R * E, eb.S, actor.action_size
  u = env.reset_noise(n) if reset_noise is None else torch.as_tensor(reset_noise, dtype=torch.float32).to(dev).reshape(n, eb.obs).contiguous()
  state, nxt = torch.empty(n, S, device=dev), torch.empty(n, S, device=dev)
  reward, returns = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
  done, finished, running = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
  eb.reset(u, state)
  lib, h = _lib.lib(), _lib.handle()
  for t in range(eb.max_episode_steps):
    action = actor._run(state.view(R, E, S), want=('action', ))['action'].view(n, A)  # evaluation.py:21 greedy action
    eb.step(action, nxt, reward, done, frozen=finished)
    _lib.check(lib.il_eval_accumulate(h, n, reward.data_ptr(), done.data_ptr(), returns.data_ptr(), finished.data_ptr(), running.data_ptr(), _lib.stream()))
    state, nxt = nxt, state
    if (t + 1) % check_every == 0 and int(running.item()) == 0: break
  if R == 1: return [float(x) for x in returns.cpu()]
  return returns.view(R, E)
