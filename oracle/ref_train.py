"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference `train.train(cfg)` (train.py:26-243) in the build container.

`train.py` imports hydra / matplotlib / seaborn (absent) and builds its environment through `gym.make` + d4rl. Here
those four imports are stubbed (decorator pass-through, no-op plotting) and `gym.make` returns a gym-shaped wrapper
around the synthetic environment twin (oracle/port.py:SyntheticEnv), so every line of the reference's own
loop, `D4RLEnv` (clamp, absorbing bit, `get_dataset`), replay memory, models and update rules executes unmodified.
Used by tests/test_oracle_loop_pinned.py to pin oracle/loop.py (the restated loop) against the real one.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from typing import Dict, Optional

import numpy as np
import torch

from . import port, refstub


class _FakeGymEnv:
  """gym.Env surface used by environments.py:20-61 on top of the synthetic dynamics (obs without the absorbing bit)."""

  def __init__(self, env_name: str, max_episode_steps: int, raw: Optional[Dict[str, torch.Tensor]], reset_seed: int):
    Box = sys.modules['gym.spaces'].Box
    self.twin = port.SyntheticEnv(env_name, False, max_episode_steps)
    self.observation_space = Box(low=-np.ones(self.twin.obs, np.float32), high=np.ones(self.twin.obs, np.float32))
    self.action_space = Box(low=-np.ones(self.twin.act, np.float32), high=np.ones(self.twin.act, np.float32))
    self._max_episode_steps = max_episode_steps
    self.ref_max_score, self.ref_min_score = 1000.0, 0.0  # SURVEY §8d placeholders
    self._raw = raw
    self._gen = torch.Generator().manual_seed(reset_seed)

  def get_dataset(self):
    return {k: v.numpy().copy() for k, v in self._raw.items()}

  def seed(self, seed): return [seed]

  def reset(self):
    return self.twin.reset(torch.rand(self.twin.obs, generator=self._gen))[0].numpy()

  def step(self, action):
    state, reward, terminal = self.twin.step(torch.as_tensor(action, dtype=torch.float32).unsqueeze(0))
    return state[0].numpy(), reward, terminal, {}

  def close(self): pass


def _to_dictconfig(d):
  return refstub.DictConfig({k: _to_dictconfig(v) if isinstance(v, dict) else v for k, v in d.items()})


_train_module = {}


def _load_train():
  if 'm' in _train_module: return _train_module['m']
  ref = refstub.load()
  for name in ('hydra', 'matplotlib', 'matplotlib.pyplot', 'seaborn'):
    if name not in sys.modules: sys.modules[name] = types.ModuleType(name)
  sys.modules['hydra'].main = lambda **kw: (lambda f: f)
  noop = lambda *a, **k: None
  plt, sns = sys.modules['matplotlib.pyplot'], sys.modules['seaborn']
  for fn in ('fill_between', 'xlim', 'xlabel', 'ylabel', 'title', 'savefig', 'close'): setattr(plt, fn, noop)
  sys.modules['matplotlib'].pyplot = plt
  sns.set, sns.lineplot = noop, noop
  if not hasattr(sys.modules['omegaconf'], 'OmegaConf'): sys.modules['omegaconf'].OmegaConf = object
  saved = {k: sys.modules.get(k) for k in ('memory', 'models', 'training', 'evaluation', 'environments', 'utils', 'train')}
  sys.modules.update(memory=ref.memory, models=ref.models, training=ref.training, evaluation=ref.evaluation, environments=ref.environments)
  sys.path.insert(0, refstub.REFERENCE_DIR)
  try:
    sys.modules.pop('utils', None)
    sys.modules.pop('train', None)
    import train  # the reference's train.py, unmodified
  finally:
    sys.path.remove(refstub.REFERENCE_DIR)
    for k, v in saved.items():
      if v is None: sys.modules.pop(k, None)
      else: sys.modules[k] = v
  _train_module['m'] = train
  return train


def run_reference_train(cfg: dict, raw: Optional[Dict[str, torch.Tensor]], max_episode_steps: int) -> Dict[str, object]:
  """Executes train.train(cfg) and returns what it wrote: agent / discriminator state dicts, metrics and the score."""
  train = _load_train()
  ref = refstub.load()
  made = []

  def make(name):
    # train.py:55: the training env is created first (load_data=True), then the evaluation env; each gets its own reset stream
    env = _FakeGymEnv(name.split('-')[0], max_episode_steps, raw, reset_seed=cfg['seed'] + 10007 * len(made))
    made.append(env)
    return env
  ref.environments.gym.make = make
  cwd, threads = os.getcwd(), torch.get_num_threads()
  with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    try:
      torch.set_num_threads(1)
      score = train.train(_to_dictconfig(cfg), file_prefix='')
      out = dict(score=float(score), agent=torch.load('agent.pth', weights_only=False), metrics=torch.load('metrics.pth', weights_only=False))
      if os.path.exists('discriminator.pth'): out['discriminator'] = torch.load('discriminator.pth', weights_only=False)
    finally:
      os.chdir(cwd)
      torch.set_num_threads(threads)
  return out
