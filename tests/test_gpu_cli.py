"""End to end through the command line: `python train.py algorithm=... env=... key=value ...` as a user of the
reference would type it (train.py:21-23), then the files the reference writes (train.py:232-239) are reloaded and checked for
the reference's schema — `agent.pth` loads into an `nn.Sequential` laid out like `_create_fcnn` (models.py:48-69)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fcnn(sizes, input_dropout=0.0, dropout=0.0):
  """A torch module with the parameter names of the reference's `_create_fcnn` (Linear layers at the same Sequential indices)."""
  from torch import nn
  layers = [nn.Dropout(input_dropout)] if input_dropout > 0 else []
  for i in range(len(sizes) - 2):
    layers.append(nn.Linear(sizes[i], sizes[i + 1]))
    if dropout > 0: layers.append(nn.Dropout(dropout))
    layers.append(nn.ReLU())
  layers.append(nn.Linear(sizes[-2], sizes[-1]))
  return nn.Sequential(*layers)


def _cli(tmp_path, *args):
  cmd = [sys.executable, 'train.py', *args, f'output_dir={tmp_path}']
  res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
  assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
  runs = glob.glob(os.path.join(str(tmp_path), '*', '*'))
  assert len(runs) == 1, runs
  return runs[0], res.stdout


def test_gail_command_line_writes_the_reference_outputs(tmp_path):
  S, A, H = 12, 3, 32  # hopper with absorbing-state indicator: 11 + 1
  out, stdout = _cli(tmp_path, 'algorithm=GAIL', 'env=hopper', 'steps=120', 'training.start=40', 'training.batch_size=32', 'training.learning_rate=3e-4', 'evaluation.interval=60',
                     'evaluation.episodes=2', 'logging.interval=20', 'save_trajectories=true', 'imitation.trajectories=2', 'memory.size=200', f'reinforcement.actor.hidden_size={H}',
                     f'reinforcement.critic.hidden_size={H}', 'imitation.discriminator.hidden_size=32')
  assert 'step 60: test return' in stdout and 'step 120: test return' in stdout
  assert sorted(os.listdir(out)) == ['agent.pth', 'discriminator.pth', 'metrics.pth', 'trajectories.pth']
  agent = torch.load(os.path.join(out, 'agent.pth'))
  assert set(agent) == {'actor', 'critic', 'log_alpha'}  # train.py:237
  actor = torch.nn.Module()
  actor.actor = _fcnn([S, H, H, 2 * A])
  actor.load_state_dict(agent['actor'])  # strict: exactly the reference's keys / shapes
  critic = torch.nn.Module()
  critic.critic_1, critic.critic_2 = torch.nn.Module(), torch.nn.Module()
  critic.critic_1.critic, critic.critic_2.critic = _fcnn([S + A, H, H, 1]), _fcnn([S + A, H, H, 1])
  critic.load_state_dict(agent['critic'])
  assert all(torch.isfinite(v).all() for v in agent['actor'].values())
  disc = torch.load(os.path.join(out, 'discriminator.pth'))
  assert any(k.endswith('parametrizations.weight.original') for k in disc), list(disc)  # spectral_norm parametrization keys (models.py:65)
  m = torch.load(os.path.join(out, 'metrics.pth'), weights_only=False)
  assert m['test_steps'] == [60, 120] and len(m['test_returns']) == 2 and len(m['test_returns'][0]) == 2
  assert m['update_steps'] == [40, 60, 80, 100, 120]  # train.py:205: only steps inside the update branch
  for k in ('predicted_rewards', 'alphas', 'entropies', 'Q_values'): assert len(m[k]) == 5
  traj = torch.load(os.path.join(out, 'trajectories.pth'), weights_only=False)
  assert len(traj) == 2  # evaluation.episodes dicts (evaluation.py:30-33)
  for t in traj:
    L = len(t['rewards'])
    assert t['states'].shape == (L, S) and t['actions'].shape == (L, A) and t['terminals'].shape == (L, )
    assert t['terminals'][-1] == 1 and t['terminals'][:-1].sum() == 0
    assert float(t['actions'].abs().max()) <= 1.0


@pytest.mark.parametrize('algorithm,extra', [('RED', ('imitation.pretraining.iterations=20', )), ('DRIL', ('imitation.pretraining.iterations=20', )), ('AdRIL', ())])
def test_other_algorithms_run_from_the_command_line(tmp_path, algorithm, extra):
  out, stdout = _cli(tmp_path, f'algorithm={algorithm}', 'env=halfcheetah', 'steps=80', 'training.start=40', 'training.batch_size=32', 'evaluation.interval=80', 'evaluation.episodes=1',
                     'logging.interval=0', 'imitation.trajectories=2', 'memory.size=400', 'reinforcement.actor.hidden_size=32', 'reinforcement.critic.hidden_size=32', *extra)
  assert 'step 80: test return' in stdout
  files = sorted(os.listdir(out))
  assert files == (['agent.pth', 'discriminator.pth', 'metrics.pth'] if algorithm in ('RED', 'DRIL') else ['agent.pth', 'metrics.pth'])
  if algorithm == 'DRIL':  # DRIL's "discriminator" is a dropout SoftActor: Dropout(0) Linear(1) Dropout(2) Tanh(3) Linear(4) (models.py:51-61, DRIL.yaml)
    disc = torch.load(os.path.join(out, 'discriminator.pth'))
    ref = torch.nn.Module()
    ref.actor = _fcnn([18, 64, 2 * 6], input_dropout=0.1, dropout=0.1)
    ref.load_state_dict(disc)
  m = torch.load(os.path.join(out, 'metrics.pth'), weights_only=False)
  assert np.isfinite(np.asarray(m['test_returns'])).all()
