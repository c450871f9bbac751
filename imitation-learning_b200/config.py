"""Hydra-free loader for the reference's configuration surface (train.py:21-23, conf/): a base YAML, an
`algorithm=<ALG>` global overlay (`# @package _global_`), an optional `optimised_hyperparameters=<name>` overlay and
dotted `key=value` overrides — `python train.py algorithm=GAIL env=hopper training.batch_size=512`."""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import yaml

CONF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'conf')


class Config(dict):
  """Attribute-style nested dict (the subset of omegaconf.DictConfig the reference uses: attribute access, .get)."""

  def __getattr__(self, k):
    try:
      v = self[k]
    except KeyError:
      raise AttributeError(k)
    if isinstance(v, dict) and not isinstance(v, Config):
      v = Config(v)
      self[k] = v
    return v

  def __setattr__(self, k, v): self[k] = v

  def get(self, k, default=None):
    return getattr(self, k) if k in self else default


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
  for k, v in src.items():
    if isinstance(v, dict) and isinstance(dst.get(k), dict): _merge(dst[k], v)
    else: dst[k] = v
  return dst


def _parse(v: str) -> Any:
  """Override values the way Hydra's grammar reads them: ints, then floats (3e-4, 1e6, inf, nan — YAML 1.1 would hand
  these back as strings), then YAML for the rest (true / false / null / lists / quoted strings)."""
  t = v.strip()
  try:
    return int(t)
  except ValueError:
    pass
  try:
    return float(t)
  except ValueError:
    pass
  return yaml.safe_load(v)


def load_config(overrides: Optional[List[str]] = None, conf_dir: str = CONF_DIR) -> Config:
  overrides = list(overrides or [])
  with open(os.path.join(conf_dir, 'train_config.yaml')) as f: cfg = yaml.safe_load(f)
  groups = {'algorithm': cfg.get('algorithm', 'SAC'), 'optimised_hyperparameters': None}
  rest = []
  for o in overrides:
    k, _, v = o.partition('=')
    if k in groups: groups[k] = v
    else: rest.append((k, v))
  for g in ('algorithm', 'optimised_hyperparameters'):
    if groups[g] in (None, 'null', ''): continue
    path = os.path.join(conf_dir, g, f'{groups[g]}.yaml')
    if not os.path.exists(path):
      raise FileNotFoundError(f'no {g} config {groups[g]!r} under {conf_dir}')
    with open(path) as f: _merge(cfg, yaml.safe_load(f) or {})
  for k, v in rest:
    node = cfg
    parts = k.lstrip('+').split('.')
    for p in parts[:-1]: node = node.setdefault(p, {})
    node[parts[-1]] = _parse(v)
  return Config(cfg)
