"""TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference modules from /root/reference.

Only usable inside the build container (the GPU box has no /root/reference). Used by
`oracle/make_golden.py` to generate the fixtures in `tests/golden/` and by the `-m "not gpu"`
tests that pin `oracle/port.py` against the real reference. The import stub for `omegaconf`
follows SURVEY.md Appendix A; `gym`/`d4rl` stubs let `evaluation.py` import.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get('IL_REFERENCE_DIR', '/root/reference')


class DictConfig(dict):
  """Attribute-dict standing in for omegaconf.DictConfig (attribute access + .get)."""

  def __getattr__(self, k):
    try:
      v = self[k]
    except KeyError:
      raise AttributeError(k)
    return DictConfig(v) if isinstance(v, dict) and not isinstance(v, DictConfig) else v

  def __setattr__(self, k, v):
    self[k] = v


def available() -> bool:
  return os.path.isfile(os.path.join(REFERENCE_DIR, 'training.py'))


_cache = {}


def load():
  """Returns the reference modules (memory, models, training) or raises if unavailable."""
  if 'mods' in _cache:
    return _cache['mods']
  if not available():
    raise RuntimeError(f'reference not present at {REFERENCE_DIR}')
  if 'omegaconf' not in sys.modules:
    om = types.ModuleType('omegaconf')
    om.DictConfig, om.OmegaConf = DictConfig, object
    sys.modules['omegaconf'] = om
  for name in ('gym', 'gym.spaces', 'd4rl'):
    if name not in sys.modules:
      m = types.ModuleType(name)
      if name == 'gym':
        m.logger = types.SimpleNamespace(set_level=lambda *_: None)
      if name == 'gym.spaces':
        class Box:  # the attributes environments.py:23-27 and train.py:61 touch
          def __init__(self, low, high):
            import numpy as np
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape
        m.Box, m.Space = Box, object
      sys.modules[name] = m
  sys.modules['gym'].spaces = sys.modules['gym.spaces']
  if not hasattr(sys.modules['gym'], 'make'): sys.modules['gym'].make = lambda name: (_ for _ in ()).throw(RuntimeError('gym is stubbed; patch gym.make (oracle/ref_train.py)'))
  # The reference uses top-level module names (memory, models, training, ...); import them under
  # those names from REFERENCE_DIR without leaving it on sys.path permanently shadowing ours.
  saved = {k: sys.modules.pop(k) for k in ('memory', 'models', 'training', 'evaluation', 'environments') if k in sys.modules}
  sys.path.insert(0, REFERENCE_DIR)
  try:
    import memory, models, training, evaluation  # noqa: E401
    import environments
    mods = types.SimpleNamespace(memory=memory, models=models, training=training, evaluation=evaluation, environments=environments, DictConfig=DictConfig)
  finally:
    sys.path.remove(REFERENCE_DIR)
    for k in ('memory', 'models', 'training', 'evaluation', 'environments'):
      sys.modules.pop(k, None)
    sys.modules.update(saved)
  _cache['mods'] = mods
  return mods
