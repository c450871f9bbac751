"""CPU-only checks of the drop-in boundary: the shared library builds/loads, exports every entry point that
include/il_b200.h declares, the ctypes struct mirrors have the C sizes, and layout helpers agree."""
import ctypes as C
import os
import re

import pytest

import il_b200
from il_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, 'include', 'il_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(il_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  lib = _lib.lib()
  names = _declared()
  assert len(names) >= 30
  for n in names:
    assert hasattr(lib, n), f'{n} declared in include/il_b200.h but not exported'
  assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_struct_sizes_match_c():
  out = (C.c_int32 * 15)()
  assert _lib.lib().il_struct_sizes(out) == 0
  py = [C.sizeof(x) for x in (_lib.Mlp, _lib.Adam, _lib.Batch, _lib.Replay, _lib.SacArgs, _lib.Gail, _lib.GailUpdateArgs, _lib.Pwil, _lib.Env, _lib.BcArgs, _lib.EvalArgs, _lib.Gailx, _lib.GailxUpdateArgs, _lib.Red, _lib.RedUpdateArgs)]
  assert list(out) == py


@pytest.mark.parametrize('dims', [[12, 256, 256, 6], [15, 256, 256, 1], [120, 64, 1], [7, 33, 5]])
def test_mlp_layout_mirror(dims):
  assert _lib.mlp_offsets(dims) == _lib.py_mlp_offsets(dims)


@pytest.mark.parametrize('S,A', [(12, 3), (18, 6), (112, 8), (5, 2)])
def test_row_layout_mirror(S, A):
  assert _lib.row_layout(S, A) == _lib.py_row_layout(S, A)
  assert _lib.py_row_layout(S, A)[1] % 4 == 0


def test_no_cpu_fallback():
  import torch
  if torch.cuda.is_available(): pytest.skip('GPU present')
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    _lib.handle()
  with pytest.raises(RuntimeError):
    il_b200.SoftActor(12, 3, type('C', (), dict(hidden_size=8, depth=1, activation='relu', get=lambda self, k, d=None: d))())


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, 'imitation-learning_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        txt = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f'{f} imports oracle/'
