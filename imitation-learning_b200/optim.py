"""torch.optim.Adam / AdamW mirrors whose state lives next to the flat replica parameter buffers
(train.py:66,84). The step itself is fused into the update kernels (csrc/mlp.cu adam_kernel: the arithmetic of
torch's `_single_tensor_adam`, hyper-parameters kept in double like the Python floats they are)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch
from torch import Tensor

from . import _lib


class Adam:
  decoupled = False

  def __init__(self, params: Iterable[Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
    params = list(params)
    assert len(params) == 1 and params[0].is_cuda, 'expects the single flat parameter tensor returned by module.parameters()'
    if weight_decay != 0 and not self.decoupled: raise NotImplementedError('coupled (L2) weight decay is not used by the reference (train.py:66)')
    self.param = params[0]
    self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
    self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.param), torch.zeros_like(self.param)
    self.step_count = torch.zeros(1, dtype=torch.int64, device=self.param.device)

  def c_struct(self) -> _lib.Adam:
    a = _lib.Adam()
    a.m, a.v, a.step = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.step_count.data_ptr()
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay
    return a

  def zero_grad(self, set_to_none: bool = True):
    pass  # gradients never persist between fused updates (training.py:29,40,47 zero them every call)

  def step_with(self, grads: Tensor):
    """One optimiser step from an explicit flat gradient (tests / custom losses)."""
    a = self.c_struct()
    _lib.check(_lib.lib().il_adam_step(_lib.handle(), self.param.data_ptr(), grads.contiguous().data_ptr(), C.byref(a), self.param.numel(), _lib.stream()))

  def state_dict(self):
    return dict(exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(), step=int(self.step_count.item()))


class AdamW(Adam):
  decoupled = True

  def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
    super().__init__(params, lr, betas, eps, weight_decay)
