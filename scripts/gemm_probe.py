"""Times the dense-layer grouped GEMM (the dominant kernel) in isolation for the three operand layouts and engines.
Usage: python scripts/gemm_probe.py [G] [mode ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import il_b200
from il_b200 import _lib

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
modes = sys.argv[2:] or ['fp32', 'tf32x3', 'tf32']
M = N = K = 256
lib, h = _lib.lib(), _lib.handle()
X = torch.randn(G, M, K, device='cuda')
W = torch.randn(G, N, K, device='cuda') / 16
Cm = torch.empty(G, M, N, device='cuda')
bias = torch.randn(G, N, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')  # 256 MB > L2


def run(mode, layout, iters=5):
  _lib.check(lib.il_set_gemm_mode(h, _lib.GEMM_MODE[mode]))
  ak, bk = {'fwd': (1, 1), 'dx': (1, 0), 'dw': (0, 0)}[layout]
  ts = []
  for i in range(iters + 2):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.il_debug_gemm(h, M, N, K, G, X.data_ptr(), X.stride(0), K, ak, W.data_ptr(), W.stride(0), K, bk, Cm.data_ptr(), Cm.stride(0), N, bias.data_ptr() if layout == 'fwd' else None,
                                 N, 0 if layout == 'fwd' else -1, None, 0, N, 0, None, 0, _lib.stream()))
    e1.record()
    torch.cuda.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1))
  ms = sorted(ts)[len(ts) // 2]
  print(f'{mode:7s} {layout:4s} G={G}: {ms:.3f} ms  {2.0 * M * N * K * G / ms / 1e9:.1f} TFLOP/s', flush=True)


for mode in modes:
  for layout in ('fwd', 'dx', 'dw'): run(mode, layout)
