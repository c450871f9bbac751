"""Pins the oracle port (oracle/port.py): (1) against the committed fixtures generated from the unmodified
reference (tests/golden/*.npz, oracle/make_golden.py); (2) when /root/reference is present (build container),
against the reference executed live on the same seeded inputs. CPU only."""
import pytest

from conftest import load_golden
from oracle import cases, refstub

RTOL, ATOL = 2e-5, 2e-6  # fp32: port and reference run the same torch CPU ops; slack covers thread-count summation order


@pytest.mark.parametrize('name', list(cases.CASES))
def test_port_matches_golden(name):
  out = cases.run_port(name, cases.make_inputs(name))
  bad = cases.compare(load_golden(name), out, rtol=RTOL, atol=ATOL)
  assert not bad, '\n'.join(bad)


@pytest.mark.skipif(not refstub.available(), reason='reference tree not present (GPU box)')
@pytest.mark.parametrize('name', ['actor_small', 'sac_small', 'bc_small', 'gail_default', 'gail_mixup', 'gmmil_hopper', 'pwil_small', 'replay_ring',
                                  'gailx_shaping', 'gailx_depth2_tanh', 'gailx_state_only_sigmoid'])
def test_port_matches_live_reference(name):
  inp = cases.make_inputs(name)
  ref = cases.run_reference(name, inp)
  bad = cases.compare(cases.compress(ref), cases.run_port(name, inp), rtol=RTOL, atol=ATOL)
  assert not bad, '\n'.join(bad)


def test_replay_index_stream_is_reference_stream():
  """memory.py:51-56 draws one np.random.randint per index from the GLOBAL stream; the port must consume it identically."""
  import numpy as np
  g = load_golden('replay_ring')
  out = cases.run_port('replay_ring', cases.make_inputs('replay_ring'))
  assert np.array_equal(g['sample_step'], out['sample_step'])
  assert np.array_equal(g['meta'], out['meta'])
