"""GPU: randomized degenerate inputs (few distinct rows, extreme byte values, exact duplicates, one or
two targets) through every dense path and the grouped kernels, against the C oracle."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings

import oracle
from fast_match_b200 import backend
from test_oracle_properties import problems

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(problems())
def test_dense_paths_on_degenerate_inputs(cuda, p):
    q, t, base = p
    od2, oidx = oracle.c_top2(q, t, base)
    for algo in (backend.FM_ALGO_MMA_SYNC, backend.FM_ALGO_TCGEN05):
        d2, idx, keys = backend.top2(_dev(q), _dev(t), t_index_base=base, want_keys=True, algo=algo)
        assert np.array_equal(d2.cpu().numpy().view(np.uint32), od2), algo
        assert np.array_equal(idx.cpu().numpy(), oidx), algo
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.pack_keys(od2, oidx)), algo


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(problems())
def test_grouped_paths_on_degenerate_inputs(cuda, p):
    q, t, _ = p
    # three groups: (q, t), (t, q) and an empty-sided one
    qp = np.concatenate([q, t, q])
    tp = np.concatenate([t, q])
    q_off = np.array([0, len(q), len(q) + len(t), 2 * len(q) + len(t)], np.int64)
    t_off = np.array([0, len(t), len(t) + len(q), len(t) + len(q)], np.int64)
    od2, oidx, ot2q = oracle.c_grouped_mutual(qp, q_off, tp, t_off)
    for algo in (backend.FM_ALGO_MMA_SYNC, backend.FM_ALGO_TCGEN05):
        d2, idx, t2q, _ = backend.grouped_mutual(_dev(qp), _dev(q_off), _dev(tp), _dev(t_off), algo=algo)
        assert np.array_equal(d2.cpu().numpy().view(np.uint32), od2), algo
        assert np.array_equal(idx.cpu().numpy(), oidx), algo
        assert np.array_equal(t2q.cpu().numpy(), ot2q), algo
