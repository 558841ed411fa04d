"""Host-side checks of the scalar arithmetic the CUDA kernels share through csrc/box_math.h (compiled here with
gcc): cost-matrix entries vs the reference-generated golden, the analytic box-loss gradient vs torch autograd of
the oracle, and the warp-parallel LSAP tie rule (re-enacted serially) vs scipy."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import detr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("harness") / "libharness.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_harness.c"), "-lm"])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_cost_entries_vs_reference_golden(harness, golden):
    g = golden
    for b in range(g["m_logits"].shape[0]):
        n = int(g["m_t_bbox"][b, 0, 0])
        probs = torch.softmax(torch.from_numpy(g["m_logits"][b]), -1).numpy().astype(np.float32)
        pb = np.ascontiguousarray(g["m_boxes"][b])
        tb = np.ascontiguousarray(g["m_t_bbox"][b, 1:1 + n])
        tc = np.ascontiguousarray(g["m_t_class"][b, 1:1 + n, 0])
        out = np.zeros((100, n), np.float32)
        harness.h_cost_matrix(_p(pb), _p(np.ascontiguousarray(probs)), 100, 92, _p(tb), _p(tc), n,
                              ctypes.c_float(1), ctypes.c_float(5), ctypes.c_float(2), _p(out))
        np.testing.assert_allclose(out, g[f"m_cost_{b}"], rtol=0, atol=1e-6)


def test_box_loss_grad_vs_autograd(harness):
    rs = np.random.RandomState(3)
    for it in range(300):
        p = np.concatenate([rs.uniform(0.05, 0.95, 2), rs.uniform(0.02, 0.6, 2)]).astype(np.float32)
        t = np.concatenate([rs.uniform(0.1, 0.9, 2), rs.uniform(0.02, 0.5, 2)]).astype(np.float32)
        if it % 5 == 0:
            p[2:] *= 2.5          # exercise the [0,1] clip
        l1 = ctypes.c_float()
        gl = ctypes.c_float()
        grad = np.zeros(4, np.float32)
        harness.h_box_loss_grad(_p(p), _p(t), ctypes.c_float(5.0), ctypes.c_float(2.0), ctypes.byref(l1),
                                ctypes.byref(gl), _p(grad))
        pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        tt = torch.tensor(t, dtype=torch.float64)
        l1_t = (pt - tt).abs().sum()
        giou = O._diag_giou(O.xcycwh_to_xy_min_xy_max(pt[None]), O.xcycwh_to_xy_min_xy_max(tt[None]))[0]
        (5 * l1_t + 2 * (1 - giou)).backward()
        assert abs(l1.value - float(l1_t.detach())) < 1e-5
        assert abs(gl.value - float((1 - giou).detach())) < 1e-5
        np.testing.assert_allclose(grad, pt.grad.numpy(), rtol=2e-3, atol=2e-4)


def test_warp_parallel_lsap_rule_vs_scipy(harness):
    from scipy.optimize import linear_sum_assignment
    rs = np.random.RandomState(11)
    for it in range(3000):
        Q = 100 if it % 2 else rs.randint(1, 128)
        n = rs.randint(0, min(Q - 1, 99) + 1)   # n < Q: scipy solves the transposed problem (always true for DETR: n <= 99 < Q = 100)
        if it % 3 == 0:
            c = rs.rand(Q, n).astype(np.float32)
        elif it % 3 == 1:
            c = rs.randint(0, 3, (Q, n)).astype(np.float32)
        else:
            c = (rs.randint(0, 2, (Q, n)) * 0.5).astype(np.float32) + rs.randint(0, 2, (Q, 1)).astype(np.float32)
        costT = np.ascontiguousarray(c.T)
        r4c = np.zeros(Q, np.int32)
        rc = harness.lsap_warp_model(_p(costT), n, Q, _p(r4c))
        assert rc == 0
        rows, cols = linear_sum_assignment(c)
        exp = -np.ones(Q, np.int32)
        exp[rows] = cols
        assert np.array_equal(r4c, exp), (Q, n, it)


def test_attention_dropout_rng():
    """Statistics of the attention-probability dropout RNG (csrc/common.cuh attn_drop_word, mirrored by
    cabi_emulator.attn_keep_mask): drop rate = round(p * 2^15) / 2^15, row / column marginals binomial, no serial or
    cross-row correlation beyond sampling noise, masks change with the seed."""
    import numpy as np
    import cabi_emulator as E
    R, K, p = 2048, 1056, 0.1
    keep = E.attn_keep_mask(range(1000, 1000 + R), range(K), p, 0x12345678_00000063, 11).numpy()
    d = (~keep).astype(np.float64)
    assert abs(d.mean() - 3277 / 32768) < 4 * np.sqrt(0.09 / (R * K))
    assert abs(d.mean(1).std() / np.sqrt(0.09 / K) - 1) < 0.1 and abs(d.mean(0).std() / np.sqrt(0.09 / R) - 1) < 0.1

    def corr(a, b):
        a, b = a - a.mean(), b - b.mean()
        return (a * b).mean() / np.sqrt((a * a).mean() * (b * b).mean())
    tol = 5 / np.sqrt(R * K)
    for lag in (1, 2, 7, 8, 9, 16, 64, 128):
        assert abs(corr(d[:, :-lag], d[:, lag:])) < tol, ("key lag", lag)
    for lag in (1, 2, 8, 64):
        assert abs(corr(d[:-lag], d[lag:])) < tol, ("row lag", lag)
    other = E.attn_keep_mask(range(1000, 1000 + R), range(K), p, 0x12345678_00000064, 11).numpy()
    assert abs(corr(d, (~other).astype(np.float64))) < tol          # next step's seed: independent mask
    assert E.attn_keep_mask(range(4), range(64), 0.0, 1, 1).all()
