"""Deterministic sin/cos/pow of the shading stage: the oracle's copy is pinned against numpy's
float64 functions (<= 1 ulp of the correctly rounded fp32 value); the GPU copy must equal the oracle's bit for
bit -- that identity is what makes whole images comparable bit for bit."""
import numpy as np
import pytest


def ulp_err(got, exact64):
    exact32 = exact64.astype(np.float32)
    ulp = np.spacing(np.abs(exact32)).astype(np.float64)
    ulp = np.maximum(ulp, np.float64(np.finfo(np.float32).tiny) * 2.0 ** -23)
    return np.abs(got.astype(np.float64) - exact64) / ulp


def sweep():
    rng = np.random.default_rng(9)
    ang = np.concatenate([np.linspace(0, 2 * np.pi, 200001), rng.random(200000) * 6.28318530718, [0.0, 6.2831855, 1.5707964, 3.1415927, 4.712389]]).astype(np.float32)
    px = np.concatenate([rng.random(300000), np.linspace(0, 1, 1001), [0.0, 1.0, 1e-30, 1e-38, 1e-42, 0.5]]).astype(np.float32)
    py = np.concatenate([1.0 / (rng.random(150000) * 10 + 1), rng.random(150000) * 10 + 0.3, np.full(1001, 1.0), [1.0, 0.5, 3.0, 2.0, 1.0, 10.0]]).astype(np.float32)
    return ang, px, py


def test_oracle_recipe_accuracy(cpu):
    ang, px, py = sweep()
    s, c = cpu.det_sincos(ang)
    a64 = ang.astype(np.float64)
    assert ulp_err(s, np.sin(a64)).max() <= 0.51 + 1e-3  # 0.5 ulp = correctly rounded; allow the rare double-rounding case
    assert ulp_err(c, np.cos(a64)).max() <= 0.51 + 1e-3
    p = cpu.det_pow(px, py)
    with np.errstate(all="ignore"):
        exact = np.power(px.astype(np.float64), py.astype(np.float64))
    ok = exact >= 1e-37  # ignore results in the fp32 denormal range for the ulp bound
    assert ulp_err(p[ok], exact[ok]).max() <= 1.0
    assert np.all((p[~ok] >= 0) & (p[~ok] < 2e-37))


def test_pow_special_cases(cpu):
    x = np.array([0, 0, 0, 1, 1, -1, np.nan, 0.5, np.inf, np.inf, 0.25], dtype=np.float32)
    y = np.array([1, 0, -1, 5, 0, 2, 1, np.nan, 2, -2, 0.5], dtype=np.float32)
    p = cpu.det_pow(x, y)
    exp = np.array([0, 1, np.inf, 1, 1, np.nan, np.nan, np.nan, np.inf, 0, 0.5], dtype=np.float32)
    assert np.array_equal(np.isnan(p), np.isnan(exp))
    assert np.array_equal(p[~np.isnan(exp)], exp[~np.isnan(exp)])


@pytest.mark.gpu
def test_gpu_copy_is_bit_identical_to_oracle(A, cpu):
    ang, px, py = sweep()
    gs, gc = A.debug_sincos(ang)
    os_, oc = cpu.det_sincos(ang)
    assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gc.view(np.uint32), oc.view(np.uint32))
    x = np.concatenate([px, [0, 0, 0, -1, np.nan, np.inf]]).astype(np.float32)
    y = np.concatenate([py, [1, 0, -1, 2, 1, -2]]).astype(np.float32)
    assert np.array_equal(A.debug_pow(x, y).view(np.uint32), cpu.det_pow(x, y).view(np.uint32))
