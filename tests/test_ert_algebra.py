"""The algebra behind early ray termination, on the CPU (numpy, fp64), against FIT_eval_transfer's update rule
(field_interpolation_tables.pxd:103-110):

    grey:      ta = max(1 - dt * T_alpha, 0);  C_c <- dt * T_c + ta * C_c          (all four channels)
    non-grey:  ta_c = max(1 - dt * T_c, 0);    C_c <- dt * T_c + ta_c * C_c        (channel by channel)

* processed from the other end as affine maps, A <- A + B e, B <- B a, the samples give the same colour,
  and cutting the walk once B < eps changes it by at most eps times the colour the rest could reach;
* without grey opacity a channel is absorbed by its own emission, so the composition is order-free (1 - C is a plain
  product) and there is nothing to terminate early unless all three channels saturate.
"""
import numpy as np


def _forward(e, a, c0):
    c = np.array(c0, dtype=np.float64)
    for ek, ak in zip(e, a):
        c = ek + ak * c
    return c


def _from_the_far_end(e, a, c0, eps=0.0):
    A = np.zeros_like(np.asarray(c0, dtype=np.float64))
    B = np.ones_like(A)
    used = 0
    for ek, ak in zip(e[::-1], a[::-1]):
        if eps > 0.0 and B.max() < eps:
            break
        A = A + B * ek
        B = B * ak
        used += 1
    return A + B * np.asarray(c0), used


def test_reversed_affine_composition_equals_the_reference_order():
    rng = np.random.default_rng(0)
    n = 400
    dt = rng.uniform(1e-4, 3e-3, size=n)
    T = rng.uniform(0.0, 40.0, size=(n, 4))
    c0 = rng.uniform(0, 0.3, size=4)
    # grey opacity: one transmittance for all channels
    ta = np.maximum(1.0 - dt * T[:, 3], 0.0)
    e, a = dt[:, None] * T, np.repeat(ta[:, None], 4, axis=1)
    want = _forward(e, a, c0)
    got, used = _from_the_far_end(e, a, c0)
    assert used == n and np.allclose(got, want, rtol=1e-12, atol=1e-15)


def test_termination_error_is_bounded_by_the_transmittance():
    rng = np.random.default_rng(1)
    n = 3000
    dt = np.full(n, 2e-3)
    T = rng.uniform(0.0, 1.0, size=(n, 4))
    T[:, 3] *= 400.0                                   # an opaque alpha table
    c0 = rng.uniform(0, 0.3, size=4)
    ta = np.maximum(1.0 - dt * T[:, 3], 0.0)
    e, a = dt[:, None] * T, np.repeat(ta[:, None], 4, axis=1)
    want = _forward(e, a, c0)
    for eps in (1e-3, 1e-6):
        got, used = _from_the_far_end(e, a, c0, eps)
        assert used < n                                # the walk really stops early
        brightest = max(_forward(e[:n - used], a[:n - used], c0).max(), 1e-300)   # what the skipped part would have shown
        assert np.abs(got - want).max() <= eps * brightest * (1 + 1e-9)


def test_without_grey_opacity_the_composition_is_order_free():
    rng = np.random.default_rng(2)
    n = 500
    dt = rng.uniform(1e-4, 2e-3, size=n)
    T = rng.uniform(0.0, 300.0, size=(n, 3))
    assert (dt[:, None] * T < 1.0).all()               # no clamp fires
    c0 = rng.uniform(0, 0.5, size=3)
    e = dt[:, None] * T
    a = 1.0 - e
    want = _forward(e, a, c0)
    perm = rng.permutation(n)
    assert np.allclose(_forward(e[perm], a[perm], c0), want, rtol=1e-12)
    assert np.allclose(1.0 - want, (1.0 - c0) * np.prod(1.0 - e, axis=0), rtol=1e-12)
