"""Real Clebsch-Gordan tables and spherical-harmonic recursion constants for the
layer-plan compiler / kernel generator (host side, build and trace time only).

Conventions are those of e3nn-jax, which the reference calls at
models/nequip.py:46-52 (``e3nn.spherical_harmonics``, ``tensor_product``):
  * real basis obtained from the complex one by the change of basis ``Q_l`` with
    the extra ``(-i)^l`` phase, CG normalised to unit Frobenius norm;
  * ``Y_1`` is the input vector itself (x, y, z), higher ``l`` by the recursion
    ``Y_l = c_l * C(l-1,1,l) . Y_{l-1} . Y_1`` with ``c_l > 0`` fixing |Y_l| = 1.
The su(2) coefficients use Racah's closed form in exact rational arithmetic.
"""
from __future__ import annotations

import functools
from fractions import Fraction
from math import factorial, sqrt
from typing import List, Tuple

import numpy as np


def _racah(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    if m3 != m1 + m2 or not (abs(j1 - j2) <= j3 <= j1 + j2):
        return 0.0
    f = factorial
    pref = Fraction((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3), f(j1 + j2 + j3 + 1))
    pref *= f(j3 + m3) * f(j3 - m3) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)
    s = Fraction(0)
    for k in range(0, j1 + j2 - j3 + 1):
        args = (k, j1 + j2 - j3 - k, j1 - m1 - k, j2 + m2 - k, j3 - j2 + m1 + k, j3 - j1 - m2 + k)
        if min(args) < 0:
            continue
        den = 1
        for a in args:
            den *= f(a)
        s += Fraction((-1) ** k, den)
    return float(s) * sqrt(float(pref))


def _real_to_complex(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    r = 1.0 / sqrt(2.0)
    for m in range(1, l + 1):
        q[l - m, l + m] = r
        q[l - m, l - m] = -1j * r
        q[l + m, l + m] = (-1) ** m * r
        q[l + m, l - m] = 1j * (-1) ** m * r
    q[l, l] = 1.0
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def real_cg(l1: int, l2: int, l3: int) -> np.ndarray:
    """(2l1+1, 2l2+1, 2l3+1) real coupling tensor, sum of squares 1."""
    c = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    for a in range(-l1, l1 + 1):
        for b in range(-l2, l2 + 1):
            if abs(a + b) <= l3:
                c[l1 + a, l2 + b, l3 + a + b] = _racah(l1, a, l2, b, l3, a + b)
    q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    t = np.einsum("ij,kl,mn,ikn->jlm", q1, q2, np.conj(q3.T), c)
    if np.abs(t.imag).max() > 1e-9:
        raise AssertionError("coupling tensor is not real")
    t = t.real
    nrm = np.linalg.norm(t)
    if nrm > 0:
        t = t / nrm
    t[np.abs(t) < 1e-13] = 0.0
    return t


def _sh_unit(lmax: int, u: np.ndarray) -> List[np.ndarray]:
    ys = [np.ones(u.shape[:-1] + (1,)), u][: lmax + 1]
    for l in range(2, lmax + 1):
        ys.append(sh_recursion_constant(l) * np.einsum("ijk,...i,...j->...k", real_cg(l - 1, 1, l), ys[-1], u))
    return ys


@functools.lru_cache(maxsize=None)
def sh_recursion_constant(l: int) -> float:
    u = np.array([[0.36, -0.48, 0.8]])
    prev = _sh_unit(l - 1, u)[-1]
    raw = np.einsum("ijk,ni,nj->nk", real_cg(l - 1, 1, l), prev, u)
    return float(1.0 / np.linalg.norm(raw))


def sh_norm_factor(l: int, normalization: str) -> float:
    """Factor applied on top of the unit-norm harmonics (e3nn ``normalization=``)."""
    if normalization == "norm":
        return 1.0
    if normalization == "component":
        return sqrt(2 * l + 1)
    if normalization == "integral":
        return sqrt((2 * l + 1) / (4.0 * np.pi))
    raise ValueError(f"unknown spherical-harmonics normalization {normalization!r}")


def sh_recursion_terms(l: int) -> List[Tuple[int, int, int, float]]:
    """Non-zeros (i, j, k, coef) of  Y_l[k] = sum coef * Y_{l-1}[i] * u[j]."""
    t = real_cg(l - 1, 1, l) * sh_recursion_constant(l)
    return [(int(i), int(j), int(k), float(t[i, j, k])) for i, j, k in zip(*np.nonzero(t))]


def cg_terms(l1: int, l2: int, l3: int) -> List[Tuple[int, int, int, float]]:
    """Non-zeros (i, j, k, coef) of the 'component'-normalised coupling sqrt(2 l3+1) C."""
    t = real_cg(l1, l2, l3) * sqrt(2 * l3 + 1)
    return [(int(i), int(j), int(k), float(t[i, j, k])) for i, j, k in zip(*np.nonzero(t))]
