"""Oracle (test infrastructure only): SO(3)/O(3) bookkeeping used by e3nn-jax.

Restates, from the published e3nn-jax algorithm (SURVEY.md Appendix A.1-A.3;
the package itself is absent from /root/reference and from this image):
  * ``Irreps`` parsing / regroup / simplify / filter        (A.1)
  * real Clebsch-Gordan coefficients ``clebsch_gordan``      (A.2)
  * real spherical harmonics ``spherical_harmonics``         (A.3)
Call sites in the reference that fix the semantics needed:
  models/nequip.py:43-52,85-88,162,180-197 and models/utils/irreps_utils.py:6-25.

The su(2) coefficients come from sympy (independent of the product's own
Racah implementation in eqnn_jax_b200/so3.py, so the two can be diffed).
"""
from __future__ import annotations

import functools
import re
from typing import Iterable, List, Sequence, Tuple, Union

import numpy as np


# ----------------------------------------------------------------------------
# Irreps
# ----------------------------------------------------------------------------
def _irrep_key(l: int, p: int):
    # e3nn ordering: 0e,0o,1o,1e,2e,2o,3o,3e,...
    return (l, -p * (-1) ** l)


class Irreps:
    """Ordered list of (mul, l, p); p=+1 even 'e', p=-1 odd 'o'."""

    def __init__(self, spec: Union[str, "Irreps", Sequence[Tuple[int, int, int]]] = ()):
        if isinstance(spec, Irreps):
            self.items: List[Tuple[int, int, int]] = list(spec.items)
        elif isinstance(spec, str):
            self.items = []
            if spec.strip():
                for tok in spec.split("+"):
                    tok = tok.strip()
                    m = re.fullmatch(r"(?:(\d+)\s*x\s*)?(\d+)([eo])", tok)
                    if m is None:
                        raise ValueError(f"bad irrep {tok!r}")
                    mul = int(m.group(1)) if m.group(1) else 1
                    self.items.append((mul, int(m.group(2)), 1 if m.group(3) == "e" else -1))
        else:
            self.items = [(int(m), int(l), int(p)) for (m, l, p) in spec]

    @staticmethod
    def spherical_harmonics(lmax: int) -> "Irreps":
        return Irreps([(1, l, (-1) ** l) for l in range(lmax + 1)])

    # -- basic properties
    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)

    def __eq__(self, other):
        try:
            other = Irreps(other)
        except Exception:
            return False
        return self.items == other.items

    def __add__(self, other):
        return Irreps(self.items + Irreps(other).items)

    def __repr__(self):
        return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in self.items)

    @property
    def dim(self) -> int:
        return sum(m * (2 * l + 1) for m, l, _ in self.items)

    @property
    def num_irreps(self) -> int:
        return sum(m for m, _, _ in self.items)

    def count(self, ir: str) -> int:
        (_, l, p), = Irreps(ir).items
        return sum(m for m, l2, p2 in self.items if (l2, p2) == (l, p))

    def slices(self) -> List[slice]:
        out, o = [], 0
        for m, l, _ in self.items:
            out.append(slice(o, o + m * (2 * l + 1)))
            o += m * (2 * l + 1)
        return out

    # -- transformations
    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(m, l, p) for m, l, p in self.items if m > 0])

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, int, int]] = []
        for m, l, p in self.items:
            if m == 0:
                continue
            if out and out[-1][1:] == (l, p):
                out[-1] = (out[-1][0] + m, l, p)
            else:
                out.append((m, l, p))
        return Irreps(out)

    def sort_perm(self) -> List[int]:
        """Stable permutation that sorts the chunks in e3nn order."""
        return sorted(range(len(self.items)), key=lambda i: _irrep_key(*self.items[i][1:]))

    def regroup(self) -> "Irreps":
        perm = self.sort_perm()
        return Irreps([self.items[i] for i in perm]).simplify()

    def filter(self, keep: Union[str, "Irreps", Iterable] = None, drop=None) -> "Irreps":
        if keep is not None:
            ks = {(l, p) for _, l, p in Irreps(keep).items}
            return Irreps([(m, l, p) for m, l, p in self.items if (l, p) in ks])
        ds = {(l, p) for _, l, p in Irreps(drop).items}
        return Irreps([(m, l, p) for m, l, p in self.items if (l, p) not in ds])


def balanced_irreps(lmax: int, feature_size: int, use_sh: bool = True) -> Irreps:
    """Restates models/utils/irreps_utils.py:6-25, including the quirk that
    ``total_dim`` is overwritten (not accumulated) at :18 and :21."""
    irreps = [(0, 0, 1)]
    n_irreps = 1 + (lmax if use_sh else lmax * 2)
    total_dim = 0
    for level in range(1, lmax + 1):
        dim = 2 * level + 1
        multi = int(feature_size / dim / n_irreps)
        if multi == 0:
            break
        if use_sh:
            irreps.append((multi, level, 1 if level % 2 == 0 else -1))
            total_dim = multi * dim
        else:
            irreps.append((multi, level, 1))
            irreps.append((multi, level, -1))
            total_dim = multi * dim * 2
    irreps[0] = (feature_size - total_dim, 0, 1)
    return Irreps(irreps)


# ----------------------------------------------------------------------------
# Clebsch-Gordan (A.2)
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def su2_clebsch_gordan(j1: int, j2: int, j3: int) -> np.ndarray:
    """<j1 m1 j2 m2 | j3 m3>, Condon-Shortley, via sympy."""
    from sympy import N
    from sympy.physics.quantum.cg import CG

    out = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            m3 = m1 + m2
            if abs(m3) <= j3:
                out[j1 + m1, j2 + m2, j3 + m3] = float(N(CG(j1, m1, j2, m2, j3, m3).doit(), 30))
    return out


def change_basis_real_to_complex(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / np.sqrt(2)
        q[l + m, l - abs(m)] = -1j / np.sqrt(2)
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / np.sqrt(2)
        q[l + m, l - abs(m)] = 1j * (-1) ** m / np.sqrt(2)
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def clebsch_gordan(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real CG tensor (2l1+1, 2l2+1, 2l3+1), Frobenius norm 1 (A.2)."""
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        return np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    C = su2_clebsch_gordan(l1, l2, l3)
    Q1, Q2, Q3 = (change_basis_real_to_complex(l) for l in (l1, l2, l3))
    C = np.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, np.conj(Q3.T), C)
    assert np.all(np.abs(C.imag) < 1e-10)
    C = C.real
    C = C / np.linalg.norm(C)
    C[np.abs(C) < 1e-14] = 0.0
    return C


# ----------------------------------------------------------------------------
# Spherical harmonics (A.3)
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _sh_recursion_constant(l: int) -> float:
    """c_l > 0 such that Y_l = c_l * C(l-1,1,l) . Y_{l-1} . Y_1 has unit norm."""
    v = np.array([0.3, -0.5, 0.81])
    v = v / np.linalg.norm(v)
    ys = _sh_norm_numpy(l - 1, v[None])[-1]
    raw = np.einsum("ijk,ni,nj->nk", clebsch_gordan(l - 1, 1, l), ys, v[None])
    return float(1.0 / np.linalg.norm(raw[0]))


def _sh_norm_numpy(lmax: int, u: np.ndarray) -> List[np.ndarray]:
    ys = [np.ones(u.shape[:-1] + (1,))]
    if lmax >= 1:
        ys.append(u.copy())
    for l in range(2, lmax + 1):
        c = _sh_recursion_constant(l)
        ys.append(c * np.einsum("ijk,...i,...j->...k", clebsch_gordan(l - 1, 1, l), ys[-1], u))
    return ys


def sh_normalization_factor(l: int, normalization: str) -> float:
    if normalization == "norm":
        return 1.0
    if normalization == "component":
        return float(np.sqrt(2 * l + 1))
    if normalization == "integral":
        return float(np.sqrt((2 * l + 1) / (4 * np.pi)))
    raise ValueError(normalization)


def spherical_harmonics(lmax: int, vec, normalize: bool = True, normalization: str = "component"):
    """Y_lm for l=0..lmax concatenated, e3nn convention (Y_1 ~ (x,y,z)).

    Works on numpy arrays or torch tensors (any float dtype); `normalize`
    divides by |r| with e3nn's guard (|r|^2==0 -> divide by 1)."""
    import torch

    is_np = isinstance(vec, np.ndarray)
    v = torch.as_tensor(vec)
    if normalize:
        n2 = (v * v).sum(-1, keepdim=True)
        n = torch.sqrt(torch.where(n2 == 0, torch.ones_like(n2), n2))
        u = v / n
    else:
        u = v
    ys = [torch.ones(u.shape[:-1] + (1,), dtype=u.dtype)]
    if lmax >= 1:
        ys.append(u)
    for l in range(2, lmax + 1):
        C = torch.as_tensor(clebsch_gordan(l - 1, 1, l), dtype=u.dtype)
        ys.append(_sh_recursion_constant(l) * torch.einsum("ijk,...i,...j->...k", C, ys[-1], u))
    out = torch.cat([sh_normalization_factor(l, normalization) * y for l, y in enumerate(ys)], dim=-1)
    return out.numpy() if is_np else out


# ----------------------------------------------------------------------------
# Wigner-D (test helper): D^l(R) with Y_l(R x) = D^l(R) Y_l(x)
# ----------------------------------------------------------------------------
def wigner_D_from_sh(l: int, R: np.ndarray, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(8 * (2 * l + 1), 3))
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    y = _sh_norm_numpy(l, x)[l]
    yr = _sh_norm_numpy(l, x @ R.T)[l]
    D, *_ = np.linalg.lstsq(y, yr, rcond=None)
    return D.T


def irreps_rotation_matrix(irreps: Irreps, R: np.ndarray) -> np.ndarray:
    """Block-diagonal representation matrix of a proper/improper rotation R."""
    det = np.sign(np.linalg.det(R))
    Rp = R * det  # proper part
    D = np.zeros((irreps.dim, irreps.dim))
    o = 0
    for mul, l, p in irreps:
        Dl = wigner_D_from_sh(l, Rp) * (p if det < 0 else 1)
        for _ in range(mul):
            D[o:o + 2 * l + 1, o:o + 2 * l + 1] = Dl
            o += 2 * l + 1
    return D
