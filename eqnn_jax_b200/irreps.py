"""Irreps bookkeeping for the layer-plan compiler (host side, trace time only).

Mirrors the subset of ``e3nn_jax.Irreps`` the reference's hot path relies on
(reference call sites: models/nequip.py:43,81-85,124-127,162,195;
models/utils/irreps_utils.py:6-25).  Memory layout convention: the last axis of
a feature array is the concatenation of chunks, each chunk ``(mul, 2l+1)``
row-major.
"""
from __future__ import annotations

import re
from typing import List, Sequence, Tuple, Union

IrrepT = Tuple[int, int]          # (l, p)
MulIrT = Tuple[int, int, int]     # (mul, l, p)


def irrep_sort_key(l: int, p: int):
    """e3nn order 0e < 0o < 1o < 1e < 2e < 2o < 3o < 3e ..."""
    return (l, -p * (-1) ** l)


class Irreps(tuple):
    def __new__(cls, spec: Union[str, "Irreps", Sequence[MulIrT], None] = None):
        if spec is None:
            items: List[MulIrT] = []
        elif isinstance(spec, str):
            items = []
            for tok in filter(None, (t.strip() for t in spec.split("+"))):
                m = re.fullmatch(r"(?:(\d+)\s*x\s*)?(\d+)([eo])", tok)
                if m is None:
                    raise ValueError(f"cannot parse irrep {tok!r}")
                items.append((int(m.group(1) or 1), int(m.group(2)), 1 if m.group(3) == "e" else -1))
        else:
            items = [(int(a), int(b), int(c)) for a, b, c in spec]
        return super().__new__(cls, items)

    @classmethod
    def spherical_harmonics(cls, lmax: int) -> "Irreps":
        return cls([(1, l, (-1) ** l) for l in range(lmax + 1)])

    def __repr__(self) -> str:
        return "+".join(f"{m}x{l}{'e' if p > 0 else 'o'}" for m, l, p in self)

    __str__ = __repr__

    def __eq__(self, other) -> bool:
        try:
            return tuple(self) == tuple(Irreps(other))
        except Exception:
            return False

    def __ne__(self, other) -> bool:
        return not self.__eq__(other)

    def __hash__(self):
        return hash(tuple(self))

    def __add__(self, other) -> "Irreps":
        return Irreps(list(self) + list(Irreps(other)))

    @property
    def dim(self) -> int:
        return sum(m * (2 * l + 1) for m, l, _ in self)

    @property
    def num_irreps(self) -> int:
        return sum(m for m, _, _ in self)

    def count(self, ir) -> int:
        ((_, l, p),) = Irreps(ir)
        return sum(m for m, l2, p2 in self if (l2, p2) == (l, p))

    def offsets(self) -> List[int]:
        out, o = [], 0
        for m, l, _ in self:
            out.append(o)
            o += m * (2 * l + 1)
        return out

    def simplify(self) -> "Irreps":
        out: List[MulIrT] = []
        for m, l, p in self:
            if m == 0:
                continue
            if out and out[-1][1:] == (l, p):
                out[-1] = (out[-1][0] + m, l, p)
            else:
                out.append((m, l, p))
        return Irreps(out)

    def sort_perm(self) -> List[int]:
        return sorted(range(len(self)), key=lambda i: irrep_sort_key(self[i][1], self[i][2]))

    def regroup(self) -> "Irreps":
        return Irreps([self[i] for i in self.sort_perm()]).simplify()

    def filter(self, keep=None, drop=None) -> "Irreps":
        if keep is not None:
            ks = {(l, p) for _, l, p in Irreps(keep)}
            return Irreps([t for t in self if (t[1], t[2]) in ks])
        ds = {(l, p) for _, l, p in Irreps(drop)}
        return Irreps([t for t in self if (t[1], t[2]) not in ds])

    def regroup_index_map(self) -> List[int]:
        """For each component of ``self.regroup()``, its offset in ``self``'s layout."""
        offs = self.offsets()
        out: List[int] = []
        for i in self.sort_perm():
            m, l, _ = self[i]
            out.extend(range(offs[i], offs[i] + m * (2 * l + 1)))
        return out


def balanced_irreps(lmax: int, feature_size: int, use_sh: bool = True) -> Irreps:
    """Hidden-irreps allocation of the reference (models/utils/irreps_utils.py:6-25).

    The running ``total_dim`` is *assigned*, not accumulated, upstream (:18,:21);
    the resulting hidden sizes (128 / 170 / 188 for lmax 1 / 2 / 3 at
    feature_size 128) are part of the model definition and are kept."""
    n_types = 1 + (lmax if use_sh else 2 * lmax)
    chunks: List[MulIrT] = []
    last_dim = 0
    for level in range(1, lmax + 1):
        width = 2 * level + 1
        mul = int(feature_size / width / n_types)
        if mul == 0:
            break
        if use_sh:
            chunks.append((mul, level, 1 if level % 2 == 0 else -1))
            last_dim = mul * width
        else:
            chunks += [(mul, level, 1), (mul, level, -1)]
            last_dim = 2 * mul * width
    return Irreps([(feature_size - last_dim, 0, 1)] + chunks)
