"""Oracle (test infrastructure only): e3nn-jax / jraph / flax operators in torch.

CPU restatement of the third-party operators the reference's hot path calls
(SURVEY.md Appendix A.4-A.10); evaluated in float64 for "truth" and float32 for
"reference-like" rounding.  torch autograd supplies the reference gradients.

Reference call sites: models/nequip.py:43 (Linear), :46-52 (SH, tensor_product),
:57 (bessel), :62-66 (MultiLayerPerceptron), :68 (IrrepsArray * array),
:85-90 (Linear, gate), :118-120 (jraph segment_*), :187 (scatter_sum).
"""
from __future__ import annotations

import functools
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .so3 import Irreps, clebsch_gordan, spherical_harmonics  # noqa: F401


# ----------------------------------------------------------------------------
# IrrepsArray
# ----------------------------------------------------------------------------
class IrrepsArray:
    """(irreps, array) pair; last axis = concatenated chunks, each (mul, 2l+1) row-major."""

    def __init__(self, irreps, array: torch.Tensor):
        self.irreps = Irreps(irreps)
        self.array = array
        assert array.shape[-1] == self.irreps.dim, (self.irreps, array.shape)

    @property
    def chunks(self) -> List[torch.Tensor]:
        out = []
        for (mul, l, _), sl in zip(self.irreps, self.irreps.slices()):
            out.append(self.array[..., sl].reshape(self.array.shape[:-1] + (mul, 2 * l + 1)))
        return out

    @staticmethod
    def from_chunks(irreps, chunks, lead_shape, dtype):
        irreps = Irreps(irreps)
        if len(chunks) == 0:
            return IrrepsArray(irreps, torch.zeros(lead_shape + (0,), dtype=dtype))
        arr = torch.cat([c.reshape(lead_shape + (-1,)) for c in chunks], dim=-1)
        return IrrepsArray(irreps, arr)

    def regroup(self) -> "IrrepsArray":
        """Stable sort of chunks in e3nn order, then merge equal neighbours (A.1)."""
        perm = self.irreps.sort_perm()
        sls = self.irreps.slices()
        arr = torch.cat([self.array[..., sls[i]] for i in perm], dim=-1) if perm else self.array
        return IrrepsArray(Irreps([self.irreps.items[i] for i in perm]).simplify(), arr)

    def filter(self, keep) -> "IrrepsArray":
        ks = {(l, p) for _, l, p in Irreps(keep)}
        sls = self.irreps.slices()
        idx = [i for i, (_, l, p) in enumerate(self.irreps) if (l, p) in ks]
        arr = torch.cat([self.array[..., sls[i]] for i in idx], dim=-1) if idx else self.array[..., :0]
        return IrrepsArray(Irreps([self.irreps.items[i] for i in idx]), arr)

    def first_irrep_copy(self) -> "IrrepsArray":
        """``slice_by_mul[:1]`` (models/nequip.py:130): first copy of the first chunk."""
        _, l, p = self.irreps.items[0]
        return IrrepsArray(Irreps([(1, l, p)]), self.array[..., : 2 * l + 1])

    def mul_by_scalars(self, w: torch.Tensor) -> "IrrepsArray":
        """``IrrepsArray * array[..., num_irreps]`` (nequip.py:68): one scalar per irrep copy."""
        assert w.shape[-1] == self.irreps.num_irreps
        reps = torch.tensor([2 * l + 1 for mul, l, _ in self.irreps for _ in range(mul)])
        return IrrepsArray(self.irreps, self.array * torch.repeat_interleave(w, reps, dim=-1))


# ----------------------------------------------------------------------------
# normalize_function (A.6)
# ----------------------------------------------------------------------------
def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


_ACTS = {"gelu": gelu_tanh, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
         "silu": torch.nn.functional.silu, "relu": torch.relu}


@functools.lru_cache(maxsize=None)
def normalization_constant(name: str) -> float:
    """c = sqrt(E_{x~N(0,1)} phi(x)^2) on e3nn's deterministic grid (A.6)."""
    from scipy.special import erfinv

    x = np.sqrt(2.0) * erfinv(np.linspace(-1.0, 1.0, 1_000_001 + 2)[1:-1])
    y = _ACTS[name](torch.from_numpy(x)).numpy()
    return float(np.sqrt(np.mean(y * y)))


def normalized_act(name: str):
    c = normalization_constant(name)
    f = _ACTS[name]
    if abs(c - 1.0) < 1e-8:
        return f
    return lambda x: f(x) / c


# ----------------------------------------------------------------------------
# e3nn.flax.Linear (A.4)
# ----------------------------------------------------------------------------
def linear_instructions(irreps_in: Irreps, irreps_out, force_irreps_out: bool = False):
    """Returns (regrouped input irreps, simplified output irreps, unsimplified
    output irreps, [(i_in, i_out, path_weight(g=1), name, shape)])."""
    irreps_in = Irreps(irreps_in).remove_zero_multiplicities().regroup()
    irreps_out = Irreps(irreps_out)
    if force_irreps_out:
        out_unsimplified = irreps_out
    else:
        out_unsimplified = irreps_out.filter(keep=irreps_in)
    out_simpl = out_unsimplified.simplify()
    ins = []
    for i_out, (mo, lo, po) in enumerate(out_simpl):
        fan_in = sum(mi for (mi, li, pi) in irreps_in if (li, pi) == (lo, po))
        for i_in, (mi, li, pi) in enumerate(irreps_in):
            if (li, pi) == (lo, po):
                name = f"w[{i_in},{i_out}] {Irreps([(mi, li, pi)])},{Irreps([(mo, lo, po)])}"
                ins.append((i_in, i_out, 1.0 / math.sqrt(fan_in), name, (mi, mo)))
    return irreps_in, out_simpl, out_unsimplified, ins


def linear_init(rng: np.random.Generator, irreps_in, irreps_out, force_irreps_out=False,
                gradient_normalization: float = 1.0, biases: bool = False) -> Dict[str, np.ndarray]:
    _, out_simpl, _, ins = linear_instructions(irreps_in, irreps_out, force_irreps_out)
    params = {}
    for _, _, alpha_sqrt, name, shape in ins:
        std = alpha_sqrt ** (1.0 - gradient_normalization)
        params[name] = rng.normal(size=shape) * std
    if biases:
        for i_out, (mo, lo, po) in enumerate(out_simpl):
            if (lo, po) == (0, 1):
                params[f"b[{i_out}] {Irreps([(mo, lo, po)])}"] = np.zeros((mo,))
    return params


def linear_apply(params: Dict[str, torch.Tensor], x: IrrepsArray, irreps_out, force_irreps_out=False,
                 gradient_normalization: float = 1.0) -> IrrepsArray:
    xin = x.regroup()
    irreps_in, out_simpl, out_unsimpl, ins = linear_instructions(x.irreps, irreps_out, force_irreps_out)
    assert irreps_in == xin.irreps
    xc = xin.chunks
    lead = x.array.shape[:-1]
    outs: List[Optional[torch.Tensor]] = [None] * len(out_simpl)
    for i_in, i_out, alpha_sqrt, name, _ in ins:
        w = params[name]
        pw = alpha_sqrt ** gradient_normalization
        y = pw * torch.einsum("uw,...ui->...wi", w, xc[i_in])
        outs[i_out] = y if outs[i_out] is None else outs[i_out] + y
    chunks = []
    for i_out, (mo, lo, po) in enumerate(out_simpl):
        c = outs[i_out]
        if c is None:
            c = torch.zeros(lead + (mo, 2 * lo + 1), dtype=x.array.dtype)
        bname = f"b[{i_out}] {Irreps([(mo, lo, po)])}"
        if bname in params:
            c = c + params[bname][..., :, None]
        chunks.append(c)
    out = IrrepsArray.from_chunks(out_simpl, chunks, lead, x.array.dtype)
    return IrrepsArray(out_unsimpl, out.array)  # rechunk: same memory layout


# ----------------------------------------------------------------------------
# e3nn.tensor_product (A.10)
# ----------------------------------------------------------------------------
def tensor_product_paths(irreps1: Irreps, irreps2: Irreps):
    """For regrouped inputs: ([(i1, i2, l_out, p_out)], unsorted output irreps)."""
    paths, out = [], []
    for i1, (m1, l1, p1) in enumerate(irreps1):
        for i2, (m2, l2, p2) in enumerate(irreps2):
            for lo in range(abs(l1 - l2), l1 + l2 + 1):
                paths.append((i1, i2, lo, p1 * p2))
                out.append((m1 * m2, lo, p1 * p2))
    return paths, Irreps(out)


def tensor_product(x: IrrepsArray, y: IrrepsArray) -> IrrepsArray:
    x, y = x.regroup(), y.regroup()
    paths, out_irreps = tensor_product_paths(x.irreps, y.irreps)
    xc, yc = x.chunks, y.chunks
    lead = x.array.shape[:-1]
    chunks = []
    for (i1, i2, lo, _) in paths:
        l1, l2 = x.irreps.items[i1][1], y.irreps.items[i2][1]
        cg = torch.as_tensor(clebsch_gordan(l1, l2, lo) * math.sqrt(2 * lo + 1), dtype=x.array.dtype)
        c = torch.einsum("...ui,...vj,ijk->...uvk", xc[i1], yc[i2], cg)
        chunks.append(c.reshape(lead + (-1, 2 * lo + 1)))
    return IrrepsArray.from_chunks(out_irreps, chunks, lead, x.array.dtype).regroup()


# ----------------------------------------------------------------------------
# e3nn.gate (A.8) with default activations
# ----------------------------------------------------------------------------
def gate(x: IrrepsArray, even_act="gelu", even_gate_act="sigmoid") -> IrrepsArray:
    scal = x.filter(keep="0e+0o")
    gated_irreps = Irreps([(m, l, p) for m, l, p in x.irreps if l > 0])
    gated = x.filter(keep=gated_irreps) if len(gated_irreps) else IrrepsArray("", x.array[..., :0])
    if any(p == -1 for _, l, p in scal.irreps):
        raise NotImplementedError("odd scalars are not on the reference's path")
    ns, ng = scal.irreps.num_irreps, gated.irreps.num_irreps
    assert ns >= ng
    extra = normalized_act(even_act)(scal.array[..., : ns - ng])
    if ng == 0:
        return IrrepsArray(Irreps([(ns, 0, 1)]), extra)
    gates = normalized_act(even_gate_act)(scal.array[..., ns - ng:])
    gated_out = gated.mul_by_scalars(gates)
    irreps = Irreps([(ns - ng, 0, 1)] + gated.irreps.items).remove_zero_multiplicities()
    return IrrepsArray(irreps, torch.cat([extra, gated_out.array], dim=-1))


# ----------------------------------------------------------------------------
# e3nn.bessel (A.7), e3nn.flax.MultiLayerPerceptron (A.5), models/mlp.py
# ----------------------------------------------------------------------------
def bessel(d: torch.Tensor, n: int, x_max: float, reference_rounding: bool = True) -> torch.Tensor:
    """e3nn.bessel: sqrt(2/c) * sin(j pi / c * x) / x  (limit j pi / c at x = 0).

    The phase j*pi/c*x reaches O(1e2..1e3) on this path, so its fp32 rounding moves the
    result by up to ~1e-4: it is the one place where the reference's fp32 arithmetic is
    not reproducible by "exact" math.  With ``reference_rounding`` the phase is formed
    with the reference's fp32 rounding points, ((j * pi) / c) * x with j, pi, c, x in
    fp32 (JAX weak-type promotion of the Python scalars), also when the oracle runs in
    fp64; everything downstream of the phase is evaluated in the working dtype."""
    x = d[..., None]
    if reference_rounding:
        x32 = x.to(torch.float32)
        j32 = torch.arange(1, n + 1, dtype=torch.float32)
        w32 = (j32 * math.pi) / x_max                       # fp32 mul, fp32 div (scalars cast to fp32)
        xs32 = torch.where(x32 == 0, torch.ones_like(x32), x32)
        phase = (w32 * xs32).to(d.dtype)
        w, xs = w32.to(d.dtype), xs32.to(d.dtype)
        val = torch.where(x == 0, w + 0 * x, torch.sin(phase) / xs)
        return float(np.float32(math.sqrt(2.0 / x_max))) * val   # jnp.sqrt(2.0 / x_max): weak-typed fp32 constant
    j = torch.arange(1, n + 1, dtype=d.dtype)
    xs = torch.where(x == 0, torch.ones_like(x), x)
    val = torch.where(x == 0, j * math.pi / x_max + 0 * x, torch.sin(j * math.pi / x_max * xs) / xs)
    return math.sqrt(2.0 / x_max) * val


def e3nn_mlp_init(rng, n_in: int, list_neurons) -> Dict[str, Dict[str, np.ndarray]]:
    params, fan = {}, n_in
    for i, h in enumerate(list_neurons):
        params[f"Dense_{i}"] = {"kernel": rng.normal(size=(fan, h))}  # gradient_normalization "path": std 1
        fan = h
    return params


def e3nn_mlp_apply(params, x: torch.Tensor, act: str = "gelu", output_activation: bool = False):
    n = len(params)
    f = normalized_act(act)
    for i in range(n):
        w = params[f"Dense_{i}"]["kernel"]
        x = (x @ w) / math.sqrt(w.shape[0])
        if i < n - 1 or output_activation:
            x = f(x)
    return x


def flax_mlp_init(rng, n_in: int, feature_sizes) -> Dict[str, Dict[str, np.ndarray]]:
    """models/mlp.py:7-28: Dense(xavier_uniform) + zero bias."""
    params, fan = {}, n_in
    for i, h in enumerate(feature_sizes):
        lim = math.sqrt(6.0 / (fan + h))
        params[f"Dense_{i}"] = {"kernel": rng.uniform(-lim, lim, size=(fan, h)), "bias": np.zeros((h,))}
        fan = h
    return params


def flax_mlp_apply(params, x: torch.Tensor, activate_final: bool = False):
    n = len(params)
    for i in range(n):
        p = params[f"Dense_{i}"]
        x = x @ p["kernel"] + p["bias"]
        if i < n - 1 or activate_final:
            x = gelu_tanh(x)
    return x


# ----------------------------------------------------------------------------
# jraph segment ops (A.9)
# ----------------------------------------------------------------------------
def segment_sum(data: torch.Tensor, ids: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros((n,) + data.shape[1:], dtype=data.dtype)
    return out.index_add(0, ids.long(), data)


def segment_mean(data, ids, n):
    num = segment_sum(data, ids, n)
    den = segment_sum(torch.ones_like(data), ids, n)
    return num / torch.clamp(den, min=1.0)


def segment_max(data, ids, n):
    out = torch.full((n,) + data.shape[1:], -float("inf"), dtype=data.dtype)
    idx = ids.long().reshape((-1,) + (1,) * (data.dim() - 1)).expand_as(data)
    return out.scatter_reduce(0, idx, data, reduce="amax", include_self=True)


SEGMENT = {"sum": segment_sum, "mean": segment_mean, "max": segment_max}
