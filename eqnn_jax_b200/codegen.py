"""Kernel generator: straight-line CUDA for the sparse Clebsch-Gordan contraction.

The tensor product at models/nequip.py:52 / :197 couples a handful of input
components with the (lmax+1)^2 harmonics through a sparse constant tensor.  A
table-driven loop would index registers dynamically (local memory); instead the
plan compiler emits one fully unrolled ``struct TP_<id>`` per (x irreps, lmax,
live set) signature, with the non-zero coefficients as literals, and the
hand-written kernels in csrc/tpconv.cu are templated on it.

Run at build time by ``__graft_entry__.build()``; output: csrc/gen/tp_gen.cuh.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Tuple

from . import so3
from .plan import TPPlan, make_tp_plan

# Signatures compiled into the library: (x irreps, lmax, live irreps or None)
_HIDDEN_LIVE = {1: "0e+1o", 2: "0e+1o+2e", 3: "0e+1o+2e+3o"}
PREBUILT: List[Tuple[str, int, Optional[str]]] = []
for _x in ("1o+1o+1x0e", "1o"):
    for _l in (1, 2, 3):
        PREBUILT.append((_x, _l, _HIDDEN_LIVE[_l]))   # interaction block (nequip.py:52), all hidden irreps
        PREBUILT.append((_x, _l, "0e+1o"))            # interaction block, irreps that survive nequip.py:162
        PREBUILT.append((_x, _l, "0e"))               # invariant read-out (nequip.py:197)
        PREBUILT.append((_x, _l, None))               # every message column: the node task's `edges` output (:154-171)


def _f(v: float) -> str:
    s = repr(float(v))
    if "e" not in s and "." not in s and "inf" not in s:
        s += ".0"
    return s + "f"


def _emit_sh(lmax: int) -> List[str]:
    out = [f"struct SH_{lmax} {{",
           f"  static constexpr int LMAX = {lmax}, NY = {(lmax + 1) ** 2};",
           "  // unit-norm real spherical harmonics of a unit (or zero) vector, e3nn component order",
           "  __device__ __forceinline__ static void sh(float ux, float uy, float uz, float* __restrict__ Y) {",
           "    const float u[3] = {ux, uy, uz};", "    Y[0] = 1.0f;"]
    if lmax >= 1:
        out.append("    Y[1] = ux; Y[2] = uy; Y[3] = uz;")
    for l in range(2, lmax + 1):
        terms: Dict[int, List[str]] = {}
        for i, j, k, c in so3.sh_recursion_terms(l):
            terms.setdefault(k, []).append(f"{_f(c)} * Y[{(l - 1) ** 2 + i}] * u[{j}]")
        for k in range(2 * l + 1):
            out.append(f"    Y[{l * l + k}] = " + " + ".join(terms.get(k, ["0.0f"])) + ";")
    out.append("  }")
    out.append("};")
    return out


def _group_columns(plan: TPPlan):
    """Live columns grouped by (l1,l2,l3) so the coef*Y products are shared across copies."""
    groups: Dict[Tuple[int, int, int], List] = {}
    for c in plan.columns:
        if c.live >= 0:
            groups.setdefault((c.l1, c.l2, c.l3), []).append(c)
    return groups


def _emit_fwd(plan: TPPlan) -> List[str]:
    out = ["  // m[live_off + k] = w[live] * sum_ij C_ijk x[i] Y[j]",
           "  __device__ __forceinline__ static void fwd(const float* __restrict__ x, const float* __restrict__ Y,",
           "                                             const float* __restrict__ w, float* __restrict__ m) {"]
    for (l1, l2, l3), cols in _group_columns(plan).items():
        terms = so3.cg_terms(l1, l2, l3)
        tag = f"{l1}{l2}{l3}"
        out.append("    {")
        for t, (i, j, k, c) in enumerate(terms):
            out.append(f"      const float c{tag}_{t} = {_f(c)} * Y[{l2 * l2 + j}];")
        for col in cols:
            for k in range(2 * l3 + 1):
                ts = [f"c{tag}_{t} * x[{col.x_off + i}]" for t, (i, j, kk, c) in enumerate(terms) if kk == k]
                out.append(f"      m[{col.live_off + k}] = w[{col.live}] * ({' + '.join(ts) if ts else '0.0f'});")
        out.append("    }")
    out.append("  }")
    return out


def _emit_bwd(plan: TPPlan) -> List[str]:
    out = ["  // dw[live] = sum_k g[live_off+k] T[k];  dx[i] += w[live] sum_jk C_ijk Y[j] g[live_off+k]",
           "  __device__ __forceinline__ static void bwd(const float* __restrict__ x, const float* __restrict__ Y,",
           "                                             const float* __restrict__ w, const float* __restrict__ g,",
           "                                             float* __restrict__ dw, float* __restrict__ dx) {",
           f"    #pragma unroll\n    for (int i = 0; i < {plan.dxp}; ++i) dx[i] = 0.0f;"]
    for (l1, l2, l3), cols in _group_columns(plan).items():
        terms = so3.cg_terms(l1, l2, l3)
        tag = f"{l1}{l2}{l3}"
        out.append("    {")
        for t, (i, j, k, c) in enumerate(terms):
            out.append(f"      const float c{tag}_{t} = {_f(c)} * Y[{l2 * l2 + j}];")
        by_i = {}
        for t, (i, j, k, c) in enumerate(terms):
            by_i.setdefault(i, []).append((t, k))
        for col in cols:
            # s_i = sum over the terms that touch x[i] of c_t * g[k]: one FMA per term, then dw += s_i x[i], dx[i] += w s_i
            out.append("      {")
            out.append("        float acc = 0.0f, s;")
            for i, tk in by_i.items():
                t0, k0 = tk[0]
                line = f"        s = c{tag}_{t0} * g[{col.live_off + k0}];"
                for t, k in tk[1:]:
                    line += f" s = fmaf(c{tag}_{t}, g[{col.live_off + k}], s);"
                out.append(line)
                out.append(f"        acc = fmaf(s, x[{col.x_off + i}], acc); "
                           f"dx[{col.x_off + i}] = fmaf(s, w[{col.live}], dx[{col.x_off + i}]);")
            out.append(f"        dw[{col.live}] = acc;")
            out.append("      }")
        out.append("    }")
    out.append("  }")
    return out


def emit_struct(idx: int, plan: TPPlan) -> List[str]:
    out = [f"// signature: {plan.signature}",
           f"struct TP_{idx} {{",
           f"  static constexpr int DX = {plan.dx}, DXP = {plan.dxp}, NY = {plan.ny}, LMAX = {plan.lmax};",
           f"  static constexpr int NLIVE = {plan.n_live}, DLIVE = {plan.d_live};"]
    out.append("  __device__ __forceinline__ static void sh(float ux, float uy, float uz, float* __restrict__ Y) {")
    out.append(f"    SH_{plan.lmax}::sh(ux, uy, uz, Y);")
    out.append("  }")
    out += _emit_fwd(plan)
    out += _emit_bwd(plan)
    out.append("};")
    return out


def generate(path: str, sigs: Iterable[Tuple[str, int, str]] = None) -> List[str]:
    sigs = list(PREBUILT if sigs is None else sigs)
    plans = [make_tp_plan(x, l, live) for x, l, live in sigs]
    seen, uniq = set(), []
    for p in plans:
        if p.signature not in seen:
            seen.add(p.signature)
            uniq.append(p)
    lines = ["// GENERATED by eqnn_jax_b200/codegen.py -- do not edit.", "#pragma once", ""]
    for l in (1, 2, 3, 4):
        lines += _emit_sh(l) + [""]
    for i, p in enumerate(uniq):
        lines += emit_struct(i, p) + [""]
    lines.append(f"#define EQNN_TP_COUNT {len(uniq)}")
    lines.append("static const char* const EQNN_TP_SIGNATURES[EQNN_TP_COUNT] = {")
    lines += [f'  "{p.signature}",' for p in uniq]
    lines.append("};")
    lines.append("#define EQNN_TP_DISPATCH(ID, ...) switch (ID) { \\")
    for i in range(len(uniq)):
        lines.append(f"  case {i}: {{ using TP = TP_{i}; __VA_ARGS__; break; }} \\")
    lines.append("  default: return eqnn_fail(EQNN_ERR_SIGNATURE, \"unknown tensor-product signature id\"); }")
    text = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return [p.signature for p in uniq]


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    print("\n".join(generate(os.path.join(here, "csrc", "gen", "tp_gen.cuh"))))
