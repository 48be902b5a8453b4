"""TEST INFRASTRUCTURE ONLY (never imported by the product): NumPy restatement of the pairing rule of
eqnn_jax_b200/csrc/pairs.cu -- which receiver-sorted CSR slots share one radial-MLP evaluation.

The reference evaluates the radial MLP per edge from the edge length alone (models/nequip.py:55-68: `e3nn.bessel(d, ...)`
-> MultiLayerPerceptron); sharing an evaluation between two edges is exact iff their fp32 lengths are bit-identical,
which this restatement checks directly."""
import numpy as np


def edge_pairs(rowptr, src, d):
    """rowptr [n+1], src [E] (global sender per slot), d [E] fp32 lengths ->
    (uidx [E], rep_slot [U], rep_partner [U] (-1 none, -2 zero class), zero_slot [n], zero_row)."""
    rowptr = np.asarray(rowptr).astype(np.int64)
    src = np.asarray(src).astype(np.int64)
    dbits = np.asarray(d, dtype=np.float32).view(np.uint32)
    n, E = len(rowptr) - 1, len(src)
    recv = np.repeat(np.arange(n), np.diff(rowptr))

    def first(row, who):
        for q in range(rowptr[row], rowptr[row + 1]):
            if src[q] == who:
                return q
        return -1

    partner = np.full(E, -1, dtype=np.int64)
    zero_slot = np.full(n, -1, dtype=np.int64)
    for p in range(E):
        r, s = recv[p], src[p]
        if s == r:
            # zero class: the row's first self-loop, if its length is +0.0
            if dbits[p] == 0 and first(r, r) == p:
                partner[p] = -2
                zero_slot[r] = p
            continue
        q = first(s, r)
        if q >= 0 and first(r, s) == p and dbits[p] == dbits[q]:
            partner[p] = q
    zs = zero_slot[zero_slot >= 0]
    zero_first = int(zs.min()) if len(zs) else -1
    idx = np.arange(E)
    is_rep = (partner == -1) | ((partner >= 0) & (idx < partner)) | ((partner == -2) & (idx == zero_first))
    rep_slot = np.nonzero(is_rep)[0]
    uidx = np.empty(E, dtype=np.int64)
    uidx[rep_slot] = np.arange(len(rep_slot))
    follow = (~is_rep) & (partner >= 0)
    uidx[follow] = uidx[partner[follow]]
    zfollow = (~is_rep) & (partner == -2)
    if zero_first >= 0:
        uidx[zfollow] = uidx[zero_first]
    zero_row = int(uidx[zero_first]) if zero_first >= 0 else -1
    return uidx, rep_slot, partner[rep_slot], zero_slot, zero_row
