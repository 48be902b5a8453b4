"""TEST INFRASTRUCTURE ONLY (never imported by the product): NumPy restatement of the pairing rule of
eqnn_jax_b200/csrc/pairs.cu -- which receiver-sorted CSR slots share one radial-MLP evaluation.

The reference evaluates the radial MLP per edge from the edge length alone (models/nequip.py:55-68: `e3nn.bessel(d, ...)`
-> MultiLayerPerceptron); sharing an evaluation between two edges is exact iff their fp32 lengths are bit-identical,
which this restatement checks directly."""
import numpy as np


def edge_pairs(rowptr, src, d):
    """rowptr [n+1], src [E] (global sender per slot), d [E] fp32 lengths -> (uidx [E], rep_slot [U], rep_partner [U])."""
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
    for p in range(E):
        r, s = recv[p], src[p]
        if s == r:
            continue
        q = first(s, r)
        if q >= 0 and first(r, s) == p and dbits[p] == dbits[q]:
            partner[p] = q
    is_rep = (partner < 0) | (np.arange(E) < partner)
    rep_slot = np.nonzero(is_rep)[0]
    uidx = np.empty(E, dtype=np.int64)
    uidx[rep_slot] = np.arange(len(rep_slot))
    follow = ~is_rep
    uidx[follow] = uidx[partner[follow]]
    return uidx, rep_slot, partner[rep_slot]
