"""CPU checks of oracle/pairs_ref.py (the NumPy restatement of the pairing rule of csrc/pairs.cu) on graphs whose answer is
known by construction, and of the claim the sharing rests on: the two directions of an edge have bit-identical fp32 lengths
(models/nequip.py:134 forms r_ij = x_s - x_r; the radial MLP of :55-68 sees only |r_ij|)."""
import numpy as np

from oracle import graph_ref as G
from oracle import pairs_ref as PR


def _csr(senders, receivers, n):
    order = np.lexsort((np.arange(len(receivers)), receivers))           # receiver-sorted, ascending edge id inside a row
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(receivers, minlength=n))])
    return rowptr, senders[order], order


def test_reverse_edges_have_bit_identical_lengths():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4000, 3)).astype(np.float32) * np.float32(37.5)
    i, j = rng.integers(0, 4000, size=20000), rng.integers(0, 4000, size=20000)
    def length(a, b):
        r = (x[a] - x[b]).astype(np.float32)
        return np.sqrt(((r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]).astype(np.float32) + r[:, 2] * r[:, 2]).astype(np.float32))
    assert np.array_equal(length(i, j).view(np.uint32), length(j, i).view(np.uint32))
    assert np.all(length(i, i).view(np.uint32) == 0)                      # self-loops: +0.0


def test_hand_built_graph():
    # edges (sender -> receiver): 0<->1 reciprocated, 2->0 one way, self-loops on 0 and 2, a duplicate of 1->0
    senders = np.array([0, 1, 2, 0, 2, 1])
    receivers = np.array([1, 0, 0, 0, 2, 0])
    x = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 2.0], [0.0, 3.0, 4.0]], dtype=np.float32)
    rowptr, src, order = _csr(senders, receivers, 3)
    d = np.linalg.norm(x[src] - x[np.repeat(np.arange(3), np.diff(rowptr))], axis=1).astype(np.float32)
    uidx, rep_slot, rep_partner, zero_slot, zero_row = PR.edge_pairs(rowptr, src, d)
    # CSR slots: row 0 <- [1, 2, 0, 1] (edge ids 1, 2, 3, 5), row 1 <- [0] (edge 0), row 2 <- [2] (edge 4)
    assert list(src) == [1, 2, 0, 1, 0, 2]
    # slot 0 (1->0) pairs with slot 4 (0->1); slot 3 (the duplicate 1->0) stays alone; slots 2 and 5 are the zero class
    assert uidx[0] == uidx[4] and uidx[3] != uidx[0]
    assert uidx[2] == uidx[5] == zero_row and list(zero_slot) == [2, -1, 5]
    assert len(rep_slot) == 4 and sorted(rep_partner.tolist()) == [-2, -1, -1, 4]
    assert np.array_equal(d[rep_slot][uidx].view(np.uint32), d.view(np.uint32))


def test_knn_graph_sharing_ratio():
    rng = np.random.default_rng(3)
    N, k = 400, 12
    x = rng.uniform(size=(N, 3)).astype(np.float32)
    s, r, _ = G.nearest_neighbors(x, k, None)
    s, r = np.asarray(s).reshape(-1), np.asarray(r).reshape(-1)
    rowptr, src, order = _csr(s, r, N)
    recv = np.repeat(np.arange(N), np.diff(rowptr))
    rr = (x[src] - x[recv]).astype(np.float32)
    d = np.sqrt(((rr[:, 0] * rr[:, 0] + rr[:, 1] * rr[:, 1]).astype(np.float32) + rr[:, 2] * rr[:, 2]).astype(np.float32))
    uidx, rep_slot, rep_partner, zero_slot, zero_row = PR.edge_pairs(rowptr, src, d)
    members = np.bincount(uidx)
    assert members[zero_row] == N and np.all(np.delete(members, zero_row) <= 2)
    assert len(rep_slot) < 0.65 * N * k                                   # most kNN edges are reciprocated
    assert np.array_equal(d[rep_slot][uidx].view(np.uint32), d.view(np.uint32))
