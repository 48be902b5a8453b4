"""GPU parity: kNN-with-PBC graph construction, CSR, edge geometry (bit-exact / integer work)."""
import numpy as np
import pytest
import torch

from oracle import graph_ref as G

pytestmark = pytest.mark.gpu


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def _knn_both(x, k, pbc_np=None, pbc_gpu=None):
    from eqnn_jax_b200 import ops
    s, t, dr = ops.knn_pbc(_dev(x), k, None if pbc_gpu is None else pbc_gpu.cell,
                           None if pbc_gpu is None else pbc_gpu.cell_inv)
    ref = [G.nearest_neighbors(x[b, :, :3], k, pbc_np) for b in range(x.shape[0])]
    return (s.cpu().numpy(), t.cpu().numpy(), dr.cpu().numpy()), ref


@pytest.mark.parametrize("N,k", [(5, 5), (10, 10), (64, 1), (300, 20), (333, 32), (257, 33), (200, 64), (2500, 20)])
@pytest.mark.parametrize("pbc", [False, True])
def test_knn_bit_exact_random(N, k, pbc):
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(N * 100 + k)
    x = rng.uniform(size=(2, N, 7)).astype(np.float32)
    std = np.array([0.28871, 0.28875, 0.28870], dtype=np.float32)
    xs = ((x - 0.5) / 0.2887).astype(np.float32)
    (s, t, dr), ref = _knn_both(xs, k, G.get_apply_pbc(std) if pbc else None, GU.get_apply_pbc(std) if pbc else None)
    for b in range(2):
        np.testing.assert_array_equal(s[b], ref[b][0])
        np.testing.assert_array_equal(t[b], ref[b][1])
        assert np.array_equal(dr[b].view(np.uint32), ref[b][2].view(np.uint32)), "dr not bit-identical"
    assert np.all(t.reshape(2, N, k)[:, :, 0] == np.arange(N)[None]), "self must be neighbour 0"


def test_knn_ties_lattice_and_duplicates():
    from eqnn_jax_b200 import graph_utils as GU
    # integer lattice: many exactly equal distances; duplicates: zero distances
    g = np.stack(np.meshgrid(*[np.arange(6)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    x = np.concatenate([g, g[:40]], 0)[None]
    for k in (7, 27, 40):
        (s, t, dr), ref = _knn_both(x, k)
        np.testing.assert_array_equal(t[0], ref[0][1])
    cell = np.eye(3, dtype=np.float32) * 6
    (s, t, dr), ref = _knn_both(x, 19, G.get_apply_pbc(cell=cell), GU.get_apply_pbc(cell=cell))
    np.testing.assert_array_equal(t[0], ref[0][1])
    assert np.array_equal(dr[0].view(np.uint32), ref[0][2].view(np.uint32))


def test_knn_triclinic_cell():
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(7)
    cell = np.array([[1.0, 0.0, 0.0], [0.3, 0.9, 0.0], [0.1, 0.2, 1.1]], dtype=np.float32)
    x = (rng.uniform(size=(1, 150, 3)) @ cell).astype(np.float32)
    (s, t, dr), ref = _knn_both(x, 12, G.get_apply_pbc(cell=cell), GU.get_apply_pbc(cell=cell))
    np.testing.assert_array_equal(t[0], ref[0][1])


def test_knn_full_size_quijote_shape_and_properties():
    """cfg-2 size (N=5000, k=20, PBC 1000 Mpc/h box standardised): bit-exact against the dense oracle."""
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(0)
    pos = rng.uniform(0, 1000, size=(1, 5000, 3))
    std = np.array([288.71092533309235, 288.7525818573022, 288.70234893905575])
    mean = np.array([499.91877075908684, 500.0947802559321, 499.964508664328])
    xs = ((pos - mean) / std).astype(np.float32)
    (s, t, dr), ref = _knn_both(xs, 20, G.get_apply_pbc(std / 1000.0), GU.get_apply_pbc(std / 1000.0))
    np.testing.assert_array_equal(t[0], ref[0][1])
    assert np.array_equal(dr[0].view(np.uint32), ref[0][2].view(np.uint32))
    d = np.sqrt((dr[0].astype(np.float64) ** 2).sum(-1)).reshape(5000, 20)
    assert np.all(np.diff(d, axis=1) >= 0), "neighbours must be sorted by distance"
    assert d.max() < 0.5 * 1000.0 / std.max() * np.sqrt(3)


def test_build_graph_matches_reference_signature():
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(1)
    x = rng.uniform(size=(3, 50, 3)).astype(np.float32)
    for n_rb in (0, 8):
        g = GU.build_graph(_dev(x), None, k=10, apply_pbc=GU.get_apply_pbc(), n_radial_basis=n_rb, r_max=0.6)
        r = G.build_graph(x, None, k=10, apply_pbc=G.get_apply_pbc(), n_radial_basis=n_rb, r_max=0.6)
        np.testing.assert_array_equal(g.senders.cpu().numpy(), r.senders)
        np.testing.assert_array_equal(g.receivers.cpu().numpy(), r.receivers)
        assert g.n_edge.tolist() == r.n_edge.tolist() and g.n_node.tolist() == r.n_node.tolist()
        assert tuple(g.edges.shape) == tuple(r.edges.shape)
        np.testing.assert_allclose(g.edges.cpu().numpy(), r.edges, rtol=0, atol=2e-6 * np.abs(r.edges).max())
    # reference PBC sanity tests (tests/test_pbcs.py:27-33)
    x = rng.uniform(size=(1, 10, 3)).astype(np.float32)
    g0 = GU.build_graph(_dev(x), None, k=10)
    g1 = GU.build_graph(_dev(x), None, k=10, apply_pbc=GU.get_apply_pbc())
    assert g0.edges.mean() > g1.edges.mean() and g1.edges.max() < np.sqrt(3) / 2


def test_nearest_neighbors_single_graph_signature():
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(2)
    x = rng.normal(size=(40, 3)).astype(np.float32)
    s, t, dr = GU.nearest_neighbors(_dev(x), 6)
    rs, rt, rdr = G.nearest_neighbors(x, 6)
    np.testing.assert_array_equal(s.cpu().numpy(), rs)
    np.testing.assert_array_equal(t.cpu().numpy(), rt)
    assert np.array_equal(dr.cpu().numpy().view(np.uint32), rdr.view(np.uint32))


def test_knn_errors():
    from eqnn_jax_b200 import ops
    with pytest.raises(RuntimeError):
        ops.knn_pbc(torch.zeros(1, 4, 3).cuda(), 5)          # k > N
    with pytest.raises(RuntimeError):
        ops.knn_pbc(torch.zeros(1, 4, 3), 2)                 # CPU tensor: no fallback


@pytest.mark.parametrize("B,N,E", [(1, 5, 20), (3, 50, 500), (2, 1000, 20000)])
def test_csr_build_is_stable_counting_sort(B, N, E):
    from eqnn_jax_b200 import ops
    rng = np.random.default_rng(E)
    snd = rng.integers(0, N, size=(B, E)).astype(np.int32)
    rcv = rng.integers(0, N, size=(B, E)).astype(np.int32)
    if E >= 500:
        rcv[:, : E // 4] = 3                                  # a hub row
    rowptr, perm, src = ops.csr_build(_dev(snd, torch.int32), _dev(rcv, torch.int32), N)
    grcv = (rcv + (np.arange(B) * N)[:, None]).reshape(-1)
    gsnd = (snd + (np.arange(B) * N)[:, None]).reshape(-1)
    want_perm = np.argsort(grcv, kind="stable")
    want_rowptr = np.concatenate([[0], np.cumsum(np.bincount(grcv, minlength=B * N))])
    np.testing.assert_array_equal(rowptr.cpu().numpy(), want_rowptr)
    np.testing.assert_array_equal(perm.cpu().numpy(), want_perm)
    np.testing.assert_array_equal(src.cpu().numpy(), gsnd[want_perm])


def test_edge_geometry_matches_reference_rounding():
    from eqnn_jax_b200 import ops
    rng = np.random.default_rng(5)
    B, N, k = 2, 60, 9
    x = rng.normal(size=(B, N, 7)).astype(np.float32)
    g = G.build_graph(x, None, k)
    rowptr, perm, src = ops.csr_build(_dev(g.senders, torch.int32), _dev(g.receivers, torch.int32), N)
    geom, radial = ops.edge_geom(_dev(x), 7, B * N, rowptr, src, 16, 0.6)
    p = perm.cpu().numpy()
    grcv = (g.receivers + (np.arange(B) * N)[:, None]).reshape(-1)[p]
    gsnd = src.cpu().numpy()
    pos = x.reshape(B * N, 7)[:, :3]
    r = pos[gsnd] - pos[grcv]
    sq = r * r
    d = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2]).astype(np.float32)
    gm = geom.cpu().numpy()
    assert np.array_equal(gm[:, 3].view(np.uint32), d.view(np.uint32)), "edge length not bit-identical"
    u = np.where(d[:, None] == 0, 0, r / np.where(d == 0, 1, d)[:, None])
    np.testing.assert_allclose(gm[:, :3], u, atol=1e-7)
    want = G.bessel_np(d, 16, 0.6)
    np.testing.assert_allclose(radial.cpu().numpy(), want, rtol=0, atol=4e-6 * np.abs(want).max())


@pytest.mark.parametrize("N,k,kind", [(5000, 20, "uniform"), (6000, 32, "clustered"), (4096, 20, "duplicates"),
                                      (20000, 32, "uniform"), (3000, 40, "uniform")])
def test_cell_list_knn_is_bit_identical_to_dense_sweep(N, k, kind):
    """The cell-list search (large periodic clouds) must return exactly the dense sweep's indices and displacements:
    same fp32 distance expression, list ordered by (distance, index) = the reference's stable argsort.  Clustered
    points force the search radius to grow; duplicated points create exact distance ties."""
    from eqnn_jax_b200 import ops
    rng = np.random.default_rng(N + k)
    L = 3.4
    if kind == "clustered":
        centres = rng.uniform(-L / 2, L / 2, size=(12, 3))
        x = centres[rng.integers(0, 12, size=N)] + 0.03 * rng.normal(size=(N, 3))
        x[: N // 10] = rng.uniform(-L / 2, L / 2, size=(N // 10, 3))       # sparse background: far k-th neighbours
    else:
        x = rng.uniform(-L / 2, L / 2, size=(N, 3))
    if kind == "duplicates":
        x[N // 2:] = x[: N - N // 2]                                        # every point twice: exact ties
    x = np.stack([x, np.roll(x, 1, axis=0)]).astype(np.float32)             # B = 2
    cell = np.eye(3, dtype=np.float32) * L
    cinv = np.linalg.inv(cell).astype(np.float32)
    xt = torch.tensor(x).cuda()
    assert ops.lib().eqnn_knn_workspace_bytes(2, N, k) > 0
    s0, r0, d0 = ops.knn_pbc(xt, k, cell, cinv, dense=True)
    s1, r1, d1 = ops.knn_pbc(xt, k, cell, cinv)
    assert torch.equal(s0, s1)
    assert torch.equal(r0, r1)
    assert torch.equal(d0.view(torch.int32), d1.view(torch.int32))


@pytest.mark.parametrize("pbc", [None, "diag", "triclinic"])
def test_radius_graph_matches_the_oracle_bit_for_bit(pbc):
    """The radius branch of build_graph (graph_utils.py:266-277; cannot run upstream): pairs with 0 < d < radius in
    np.where order, compute_distances' arithmetic.  Indices, counts and distances must be identical to the oracle."""
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(3)
    B, N, radius = 3, 300, 0.21
    x = rng.uniform(0.0, 1.0, size=(B, N, 3)).astype(np.float32)
    x[1, 7] = x[1, 3]                                   # a duplicate: distance exactly 0 is excluded
    cell = None
    if pbc == "diag":
        cell = np.diag([1.0, 0.9, 1.1]).astype(np.float32)
    elif pbc == "triclinic":
        cell = np.array([[1.0, 0.0, 0.0], [0.2, 0.9, 0.0], [0.1, 0.3, 1.1]], dtype=np.float32)
    o_pbc = G.get_apply_pbc(cell=cell) if cell is not None else None
    g_pbc = GU.get_apply_pbc(cell=cell) if cell is not None else None
    ws, wt, wd, wn = G.radius_graph(x, radius, o_pbc)
    g = GU.build_graph(torch.tensor(x).cuda(), None, k=0, radius=radius, apply_pbc=g_pbc)
    assert np.array_equal(g.n_edge.cpu().numpy(), wn)
    assert np.array_equal(g.senders.cpu().numpy(), ws) and np.array_equal(g.receivers.cpu().numpy(), wt)
    assert g.edges.shape == (1, len(wd), 1)
    assert np.array_equal(g.edges.cpu().numpy()[0, :, 0], wd)
    assert wn.sum() > 0 and not np.any(ws == wt)
    # Bessel edge features of the same graph (graph_utils.py:300-302)
    gb = GU.build_graph(torch.tensor(x).cuda(), None, k=0, radius=radius, apply_pbc=g_pbc, n_radial_basis=8, r_max=0.3)
    want = G.bessel_np(wd[:, None], 8, 0.3).reshape(len(wd), 8)
    np.testing.assert_allclose(gb.edges.cpu().numpy()[0], want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


def test_radius_graph_empty_and_full():
    from eqnn_jax_b200 import graph_utils as GU
    x = torch.tensor(np.random.default_rng(0).uniform(0, 1, size=(2, 40, 3)).astype(np.float32)).cuda()
    g = GU.build_graph(x, None, k=0, radius=1e-6)
    assert int(g.n_edge.sum()) == 0 and g.senders.numel() == 0
    g = GU.build_graph(x, None, k=0, radius=10.0)
    assert g.n_edge.tolist() == [40 * 39, 40 * 39]
