"""GPU parity: get_equivariant_graph (models/utils/equivariant_graph_utils.py:27-102) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import graph_ref as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lmax,norm,pbc,vel,n_rb", [
    (1, "integral", True, False, 0),       # train_velocities.py:454-458 (lmax_attributes=1, PBC, |r| messages)
    (2, "integral", True, True, 16),       # train_cosmology.py:185 + steerable velocities + Bessel messages
    (3, "component", False, True, 8),
    (2, "norm", False, False, 0),
])
def test_steerable_attributes_match_oracle(lmax, norm, pbc, vel, n_rb):
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
    rng = np.random.default_rng(lmax)
    B, N, k = 3, 200, 7
    x = rng.uniform(0.0, 3.4, size=(B, N, 3)).astype(np.float32)
    v = rng.normal(size=(B, N, 3)).astype(np.float32)
    v[0, 0] = 0.0                                                     # zero vector: e3nn's |r| = 0 guard
    cell = np.eye(3, dtype=np.float32) * 3.4
    opbc = G.get_apply_pbc(cell=cell) if pbc else None
    gpbc = GU.get_apply_pbc(cell=cell) if pbc else None
    g = G.build_graph(x, None, k, apply_pbc=opbc)                     # kNN with self loops: r_ii = 0 edges included
    want = G.get_equivariant_graph(None, x, v, g.senders, g.receivers, g.n_node, g.n_edge, None, None, lmax,
                                   steerable_velocities=vel, spherical_harmonics_norm=norm, apply_pbc=opbc,
                                   n_radial_basis=n_rb, r_max=1.2)
    got = get_equivariant_graph(None, torch.tensor(x).cuda(), torch.tensor(v).cuda(), torch.tensor(g.senders).cuda(),
                                torch.tensor(g.receivers).cuda(), g.n_node, g.n_edge, None, None, lmax,
                                steerable_velocities=vel, spherical_harmonics_norm=norm, apply_pbc=gpbc,
                                n_radial_basis=n_rb, r_max=1.2)
    assert str(got.steerable_edge_attrs.irreps) == "+".join(f"1x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(lmax + 1))
    for name in ("steerable_edge_attrs", "steerable_node_attrs", "additional_messages"):
        a = getattr(got, name).array.cpu().numpy().astype(np.float64)
        b = np.asarray(getattr(want, name), dtype=np.float64)
        assert a.shape == b.shape, (name, a.shape, b.shape)
        err = np.abs(a - b).max()
        assert err <= 1e-5 * max(np.abs(b).max(), 1e-30), f"{name}: max err {err:.3e} (scale {np.abs(b).max():.3e})"


def test_cpu_tensor_fails_loudly():
    from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
    with pytest.raises(RuntimeError):
        get_equivariant_graph(None, torch.zeros(1, 4, 3), None, torch.zeros(1, 4, dtype=torch.int32),
                              torch.zeros(1, 4, dtype=torch.int32), None, None, None, None, 1)
