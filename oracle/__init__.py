"""CPU oracle for the eqnn-jax equivariant message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``eqnn_jax_b200``)
may import, call, link or execute anything in this directory.  The only
permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

PARITY UNPINNED.  The reference (smsharma/eqnn-jax, /root/reference) is pure
Python on top of jax / flax / jraph / e3nn-jax.  None of those packages is
installed in this image (nor on the GPU boxes) and none is vendored in the
reference tree (``requirements.txt:1-5`` lists them un-pinned), so the
reference cannot be executed here and its own tests hold no golden vectors for
this path (only equivariance at rtol 0.05, ``tests/test_equivariance.py:112``).
This oracle is therefore a *restatement* of
  - ``models/nequip.py``, ``models/utils/graph_utils.py:13-72,248-324``,
    ``models/utils/irreps_utils.py``, ``models/mlp.py`` (reference code), and
  - the published algorithms of e3nn-jax (~0.20), jraph 0.0.6.dev0 and flax
    that those files call,
anchored on what *can* be pinned without running JAX:
  - closed-form values (C(1,1,0)=delta/sqrt3, C(1,1,1)=eps/sqrt6, the l<=2 real
    spherical harmonics written out in SURVEY.md A.3),
  - an independent Clebsch-Gordan computation through sympy,
  - rotation equivariance of every operator (Wigner-D built from the SH),
  - the parameter counts printed by ``notebooks/examples.ipynb``,
  - the reference's own property tests (equivariance, PBC sanity) re-run on it.
"""
