"""CPU oracle for the bayeSDE fixed-grid SDE-BNN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bayesde_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  The product path fails loudly when the CUDA
extension is missing instead of falling back to this code.

PARITY STATUS: **parity unpinned** at the ``sdeint_ito_fixed_grid`` / ``SDEBNN``
boundary — the reference holds no golden vectors or known-answer tests for the
fixed-grid path and JAX 0.1.77 / torchsde are not
installable here, so the JAX front-end restatement (``oracle/jax_path.py``) is
pinned only through (i) an independent explicit-loop restatement of
``lax.conv_general_dilated`` (``oracle/xla_conv_loops.py``), (ii) fp64
finite-difference checks of its own autograd, and (iii) the reference code that
*does* run here: the torchbnn conv stack / w-net (``make_y_net``/``make_w_net``)
and ``examples/torch/sdebnn_toy1d.py``'s Euler-Maruyama + KL loop, whose outputs
are committed as fixtures under ``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
