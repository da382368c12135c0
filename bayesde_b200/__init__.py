"""bayesde_b200 -- B200-native (sm_100a) implementation of bayeSDE's fixed-grid SDE-BNN hot path.

Public surface (same names / call shapes as the reference):
    sdeint_ito_fixed_grid(f, g, y0, ts, rng, args=(), dt=1e-6, method='milstein', rep=0)
    sdeint_ito(f, g, y0, ts, rng, args=(), dt=1e-6, method='milstein', rep=0)   (stochastic adjoint, Philox Brownian tree)
    brax.SDEBNN(...), brax.MeanField, brax.bnn_serial, arch.MLP, arch.build_fx, arch.Augment, ...
    brax_utils.sdeint / _sdeint          (JAX toy solver, brax/utils/sdeint.py)
    brax_layers.SDELayer / mlp / ConcatSquashLineartx / ...  (JAX toy dynamics, brax/_impl/layers.py; one fused launch per solve)
    torch_toy.em_solver / VariationalSDE (torch toy, examples/torch/sdebnn_toy1d.py)
    evaluation.predict / accuracy / evaluate / update_ema / score_model (script :52-105, brax/utils/utils.py:122-307)
    torch_models.SDENet / make_y_net / make_w_net (torch front end, torchbnn/_impl/models.py)
All arithmetic on the hot path runs in libbsde.so (hand-written CUDA, C ABI in include/bsde.h);
there is no CPU / eager fallback.
"""
from . import arch, brax, brax_layers, brax_utils, evaluation, torch_models, torch_toy  # noqa: F401
from .sdeint import sdeint_ito_fixed_grid, sdeint_ito, make_brownian_motion, PRNGKey, split  # noqa: F401

__all__ = ["sdeint_ito_fixed_grid", "sdeint_ito", "make_brownian_motion", "PRNGKey", "split", "arch", "brax", "brax_utils", "brax_layers", "torch_toy", "torch_models", "evaluation"]
