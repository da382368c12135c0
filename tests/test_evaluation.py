"""Evaluation path: calibration scores against values produced by the REFERENCE's own functions
(tests/golden/make_golden_metrics.py), EMA, and -- on the GPU -- predict / accuracy / evaluate against the oracle's
per-sample predictions with the same Brownian increments and head noise.

Tolerances: scores 1e-12 (fp64 numpy on both sides); S-sample mean log-probabilities and weight trajectories vs the fp64
oracle rtol 1e-4 (fp32 conv path)."""
import numpy as np
import pytest
import torch

from bayesde_b200 import evaluation as ev
from oracle import jax_path as jp

GOLD = np.load(__file__.rsplit("/", 1)[0] + "/golden/metrics.npz")
DEV = "cuda:0"


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_scores_match_reference_functions(tag):
    p, oh = GOLD[f"{tag}_p"], GOLD[f"{tag}_onehot"]
    cal = ev.get_calibration(oh, p, num_bins=10)
    assert abs(cal["ece"] - GOLD[f"{tag}_cal_ece"]) < 1e-12
    np.testing.assert_allclose(cal["reliability_diag"][0], GOLD[f"{tag}_cal_conf"], rtol=0, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(cal["reliability_diag"][1], GOLD[f"{tag}_cal_acc"], rtol=0, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(cal["nb_items"], GOLD[f"{tag}_cal_items"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(np.asarray(ev.score_model(p, oh, bins=15)), GOLD[f"{tag}_score"], rtol=1e-12, atol=1e-12)
    conf, pred, true = p.max(1), p.argmax(1), oh.argmax(1)
    assert abs(ev.ECE(conf, pred, true, 0.1) - GOLD[f"{tag}_ece10"]) < 1e-12
    assert abs(ev.MCE(conf, pred, true, 0.1) - GOLD[f"{tag}_mce10"]) < 1e-12
    assert abs(ev.brier_score(oh, p) - GOLD[f"{tag}_brier"]) < 1e-15


def test_scores_edge_cases():
    p = np.array([[1.0, 0.0], [0.0, 1.0]])
    oh = np.array([[1.0, 0.0], [1.0, 0.0]])
    with np.errstate(invalid="ignore"):
        assert ev.get_calibration(oh, p)["ece"] == 0.0       # reference semantics: conf == 1.0 is outside the half-open bins [lo, hi)
    acc, err, ece, mce, brier, loss = ev.score_model(p, oh, bins=10)
    assert acc == 50.0 and err == 50.0 and ece == pytest.approx(0.5) and np.isfinite(loss)
    assert ev.brier_score(np.array([1.0, 0.0]), np.array([0.5, 0.5])) == pytest.approx(0.25)


def test_update_ema():
    e = [torch.ones(3), torch.zeros(2)]
    p = [torch.zeros(3), torch.ones(2)]
    ev.update_ema(e, p, momentum=0.9)
    assert torch.allclose(e[0], torch.full((3,), 0.9)) and torch.allclose(e[1], torch.full((2,), 0.1))


@pytest.mark.gpu
def test_gpu_accuracy_and_evaluate_match_oracle(built_lib):
    from bayesde_b200.classification import SDEBNNClassifier
    S, B = 3, 4
    kw = dict(input_size=(8, 8, 3), aug=1, nblocks=(1, 1), fx_dim=8, fw_dims=(2, 8, 2), diff_coef=0.1, nsteps=5)
    net_r = jp.SDEBNNClassifier(batch=B, **kw)
    params_r, head_r = net_r.init(0, torch.float64, fw_last_std=0.05)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 8, 8, 3, generator=g, dtype=torch.float64)
    labels = torch.tensor([1, 7, 3, 3])
    onehot = torch.nn.functional.one_hot(labels, 10).double()
    dWs = [torch.stack([jp.sample_increments(b.P, 5, g, torch.float64) for _ in range(S)]) for b in net_r.blocks()]
    eps = torch.randn(S, net_r.feat * 10 + 10, generator=g, dtype=torch.float64)
    logps, wlast = [], []
    for s in range(S):
        logp, _ = net_r.predict(params_r, head_r, x, [d[s] for d in dWs], eps[s])
        logps.append(logp)
        # trajectory of the LAST block (the info dict keeps one key, brax.py:211)
        h = jp.augment(x, net_r.aug)
        bi = 0
        for st in net_r.stages:
            if st[0] == "sde":
                h, _, wpath = st[1].apply(params_r[bi], h, dWs[bi][s])
                bi += 1
            else:
                h = jp.squeeze_downsample(h)
        wlast.append(wpath)
    avg_r = torch.stack(logps).mean(0)
    wts_r = torch.stack(wlast).mean(0)

    net = SDEBNNClassifier(nsamples=S, device=DEV, **kw)
    net.load_oracle_params(params_r, head_r)
    rng = net.explicit_rng([d.float().to(DEV) for d in dWs], eps.float().to(DEV))
    nc, nt, avg, wts = ev.accuracy(net, (x.float().to(DEV), onehot.float().to(DEV)), S, rng)
    assert nt == B and nc == int((avg_r.argmax(1) == labels).sum())
    assert torch.allclose(avg.double().cpu(), avg_r, rtol=1e-4, atol=1e-5)
    assert wts.shape == wts_r.shape and torch.allclose(wts.double().cpu(), wts_r, rtol=1e-4, atol=1e-6)
    # evaluate(): two batches from a list "loader", integer seeds (in-kernel Philox): shapes, determinism, score plumbing
    loader = [(x.float().to(DEV), labels.to(DEV)), (x.float().to(DEV).flip(0), labels.flip(0).to(DEV))]
    acc1, logits, lab, nll, kl, w = ev.evaluate(net, loader, S, rng=5, kl_coef=1e-3)
    acc2, logits2, *_ = ev.evaluate(net, loader, S, rng=5, kl_coef=1e-3)
    assert logits.shape == (2 * B, 10) and lab.shape == (2 * B, 10) and w.shape[0] == 2 * wts.shape[0]
    assert acc1 == acc2 and np.array_equal(logits, logits2) and np.isfinite(nll) and kl > 0
    scores = ev.score_model(np.exp(logits), lab)
    assert scores[0] == pytest.approx(100 * acc1)
