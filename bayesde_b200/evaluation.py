"""Evaluation path of examples/jax/sdebnn_classification.py: S-sample predictive averaging with the
weight trajectory of the network (`full_output` / `sdebnn_w`), EMA of the parameters, and the calibration scores of
brax/utils/utils.py:122-307.  Forward only: the solves run through the same K1 / K2 kernels as training (one launch set
for all S samples); the scores are host-side numpy, as they are in the reference.

    predict(model, x, rng)                          script :52-54    -> (logp [S,B,C], kl [S], {"sdebnn_w": [S,T,P]})
    accuracy(model, (x, onehot), nsamples, rng)     script :56-72    -> (n_correct, n_total, avg_preds, avg_wts)
    update_ema(ema, params, momentum)               script :74-75
    evaluate(model, loader, nsamples, rng, kl_coef) script :78-105   -> (acc, logits, labels, nll, kl, wts)
    get_calibration / ECE / MCE / brier_score / score_model        brax/utils/utils.py:122-307
"""
import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# prediction with S weight samples
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def predict(model, inputs, rng, nsamples=None):
    """`_predict(params, inputs, rng=rng, full_output=True)` for `nsamples` rngs at once (the script vmaps `predict` over
    `jax.random.split(rng, nsamples)`, :60-61).  As in the reference every SDEBNN block reports under the same key, so the
    info dict holds the trajectory of the LAST block (`infodict.update`, brax.py:211)."""
    S = model.nsamples if nsamples is None else nsamples
    xs = inputs.unsqueeze(0).expand(S, *inputs.shape).contiguous()
    logp, kl, info = model._predict(model.params(), xs, rng=rng, mc_samples=S, full_output=True)
    return logp, kl, info


@torch.no_grad()
def accuracy(model, data, nsamples, rng):
    """script :56-72: mean of the S log-probability tensors -> argmax; mean weight trajectory."""
    inputs, targets = data
    target_class = targets.argmax(dim=1)
    preds, _, info = predict(model, inputs, rng, nsamples)
    avg_preds = preds.mean(0)
    predicted_class = avg_preds.argmax(dim=1)
    n_correct = int((predicted_class == target_class).sum())
    avg_wts = info["sdebnn_w"].mean(0)
    return n_correct, inputs.shape[0], avg_preds, avg_wts


@torch.no_grad()
def update_ema(ema_params, params, momentum=0.999):
    """script :74-75 on a list of tensors (e.g. `model.parameters()`), in place on `ema_params`."""
    for e, p in zip(ema_params, params):
        e.mul_(momentum).add_(p.detach(), alpha=1 - momentum)
    return ema_params


@torch.no_grad()
def _nll(model, inputs, targets, rng):
    """script :27-31 with ONE weight sample: nll = -mean(sum(preds * targets, axis=1))."""
    xs = inputs.unsqueeze(0)
    logp, kl, _ = model._predict(model.params(), xs, rng=rng, mc_samples=1)
    nll = -(logp[0] * targets).sum(dim=1).mean()
    return logp[0], nll, kl[0]


@torch.no_grad()
def evaluate(model, data_loader, nsamples, rng, kl_coef=0.0, num_classes=10):
    """script :78-105.  `data_loader` yields (inputs NHWC float, integer labels) already on the model's device; `rng` is an
    integer seed advanced twice per batch (the script draws `rng_generator.next()` for `accuracy` and for `_nll`).
    Returns (accuracy, logits [N,C], one-hot labels [N,C], nll / n_total, kl / n_total, wts [nbatches * T, P]) -- the sums of the
    per-batch MEAN nll and of kl are divided by the number of images, as the script does."""
    n_total = n_correct = 0
    nll = kl = 0.0
    logits, labels, wts = [], [], []
    seed = int(rng)
    for inputs, y in data_loader:
        targets = torch.nn.functional.one_hot(y.long(), num_classes).to(inputs.dtype)
        c, n, lg, w = accuracy(model, (inputs, targets), nsamples, seed)
        _, bnll, bkl = _nll(model, inputs, targets, seed + 1)
        seed += 2
        n_correct += c
        n_total += n
        nll += float(bnll)
        kl += float(bkl)
        logits.append(lg.cpu().numpy())
        labels.append(targets.cpu().numpy())
        wts.append(w.cpu().numpy())
    return (n_correct / n_total, np.concatenate(logits, 0), np.concatenate(labels, 0), nll / n_total, kl / n_total,
            np.concatenate(wts, 0))


# ------------------------------------------------------------------------------------------------
# calibration scores (numpy, host): brax/utils/utils.py:122-307
# ------------------------------------------------------------------------------------------------
def get_calibration(y, p_mean, num_bins=10):
    """utils.py:122-186: reliability diagram over `num_bins` equal confidence bins [lo, hi) and the ECE weighted by bin counts."""
    p_mean = np.asarray(p_mean)
    pred = p_mean.argmax(1)
    conf = p_mean.max(1)
    true = np.asarray(y).argmax(1)
    edges = np.linspace(0, 1, num_bins + 1)
    count = np.zeros(num_bins)
    mean_conf = np.full(num_bins, np.nan)
    acc = np.full(num_bins, np.nan)
    for b in range(num_bins):
        sel = (conf >= edges[b]) & (conf < edges[b + 1])
        count[b] = sel.sum()
        if count[b] > 0:
            mean_conf[b] = conf[sel].mean()
            acc[b] = (pred[sel] == true[sel]).mean()
    used = count > 0
    ece = 0.0
    if count.sum() > 0:
        wgt = count[used].astype(np.float32) / count[used].sum()
        ece = np.average(np.abs(mean_conf[used] - acc[used]), weights=wgt)
    return {"reliability_diag": (mean_conf, acc), "ece": ece, "nb_items": count / count.sum()}


def _bin_stats(conf, pred, true, bin_size):
    """utils.py:188-211 for every bin (lo, hi]: (accuracy, mean confidence, count); empty bins give zeros."""
    uppers = np.arange(bin_size, 1 + bin_size, bin_size)
    conf, pred, true = np.asarray(conf), np.asarray(pred), np.asarray(true)
    out = []
    for hi in uppers:
        sel = (conf > hi - bin_size) & (conf <= hi)
        k = int(sel.sum())
        out.append((0.0, 0.0, 0) if k < 1 else (float((pred[sel] == true[sel]).sum()) / k, float(conf[sel].sum() / k), k))
    return out


def ECE(conf, pred, true, bin_size=0.1):
    """utils.py:213-236."""
    n = len(conf)
    return sum(abs(a - c) * k / n for a, c, k in _bin_stats(conf, pred, true, bin_size))


def MCE(conf, pred, true, bin_size=0.1):
    """utils.py:238-248."""
    return max(abs(a - c) for a, c, _ in _bin_stats(conf, pred, true, bin_size))


def brier_score(targets, probs):
    """utils.py:250-255."""
    targets, probs = np.asarray(targets), np.asarray(probs)
    if targets.ndim == 1:
        targets, probs = targets[..., None], probs[..., None]
    return np.mean(np.sum((probs - targets) ** 2, axis=1))


def score_model(probs, y_true, verbose=False, normalize=False, bins=15):
    """utils.py:258-307 -> (accuracy %, error %, ece, mce, brier, log loss).  As in the reference the Brier term is taken between
    the integer class label and the probability assigned to it, and NaN probabilities count as 0 in the log loss."""
    probs = np.array(probs, dtype=np.float64)
    preds = probs.argmax(1)
    true = np.asarray(y_true).argmax(1)
    confs = probs.max(1) / probs.sum(1) if normalize else probs.max(1)
    accuracy = float((preds == true).mean()) * 100
    error = 100 - accuracy
    ece = ECE(confs, preds, true, bin_size=1 / bins)
    mce = MCE(confs, preds, true, bin_size=1 / bins)
    probs[np.isnan(probs)] = 0.0
    # sklearn.metrics.log_loss: rows renormalised, probabilities clipped to [eps, 1 - eps] with the dtype's eps
    eps = np.finfo(probs.dtype).eps
    q = np.clip(probs / probs.sum(1, keepdims=True), eps, 1 - eps)
    loss = float(-np.log(q[np.arange(len(true)), true]).mean())
    p_true = probs[np.arange(len(true)), true]
    brier = brier_score(true, p_true)
    if verbose:
        print("Accuracy:", accuracy, "Error:", error, "ECE:", ece, "MCE:", mce, "Loss:", loss, "brier:", brier)
    return accuracy, error, ece, mce, brier, loss
