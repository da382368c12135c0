"""Golden values for the calibration scores produced by the REFERENCE's own functions.

    python tests/golden/make_golden_metrics.py      (needs /root/reference; the output .npz is committed)

brax/utils/utils.py imports jax at module level (absent here), so the six pure-numpy functions under test
(get_calibration, compute_acc_bin, ECE, MCE, brier_score, score_model; utils.py:122-307) are located with `ast` in the
reference file and executed unmodified in a namespace that holds numpy and scikit-learn's `metrics` / `log_loss`, which
is all they use."""
import ast
import io
import os
from contextlib import redirect_stdout

import numpy as np
from sklearn import metrics
from sklearn.metrics import log_loss

SRC = "/root/reference/brax/utils/utils.py"
WANT = ["get_calibration", "compute_acc_bin", "ECE", "MCE", "brier_score", "score_model"]


def load():
    text = open(SRC).read()
    tree = ast.parse(text)
    ns = {"np": np, "metrics": metrics, "log_loss": log_loss}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    return ns


def main():
    ns = load()
    out = {}
    rng = np.random.RandomState(0)
    for tag, (n, c, sharp) in {"a": (200, 10, 3.0), "b": (57, 4, 0.5), "c": (1000, 10, 8.0)}.items():
        logits = rng.randn(n, c) * sharp
        labels = rng.randint(0, c, n)
        flip = rng.rand(n) < 0.6
        logits[np.arange(n)[flip], labels[flip]] += sharp * 1.5
        p = np.exp(logits - logits.max(1, keepdims=True))
        p /= p.sum(1, keepdims=True)
        onehot = np.eye(c)[labels]
        with redirect_stdout(io.StringIO()):
            cal = ns["get_calibration"](onehot, p, num_bins=10)
            sc = ns["score_model"](p, onehot, bins=15)
        conf, pred = p.max(1), p.argmax(1)
        out[f"{tag}_p"], out[f"{tag}_onehot"] = p, onehot
        out[f"{tag}_cal_ece"] = np.float64(cal["ece"])
        out[f"{tag}_cal_conf"], out[f"{tag}_cal_acc"] = cal["reliability_diag"]
        out[f"{tag}_cal_items"] = cal["nb_items"]
        out[f"{tag}_score"] = np.asarray(sc, dtype=np.float64)
        out[f"{tag}_ece10"] = np.float64(ns["ECE"](conf, pred, labels, bin_size=0.1))
        out[f"{tag}_mce10"] = np.float64(ns["MCE"](conf, pred, labels, bin_size=0.1))
        out[f"{tag}_brier"] = np.float64(ns["brier_score"](onehot, p))
        print(tag, "score_model", sc)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
