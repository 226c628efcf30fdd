"""TEST INFRASTRUCTURE ONLY — numpy float64 restatement of the reference's two-layer MLP
(modules/modeling.py:788-807: Linear -> BatchNorm1d (training mode) -> ReLU -> Linear) with its
analytic backward, pinned against the reference class itself (tests/golden/mlp.npz written by
oracle/gen_golden.py).  Imported by tests/ and tests/multi_gpu_check.py only."""
import numpy as np


def forward(x, W1, b1, gamma, beta, W2, b2, eps=1e-5, running=None):
    """running = (mean, var): eval mode.  Returns (y, cache)."""
    x = np.asarray(x, np.float64)
    h = x @ np.asarray(W1, np.float64).T + np.asarray(b1, np.float64)
    if running is None:
        mean, var = h.mean(0), h.var(0)
    else:
        mean, var = np.asarray(running[0], np.float64), np.asarray(running[1], np.float64)
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (h - mean) * invstd
    z = xhat * np.asarray(gamma, np.float64) + np.asarray(beta, np.float64)
    a = np.maximum(z, 0.0)
    y = a @ np.asarray(W2, np.float64).T + np.asarray(b2, np.float64)
    return y, dict(x=x, h=h, mean=mean, var=var, invstd=invstd, xhat=xhat, z=z, a=a)


def running_stats(cache, rm, rv, momentum=0.1):
    """nn.BatchNorm1d's update: biased batch variance normalises, the unbiased one is tracked."""
    n = cache["h"].shape[0]
    new_rm = (1 - momentum) * np.asarray(rm, np.float64) + momentum * cache["mean"]
    new_rv = (1 - momentum) * np.asarray(rv, np.float64) + momentum * cache["var"] * n / max(n - 1, 1)
    return new_rm, new_rv


def backward(dy, cache, W1, gamma, W2):
    dy = np.asarray(dy, np.float64)
    W1, W2, gamma = (np.asarray(t, np.float64) for t in (W1, W2, gamma))
    n = dy.shape[0]
    dW2 = dy.T @ cache["a"]
    db2 = dy.sum(0)
    da = dy @ W2
    dz = da * (cache["z"] > 0)
    dbeta = dz.sum(0)
    dgamma = (dz * cache["xhat"]).sum(0)
    dh = gamma * cache["invstd"] * (dz - dbeta / n - cache["xhat"] * dgamma / n)
    dW1 = dh.T @ cache["x"]
    db1 = dh.sum(0)
    dx = dh @ W1
    return dict(dx=dx, dW1=dW1, db1=db1, dgamma=dgamma, dbeta=dbeta, dW2=dW2, db2=db2, da=da)


def relu_flip_slack(cache, grads, W1, gamma, width):
    """ReLU is not smooth: an implementation whose pre-activations differ from this one by up to `width`
    may take the other branch for elements with |z| < width, and each such element moves dZ by the full
    upstream value.  Upper bounds (L2) of what those elements can contribute to dx, dW1, dgamma, dbeta —
    the slack a comparison against another precision has to grant on top of its relative tolerance."""
    amb = np.abs(cache["z"]) < width
    r, c = np.nonzero(amb)
    w = np.abs(grads["da"][r, c]) * np.abs(np.asarray(gamma, np.float64)[c] * cache["invstd"][c])
    W1 = np.asarray(W1, np.float64)
    return dict(count=int(amb.sum()),
                dx=float(np.sum(w * np.linalg.norm(W1[c], axis=1))),
                dW1=float(np.sum(w * np.linalg.norm(cache["x"][r], axis=1))),
                dgamma=float(np.sum(np.abs(grads["da"][r, c] * cache["xhat"][r, c]))),
                dbeta=float(np.sum(np.abs(grads["da"][r, c]))))


def visual_tail(temporal, original):
    """modules/module_cross.py:207-213 in float64: (visual_output, per-row norms)."""
    h = np.asarray(original, np.float64) + (np.asarray(temporal, np.float64) if temporal is not None else 0.0)
    n = np.linalg.norm(h, axis=-1, keepdims=True)
    return (h / n).mean(1), h, n


def visual_tail_backward(g, h, n):
    g = np.asarray(g, np.float64)[:, None, :]
    hh = h / n
    return (g - hh * (hh * g).sum(-1, keepdims=True)) / (n * h.shape[1])
