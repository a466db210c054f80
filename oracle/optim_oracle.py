"""TEST INFRASTRUCTURE (oracle): numpy restatement of the optimizer steps the reference applies to its embedding tables.

The reference builds ``torch.optim`` optimizers by name (recbole ``Trainer._build_optimizer`` [recbole-1.0.1]: learner in
adam / sgd / adagrad / rmsprop / sparse_adam, ``overall.yaml:20-21``).  The functions below restate the published update
rules of ``torch.optim.SGD`` (momentum 0), ``torch.optim.Adagrad`` (lr_decay 0, weight_decay 0,
initial_accumulator_value 0) and ``torch.optim.SparseAdam`` for ONE step on a dense ``[N, D]`` gradient whose untouched
rows are zero; ``tests/test_optim_oracle.py`` pins them against the torch implementations themselves.  Only ``tests/``
may import this module."""
import numpy as np


def sgd_step(w, g, lr):
    """torch.optim.SGD.step, momentum = dampening = weight_decay = 0:  w <- w - lr g."""
    return (w - np.float32(lr) * g).astype(np.float32)


def adagrad_step(w, s, g, lr, eps=1e-10):
    """torch.optim.Adagrad.step (_single_tensor_adagrad), lr_decay = 0:  s <- s + g*g;  w <- w - lr g / (sqrt(s) + eps).
    Rows with g == 0 are unchanged, so the dense rule restricted to touched rows is the dense rule."""
    s = (s + g * g).astype(np.float32)
    w = (w - np.float32(lr) * g / (np.sqrt(s) + np.float32(eps))).astype(np.float32)
    return w, s


def sparse_adam_step(w, m, v, g, touched, t, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.SparseAdam.step (_functional.sparse_adam) at optimizer step ``t`` (1-based) on the rows in ``touched``:
    m <- m + (1-b1)(g-m);  v <- v + (1-b2)(g*g-v);  w <- w - lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)."""
    a1, a2 = np.float32(1.0 - float(betas[0])), np.float32(1.0 - float(betas[1]))  # (1 - beta) formed in double, as torch does
    w, m, v = w.copy(), m.copy(), v.copy()
    r = np.unique(np.asarray(touched))
    m[r] = m[r] + a1 * (g[r] - m[r])
    v[r] = v[r] + a2 * (g[r] * g[r] - v[r])
    step_size = np.float32(lr * np.sqrt(1.0 - float(betas[1]) ** t) / (1.0 - float(betas[0]) ** t))
    w[r] = w[r] - step_size * m[r] / (np.sqrt(v[r]) + np.float32(eps))
    return w.astype(np.float32), m.astype(np.float32), v.astype(np.float32)
