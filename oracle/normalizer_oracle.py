"""CPU oracle of the observation normaliser -- TEST INFRASTRUCTURE, never on the product path.

Restates ``EmpiricalNormalization.forward`` / ``update`` (rsl_rl/modules/normalizer.py:43-75 in /root/reference/rsl_rl) on a plain
state dict ``{"mean", "var", "std": [1, O] fp32, "count": int}``.  Pinned by ``tests/test_normalizer.py`` against the UNMODIFIED
reference module (container only) and ``tests/golden/normalizer.npz`` (``tests/golden/make_normalizer_golden.py``).
"""
import torch


def new_state(num_obs):
    return {"mean": torch.zeros(1, num_obs), "var": torch.ones(1, num_obs), "std": torch.ones(1, num_obs), "count": 0}


def update(st, x, until=None):
    if until is not None and st["count"] >= until:
        return
    n = x.shape[0]
    st["count"] += n
    rate = torch.tensor(n, dtype=torch.float32) / torch.tensor(st["count"], dtype=torch.long)       # int / LongTensor -> fp32
    var_x = torch.var(x, dim=0, unbiased=False, keepdim=True)
    mean_x = torch.mean(x, dim=0, keepdim=True)
    delta = mean_x - st["mean"]
    st["mean"] = st["mean"] + rate * delta
    st["var"] = st["var"] + rate * (var_x - st["var"] + delta * (mean_x - st["mean"]))
    st["std"] = torch.sqrt(st["var"])


def forward(st, x, eps=1e-2, until=None, training=True):
    if training:
        update(st, x, until)
    return (x - st["mean"]) / (st["std"] + eps)


def batches(seed, n, o, steps):
    """seeded observation-like batches: columns of very different scale, one constant column, one zero column"""
    g = torch.Generator().manual_seed(seed)
    scale = torch.logspace(-2, 1.5, o)
    shift = torch.linspace(-3, 3, o)
    out = []
    for s in range(steps):
        x = torch.randn(n, o, generator=g) * scale + shift * (1 + 0.1 * s)
        x[:, 3] = 0.7
        x[:, 5] = 0.0
        out.append(x.clamp(-100, 100))
    return out
