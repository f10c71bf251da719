"""Generate tests/golden/normalizer.npz from the UNMODIFIED reference module rsl_rl/modules/normalizer.py (container only; the
file is loaded on its own, it only needs torch): three training-mode forwards and one eval-mode forward, without ``until``
(case a) and with learning stopped after the second batch (case b).

    python tests/golden/make_normalizer_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import normalizer_oracle as no  # noqa: E402

REF = "/root/reference/rsl_rl/rsl_rl/modules/normalizer.py"
CASES = {"a": dict(n=96, o=48, seed=0, until=None), "b": dict(n=64, o=235, seed=1, until=100)}


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_normalizer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.EmpiricalNormalization


def main():
    Ref = load_reference()
    out = {}
    for tag, c in CASES.items():
        m = Ref(shape=[c["o"]], until=c["until"])
        m.train()
        xs = no.batches(c["seed"], c["n"], c["o"], 4)
        for s, x in enumerate(xs):
            if s == 3:
                m.eval()
            y = m(x)
            out[f"{tag}__s{s}__out"] = y.numpy()
            for k in ("_mean", "_var", "_std"):
                out[f"{tag}__s{s}__{k}"] = getattr(m, k).numpy().copy()
            out[f"{tag}__s{s}__count"] = np.int64(int(m.count))
    path = os.path.join(ROOT, "tests", "golden", "normalizer.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
