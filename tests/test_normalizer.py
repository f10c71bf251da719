"""Observation normaliser + storage write (rsl_rl/modules/normalizer.py:14-79, rollout_storage.py:95-100).

not gpu: the oracle against the golden vectors generated from the UNMODIFIED reference module and (container only) against that
module itself; ABI checks.  gpu: elg_normalize_observations / the EmpiricalNormalization host class against the golden vectors and
against the oracle at BASELINE's 4096 x 235 and 65 536 x 48, in-place and storage-slot destinations, `until`, eval mode,
state_dict interchange, run-to-run bit reproducibility.

Tolerances: state (mean / var / std) rtol 1e-5, atol 1e-6.  The output (x - mean) / (std + eps) inherits the state's tolerance,
dm = 1e-6 + 1e-5 |mean| and ds = 1e-6 + 1e-5 std, amplified by the division:  |d out| <= (dm + |out| ds) / (std + eps) + 1e-5 |out|
(for a constant column, std = 0 and eps = 0.01, that is 1e-4 per unit of dm)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from extended_legged_gym_b200 import _lib  # noqa: E402
from oracle import normalizer_oracle as no  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_normalizer_golden as mk  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "normalizer.npz")
DEV = "cuda:0"
RTOL, ATOL, EPS = 1e-5, 1e-6, 1e-2


def assert_state(mean, var, std, count, z, tag, s):
    for got, k in ((mean, "_mean"), (var, "_var"), (std, "_std")):
        torch.testing.assert_close(got.cpu().view(1, -1), torch.from_numpy(z[f"{tag}__s{s}__{k}"]), rtol=RTOL, atol=ATOL, msg=lambda m, k=k: f"{k} step {s}: {m}")
    assert int(count) == int(z[f"{tag}__s{s}__count"])


def assert_out(got, want, std, mean):
    std, mean = std.cpu().view(1, -1), mean.cpu().view(1, -1)
    tol = ((ATOL + RTOL * mean.abs()) + want.abs() * (ATOL + RTOL * std)) / (std + EPS) + RTOL * want.abs()
    bad = (got.cpu() - want).abs() > tol
    assert not bool(bad.any()), f"{int(bad.sum())} outputs off, worst {float(((got.cpu() - want).abs() / tol).max()):.2f} x tolerance"


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_fixture(tag):
    z, c = np.load(GOLDEN), mk.CASES[tag]
    st = no.new_state(c["o"])
    for s, x in enumerate(no.batches(c["seed"], c["n"], c["o"], 4)):
        y = no.forward(st, x, EPS, c["until"], training=s < 3)
        assert_state(st["mean"], st["var"], st["std"], st["count"], z, tag, s)
        assert_out(y, torch.from_numpy(z[f"{tag}__s{s}__out"]), st["std"], st["mean"])


@pytest.mark.skipif(not os.path.exists(mk.REF), reason="the reference checkout is only present in the build container")
def test_oracle_matches_live_reference_module():
    m = mk.load_reference()(shape=[30], until=500)
    m.train()
    st = no.new_state(30)
    for x in no.batches(5, 200, 30, 4):
        y, y2 = m(x), no.forward(st, x, EPS, 500)
        assert torch.equal(y, y2) and torch.equal(m._mean, st["mean"]) and torch.equal(m._var, st["var"]) and int(m.count) == st["count"]
    assert st["count"] == 600      # the fourth batch no longer learns


def test_normalizer_abi_argument_checks():
    lib = _lib.load()
    assert lib.elg_normalizer_scratch_bytes(4096, 235) == 256 + 16 * 235 + 16 * 235 * 128      # (mean, M2) doubles per CTA of the single-launch form
    assert lib.elg_normalizer_scratch_bytes(100, 48) == 256 + 16 * 48 + 16 * 48 * 128
    assert lib.elg_normalizer_scratch_bytes(10_000_000, 48) > 256 + 4 * 48 * (4 + 3 * 31)
    call = lib.elg_normalize_observations
    assert call(8, 0, 16, 16, 16, 16, 16, 0.01, -1, 1, 16, 16, None, None, None, None, None) == -1
    assert call(8, 4, None, 16, 16, 16, 16, 0.01, -1, 1, 16, 16, None, None, None, None, None) == -4
    assert call(8, 4, 16, 16, 16, 16, 16, 0.01, -1, 1, 16, None, None, None, None, None, None) == -4         # training without scratch
    assert call(8, 4, 16, 16, 16, 16, 16, 0.01, -1, 1, 16, 24, None, None, None, None, None) == -1           # misaligned scratch
    assert call(8, 4, 16, 16, 16, 16, 16, 0.01, -1, 0, 16, None, None, 16, None, None, None) == -4           # reward destination without source
    assert call(0, 4, 16, 16, 16, 16, 16, 0.01, -1, 1, 16, 16, None, None, None, None, None) == -1           # cannot learn from an empty batch
    assert call(0, 4, 16, 16, 16, 16, 16, 0.01, -1, 0, 16, None, None, None, None, None, None) == 0


# ---------------------------------------------------------------------------------------------------------------
def make_module(o, until=None):
    from extended_legged_gym_b200.utils.normalizer import EmpiricalNormalization
    return EmpiricalNormalization(shape=[o], until=until).to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_kernel_matches_reference_fixture(tag):
    z, c = np.load(GOLDEN), mk.CASES[tag]
    m = make_module(c["o"], c["until"])
    m.train()
    for s, x in enumerate(no.batches(c["seed"], c["n"], c["o"], 4)):
        if s == 3:
            m.eval()
        y = m(x.to(DEV))
        torch.cuda.synchronize()
        assert_state(m._mean, m._var, m._std, m.count, z, tag, s)
        assert_out(y, torch.from_numpy(z[f"{tag}__s{s}__out"]), m._std, m._mean)


@pytest.fixture(params=[0, 1, 2, 4, 8, 12, 16], ids=["default", "two_launches", "grid_handover", "cluster_of_1", "cluster_of_2", "cluster_of_4", "cluster_of_8"])
def launch_form(request):
    """the forms of the training-mode call: 0 = column-parallel single launch (rows of a column group split over a thread-block
    cluster, partial sums through distributed shared memory) up to 65 536 rows, else the statistics + apply pair; 1 = always the
    pair; 2 = row-parallel single launch with a grid-wide hand-over (A/B form); 4 / 8 / 12 / 16 = form 0 with the cluster size forced"""
    lib = _lib.load()
    lib.elg_set_normalizer_tuning(request.param)
    yield request.param
    lib.elg_set_normalizer_tuning(0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,o", [(4096, 235), (65536, 48), (4099, 235), (31, 300), (1, 7), (32832, 235), (16384, 235), (300, 1024)])
def test_kernel_matches_oracle_full_size_and_storage_slot(n, o, launch_form):
    m = make_module(o, until=int(1e8))
    m.train()
    st = no.new_state(o)
    storage = torch.zeros(3, n, o, device=DEV)                 # rollout_storage.observations
    rewards, dones = torch.zeros(3, n, 1, device=DEV), torch.zeros(3, n, 1, device=DEV, dtype=torch.uint8)
    g = torch.Generator().manual_seed(n)
    for s, x in enumerate(no.batches(n % 97, n, o, 3)):
        rew, done = torch.randn(n, generator=g), torch.rand(n, generator=g) < 0.1
        xd = x.to(DEV)
        got = m.forward_into(xd, storage[s], rew.to(DEV), rewards[s], done.to(DEV), dones[s])
        want = no.forward(st, x, EPS, int(1e8))
        torch.cuda.synchronize()
        assert got.data_ptr() == storage[s].data_ptr() and torch.equal(xd.cpu(), x), "the input must stay untouched"
        torch.testing.assert_close(m._mean.cpu(), st["mean"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(m._var.cpu(), st["var"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(m._std.cpu(), st["std"], rtol=RTOL, atol=ATOL)
        assert int(m.count) == st["count"]
        assert_out(storage[s], want, m._std, m._mean)
        assert torch.equal(rewards[s].cpu().view(-1), rew) and torch.equal(dones[s].cpu().view(-1).bool(), done)
    # in place, eval mode: state frozen, same arithmetic
    m.eval()
    x = no.batches(3, n, o, 1)[0]
    xd = x.to(DEV)
    before = (m._mean.clone(), m._var.clone(), int(m.count))
    m.forward_into(xd, xd)
    torch.cuda.synchronize()
    assert torch.equal(m._mean, before[0]) and torch.equal(m._var, before[1]) and int(m.count) == before[2]
    assert torch.equal(xd.cpu(), ((x.to(DEV) - m._mean) / (m._std + EPS)).cpu()), "eval-mode output is not bit-identical to the torch expression"
    torch.testing.assert_close(m.inverse(xd).cpu(), x, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_update_only_until_and_reproducibility(launch_form):
    n, o = 4096, 235
    xs = [x.to(DEV) for x in no.batches(2, n, o, 3)]
    runs = []
    for _ in range(2):
        m = make_module(o, until=5000)
        m.train()
        m.update(xs[0])                       # statistics only
        y1 = m(xs[1])                         # count 4096 < 5000: learns -> 8192
        y2 = m(xs[2])                         # count 8192 >= 5000: frozen, no host read involved
        torch.cuda.synchronize()
        assert int(m.count) == 8192
        runs.append((y1.clone(), y2.clone(), m._mean.clone(), m._var.clone()))
    for a, b in zip(*runs):
        assert torch.equal(a, b), "two identical runs differ bitwise"
    st = no.new_state(o)
    no.update(st, xs[0].cpu(), 5000)
    no.forward(st, xs[1].cpu(), EPS, 5000)
    want = no.forward(st, xs[2].cpu(), EPS, 5000)
    assert st["count"] == 8192
    assert_out(runs[0][1], want, st["std"], st["mean"])


@pytest.mark.gpu
def test_state_dict_interchanges_with_the_reference_layout():
    m = make_module(12)
    m.train()
    m(torch.randn(64, 12, device=DEV))
    sd = m.state_dict()
    assert sorted(sd) == ["_mean", "_std", "_var", "count"] and sd["_mean"].shape == (1, 12) and sd["count"].dtype == torch.long
    m2 = make_module(12)
    m2.load_state_dict(sd)
    m2.eval()
    x = torch.randn(5, 12, device=DEV)
    assert torch.equal(m2(x), (x - m._mean) / (m._std + EPS))
    assert m.mean.shape == (12,) and m.std.shape == (12,)
    with pytest.raises(ValueError):
        m(torch.randn(5, 13, device=DEV))
    with pytest.raises(ValueError):
        m(torch.randn(5, 12))
