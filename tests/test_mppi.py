"""MPPI cost-weighted update and the multi-GPU reductions (section 8e).
not-gpu: oracle against the fixture generated from the reference's in-tree update; the sharded update and the episode
statistics over a 2-rank gloo group on CPU (host-side plumbing, the oracle's local stages standing in for the
kernels).  gpu: the three CUDA stages against the oracle and the fixture."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import mppi_oracle as mo, ref_harness  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mppi.npz")
DEV = "cuda:0"


def case(tag):
    z = np.load(GOLDEN)
    return (torch.from_numpy(z[f"{tag}__rewards"]), torch.from_numpy(z[f"{tag}__samples"]), float(z[f"{tag}__temp"][0]),
            torch.from_numpy(z[f"{tag}__mean_traj"]))


class OracleOps:
    costs = staticmethod(mo.local_costs)
    partials = staticmethod(mo.local_partials)
    finish = staticmethod(mo.finish)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_fixture(tag):
    r, u, temp, want = case(tag)
    assert torch.equal(mo.mppi_update(r, u, temp), want)
    # the three-stage split used for sharding is the same update
    from extended_legged_gym_b200.utils.mppi import mppi_update
    got = mppi_update(r, u, temp, ops=OracleOps)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.available(), reason="needs /root/reference")
def test_oracle_matches_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_mppi_golden as mg
    g = torch.Generator().manual_seed(5)
    r, u = torch.randn(1, 50, 8, generator=g), torch.randn(1, 50, 4, 3, generator=g)
    assert torch.equal(mo.mppi_update(r, u, 0.07)[0], mg.reference_update(r[0], u[0], 0.07))


def _worker(rank, world, port, tag, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from extended_legged_gym_b200.utils.distributed import ShardedEpisodeStats, shard_range
        from extended_legged_gym_b200.utils.mppi import mppi_update
        r, u, temp, want = case(tag)
        S = r.shape[1] - r.shape[1] % world
        lo, hi = shard_range(S, rank, world)
        got = mppi_update(r[:, lo:hi], u[:, lo:hi], temp, ops=OracleOps)
        full = mo.mppi_update(r[:, :S], u[:, :S], temp)
        ok_mppi = torch.allclose(got, full, rtol=1e-5, atol=1e-6)
        # episode statistics: 10 envs split 2 ways, different resets per rank; SoA [terms, N] like LeggedRobot._episode_sums_all
        from extended_legged_gym_b200 import _lib
        N, nt = 10, _lib.NUM_REWARD_TERMS
        full_sums = torch.zeros(nt, N)
        full_sums[_lib.TERM_ID["torques"]] = torch.arange(N, dtype=torch.float) * 1.5
        full_sums[_lib.TERM_ID["dof_acc"]] = 2.0
        lo, hi = shard_range(N, rank, world)
        resets_global = torch.tensor([1, 2, 7, 8, 9])
        mine = resets_global[(resets_global >= lo) & (resets_global < hi)] - lo
        st = ShardedEpisodeStats("cpu", use_torch_distributed=True)
        st.accumulate(full_sums[:, lo:hi].clone(), mine)
        out = st.reduce(20.0, terrain_levels=torch.arange(lo, hi), names=["torques", "dof_acc"])
        want_a = full_sums[_lib.TERM_ID["torques"]][resets_global].mean() / 20.0
        ok_stats = (abs(float(out["rew_torques"]) - float(want_a)) < 1e-6 and abs(float(out["rew_dof_acc"]) - 0.1) < 1e-6 and
                    int(out["num_resets"]) == 5 and abs(float(out["terrain_level"]) - 4.5) < 1e-6 and float(st.buf.abs().sum()) == 0.0)
        q.put((rank, bool(ok_mppi), bool(ok_stats)))
    finally:
        dist.destroy_process_group()


def test_sharded_update_and_episode_stats_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "a", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def test_shard_range_and_abi_checks():
    from extended_legged_gym_b200 import _lib
    from extended_legged_gym_b200.utils.distributed import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_range(65536, r, 8) for r in range(8)][-1] == (57344, 65536)
    lib = _lib.load()
    assert lib.elg_mppi_partials(None, 1, 4, 3, 2, None, 5, 0.05, None, None) == -1     # local range outside the total
    assert lib.elg_mppi_partials(None, 1, 4, 0, 2, None, 5, 0.0, None, None) == -1      # temperature must be > 0
    assert lib.elg_mppi_costs(None, 1, 1, 1, None, None) == -4


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_mppi_kernels_match_reference_fixture(tag):
    from extended_legged_gym_b200.utils.mppi import mppi_update
    r, u, temp, want = case(tag)
    got = mppi_update(r.to(DEV), u.to(DEV), temp)
    torch.cuda.synchronize()
    assert got.shape == want.shape and torch.allclose(got.cpu(), want, rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
def test_mppi_kernels_baseline_config_5_and_shards():
    """64 mains x 512 rollouts x horizon 20, 5 nodes x 12 dof; also the sharded evaluation of the same update:
    partials of two halves summed == the full update (what the all-reduce does)."""
    from extended_legged_gym_b200.utils.mppi import _CudaOps, mppi_update
    g = torch.Generator().manual_seed(1)
    M, S, T, K, D, temp = 64, 512, 20, 5, 12, 0.05
    r = (torch.randn(M, S, T, generator=g) * 0.2 + torch.randn(M, S, 1, generator=g)).to(DEV)
    u = torch.randn(M, S, K, D, generator=g).to(DEV)
    got = mppi_update(r, u, temp)
    want = mo.mppi_update(r.cpu(), u.cpu(), temp)
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-5)
    costs = _CudaOps.costs(r)
    p = _CudaOps.partials(costs, 0, u[:, :256], temp) + _CudaOps.partials(costs, 256, u[:, 256:], temp)
    assert torch.allclose(_CudaOps.finish(p, (K, D)), got, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_rollout_batch_glue():
    """rollout_batch: rewards [num_rollout_envs, horizon]; mains untouched, rollouts re-synced at the end."""
    from test_rollout_clone import make_rollout_env
    from extended_legged_gym_b200.utils.mppi import mppi_update, rollout_batch
    m, r, horizon = 4, 8, 3
    env, _ = make_rollout_env(m, r, seed=3)
    env.cfg.domain_rand.push_robots = False
    env.step(torch.zeros(m, env.num_actions, device=DEV))
    main_before = env.root_states[env.main_env_indices].clone()
    us = torch.randn(m * r, horizon, env.num_actions, device=DEV) * 0.3
    rewards = rollout_batch(env, us)
    assert rewards.shape == (m * r, horizon) and bool(torch.isfinite(rewards).all())
    assert torch.equal(env.root_states[env.main_env_indices], main_before)
    rs = env.root_states.view(m, 1 + r, 13)
    assert torch.equal(rs[:, 1:], rs[:, :1].expand(-1, r, -1))
    nodes = us.view(m, r, horizon, env.num_actions)
    traj = mppi_update(rewards.view(m, r, horizon), nodes, 0.05)
    assert traj.shape == (m, horizon, env.num_actions)


@pytest.mark.gpu
def test_rollout_batch_graph_equals_step_by_step():
    """RobotBatchRollout.rollout_batch as one CUDA graph (replayed) == horizon x step_rollout through the public API"""
    import common
    from extended_legged_gym_b200 import synthetic
    from extended_legged_gym_b200.envs import RobotTrajGradSampling
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    mains, rollouts, horizon = 8, 24, 6
    n = mains * (1 + rollouts)

    def build():
        cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=3)
        cfg.env.num_envs, cfg.env.rollout_envs = mains, rollouts
        env = RobotTrajGradSampling(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=synthetic.make_height_field(seed=0), state=st), DEV, True)
        env.set_env_state(st)
        env._cache_main_env_states()
        return env
    a, b = build(), build()
    us = (torch.randn(mains * rollouts, horizon, 12, generator=torch.Generator().manual_seed(1)) * 0.3).to(DEV)
    want = torch.zeros(mains * rollouts, horizon, device=DEV)
    b._sync_main_to_rollout()
    for i in range(horizon):
        _, _, r, _, _ = b.step_rollout(us[:, i])
        want[:, i] = r
    b._sync_main_to_rollout()
    for rep in range(3):                       # capture, then replays: the result must not drift
        got = a.rollout_batch(us).clone()
        torch.cuda.synchronize()
        if rep == 0:
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), "graph-captured rollout_batch differs from the step_rollout loop"
            first = got
    # replays start from the re-synchronised rollouts: with unchanged mains they reproduce the same rewards
    assert torch.equal(got, first)
    for k in ("root_states", "dof_state", "last_actions", "feet_air_time"):
        assert torch.equal(getattr(a, k), getattr(b, k)), f"{k} differs after rollout_batch"
