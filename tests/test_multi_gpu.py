"""Multi-rank correctness ON HARDWARE (parity hole ii): the CUDA stages + NCCL collectives issued by the extension.
Needs >= 2 GPUs (skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).
  * elg_mppi_update with the rollouts of every main env sharded over 2 ranks == the single-rank update on the same data
  * ShardedEpisodeStats fed by the fused reset kernel on each rank's env shard + elg_episode_stats_allreduce ==
    extras["episode"] of ONE env object owning all envs (legged_robot.py:200-213)"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


LOG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _worker(rank, world, port, q):
    os.makedirs(LOG_DIR, exist_ok=True)
    log = open(os.path.join(LOG_DIR, f"multi_gpu_rank{rank}.log"), "w")

    def say(msg):
        log.write(msg + "\n")
        log.flush()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    say("process group up")
    try:
        import common
        from extended_legged_gym_b200 import _lib, synthetic
        from extended_legged_gym_b200.envs import LeggedRobot
        from extended_legged_gym_b200.sim_backend import SyntheticSim
        from extended_legged_gym_b200.utils.distributed import ElgComm, ShardedEpisodeStats, shard_range
        from extended_legged_gym_b200.utils.mppi import mppi_update, mppi_update_native
        comm = ElgComm(dev)
        say("ElgComm up")
        # ---- MPPI: same data on every rank (seeded), each takes its share of the samples
        g = torch.Generator().manual_seed(3)
        M, S, T, K, D = 16, 96, 20, 5, 12
        r = torch.randn(M, S, T, generator=g).to(dev)
        u = torch.randn(M, S, K, D, generator=g).to(dev)
        lo, hi = shard_range(S, rank, world)
        got = mppi_update(r[:, lo:hi].contiguous(), u[:, lo:hi].contiguous(), 0.05, comm=comm)
        full = mppi_update_native(r, u, 0.05, comm=None)    # single rank, no collective
        ok_mppi = torch.allclose(got, full, rtol=1e-4, atol=1e-5)
        say(f"sharded mppi_update: {ok_mppi}")
        # inside a CUDA graph as well (collectives on the capture stream)
        s = torch.cuda.Stream(device=dev)
        gr = torch.cuda.CUDAGraph()
        rs, us = r[:, lo:hi].contiguous(), u[:, lo:hi].contiguous()
        with torch.cuda.stream(s):
            mppi_update(rs, us, 0.05, comm=comm)
            s.synchronize()
            with torch.cuda.graph(gr, stream=s, capture_error_mode="thread_local"):     # (NCCL's helper threads call CUDA too)
                out_g = mppi_update(rs, us, 0.05, comm=comm)
            gr.replay()
            s.synchronize()
        ok_graph = torch.allclose(out_g, full, rtol=1e-4, atol=1e-5)
        say(f"graph-captured sharded mppi_update: {ok_graph}")
        # ---- episode statistics: N envs in one object vs two shards, fused (in-kernel) reset path, same uniforms
        N, case = 512, "anymal_c_rough"
        cfg, spec, st = common.make_case_state(case, N, seed=4, adversarial=True)
        hf = synthetic.make_height_field(seed=0)
        U = torch.rand(N, _lib.RESET_UNIFORMS, generator=torch.Generator().manual_seed(9))

        def build(lo, hi):
            c, sp, _ = common.make_case_state(case, 4, seed=0)
            c.env.num_envs = hi - lo
            c.domain_rand.push_robots = False
            c.terrain.curriculum = False
            sub = {k: (v.view(N, -1)[lo:hi].reshape(-1, *v.shape[1:]) if v.shape[0] != N else v[lo:hi]).clone() for k, v in st.items()}
            torch.manual_seed(1234)
            env = LeggedRobot(c, None, SyntheticSim(c, hi - lo, dev, spec=sp, height_samples=hf, state=sub), dev, True)
            env.set_env_state(sub)
            env.reset_uniforms = U[lo:hi].to(dev).contiguous()
            env.noise_u = None
            env.add_noise = False
            return env
        whole = build(0, N)
        whole.episode_stats = ShardedEpisodeStats(dev)              # no comm: local totals of the one-object run
        lo, hi = shard_range(N, rank, world)
        part = build(lo, hi)
        part.episode_stats = ShardedEpisodeStats(dev, comm=comm)
        for env in (whole, part):
            for _ in range(3):
                env.torques = env._compute_torques(env.actions).view(env.torques.shape)
                env.post_physics_step()
        want = whole.episode_stats.reduce(whole.max_episode_length_s, names=list(whole.episode_sums))
        got_s = part.episode_stats.reduce(part.max_episode_length_s, names=list(part.episode_sums))
        torch.cuda.synchronize()
        ok_stats = int(want["num_resets"]) > 0 and int(got_s["num_resets"]) == int(want["num_resets"])
        for k, v in want.items():
            ok_stats = ok_stats and abs(float(got_s[k]) - float(v)) <= 1e-5 * abs(float(v)) + 1e-6
        say(f"sharded episode statistics: {ok_stats}")
        q.put((rank, bool(ok_mppi), bool(ok_graph), bool(ok_stats)))
        comm.close()
    except Exception as ex:  # noqa: BLE001
        import traceback
        say("FAILED: " + traceback.format_exc())
        q.put((rank, False, False, False))
    finally:
        say("leaving")
        log.close()
        os._exit(0)          # no collective tear-down: a rank that failed must not leave its peer waiting in destroy_process_group


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_mppi_and_episode_stats_two_ranks_nccl():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=240) for _ in procs]
    finally:
        for p in procs:          # never leave a rank behind (a hung child would block the interpreter's exit)
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    for rank, ok_mppi, ok_graph, ok_stats in res:
        assert ok_mppi, f"rank {rank}: sharded elg_mppi_update differs from the single-rank update"
        assert ok_graph, f"rank {rank}: graph-captured sharded update differs"
        assert ok_stats, f"rank {rank}: sharded episode statistics differ from the one-object run"
