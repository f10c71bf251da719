#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--envs E] [--config anymal_c_rough]

Metric: env-steps/s of the fused post-physics step.  One "step" = one pass of the hot path over one
batch of synthetic PhysX state = 1x _compute_torques + the post_physics_step body (derive, heading,
187-point height scan, termination, the reward registry, observations with in-kernel noise, history)
-- SURVEY.md section 8d.  Workload at N=1: BASELINE configs[1], anymal_c_rough, 4096 envs, one B200.
For N>1 the envs shard across ranks with no data-path collective (weak scaling: 4096 envs per GPU); the one genuine
reduction of the path -- the episode statistics -- is an NCCL all-reduce issued by the extension on the compute stream as
the LAST NODE OF THE TIMED GRAPH (one per K steps).  `secondary` carries BASELINE configs[2]
(a1 rough, 65 536 envs in total, strong-scaled over the ranks, reset path + statistics all-reduce inside the timed region),
configs[3] (depth camera) and configs[4] (MPPI: 64 mains x 512 rollouts in total, sharded over the ranks).

  value     device-resident throughput: K steps (2 kernels each) + the statistics all-reduce replayed from one CUDA graph,
            CUDA events, max over ranks.  Every step works on a different replica of the state (R replicas, > 2x L2 in
            total) so no step finds its inputs in L2.
  e2e       the same step through the public Python API (LeggedRobot._compute_torques + post_physics_step)
            with the PhysX state in pinned HOST memory: ONE H2D of the packed state + actions block and ONE D2H of the
            packed obs / rew / reset block inside the timed region, every step.
  roofline  step kernel alone (same graph technique), algorithmic bytes of SURVEY.md section 8d over its
            mean launch duration, against MEASURED_PEAKS.json hbm_gbs: cold L2 (`frac`) and steady state (one replica,
            L2-resident, `steady_state`), next to the launch floor measured in the same run (`floor`: an empty PDL kernel and
            a pure bulk-copy round trip of the step's bytes with the step's launch shape).
  cpu_baseline  the oracle port of the reference's torch-CPU implementation on this box's host cores.

--impl reference runs ONLY that CPU implementation (all host threads) and prints the same line shape.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "env-steps/s of fused post-physics step"
UNIT = "env-steps/s"
L2_BYTES = 126 * 1024 * 1024


# ----------------------------------------------------------------------------------------------
def algorithmic_bytes_per_env(D, F, P, T, H, O, R, C_, heading):
    """every distinct tensor element the reference reads / writes, counted once."""
    reads = 4 * (13 + 2 * D + 3 * (T + P + F) + 6 * F + D + D + D + 6 + 6 + C_ + 2 * F + R) + F + 8
    writes = 4 * (D + 9 + 6 + 6 * F + (1 if heading else 0) + H + 1 + R + 2 * F + O + D + D + 6) + F + 2 + 8
    return reads, writes


def case_for(config):
    import common
    return {"anymal_c_rough": "anymal_c_rough", "anymal_c_flat": "anymal_c_flat", "a1": "a1_rough", "a1_rough": "a1_rough",
            "go2_rough": "go2_rough", "go2": "go2_rough"}[config], common


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        try:      # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by its uuid
            self.index = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
        except Exception:
            pass

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "period_ms": 10}


# ----------------------------------------------------------------------------------------------
def cpu_reference_run(case, common, n_envs, steps, warmup, threads):
    """The reference's torch-CPU implementation (oracle port) of the same step on the host cores."""
    from oracle.legged_oracle import LeggedOracle
    from extended_legged_gym_b200 import synthetic
    torch.set_num_threads(threads)
    cfg, spec, st = common.make_case_state(case, n_envs, seed=0)
    hf = synthetic.make_height_field(seed=0)
    ora = LeggedOracle(cfg, spec, st, hf)
    for _ in range(warmup):
        ora.hot_step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ora.hot_step()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return n_envs / med, med, sum(ts)


def workload_config(config, n_envs, world, H=None, O=None, R_terms=None, n_rep=None, bytes_per_step=None):
    """`config` of the JSON line -- the SAME dict for both arms (what differs between them lives outside it)."""
    return {"workload": f"{config} post-physics step (torques+derive+heights+termination+rewards+obs+noise+history), {n_envs} envs per GPU",
            "num_envs_per_gpu": n_envs, "num_envs_total": n_envs * world,
            "sharding": f"envs x{world}, no data-path collective; episode statistics all-reduced once per K steps"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    case, common = case_for(args.config)
    threads = os.cpu_count() or 1
    steps = max(args.steps, 3)
    value, med, _ = cpu_reference_run(case, common, args.envs, min(steps, 200), max(args.warmup, 3), threads)
    sample = (f"{args.envs} envs x {min(steps, 200)} steps, median step, oracle port of the reference's torch-CPU code on {threads} threads" +
              ("" if world == 1 else f" (a bounded sample: one GPU's share of the {args.envs * world}-env workload; throughput, not the same total work)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": min(steps, 200),
            "warmup": max(args.warmup, 3), "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, args.envs, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        return None
    return None


def write_combined_host_block(like):
    """Pinned WRITE-COMBINED host memory (cudaHostAllocWriteCombined) holding a copy of `like`: the host only ever writes the
    inbound block, and the PCIe read of write-combined memory skips the snoop of the CPU caches (scripts/h2d_probe.py: 49 vs
    39 GB/s at these sizes).  Falls back to ordinary pinned memory."""
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so")
        p = ctypes.c_void_p()
        nbytes = like.numel() * like.element_size()
        if rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04)) != 0:
            raise RuntimeError("cudaHostAlloc failed")
        buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
        t = torch.frombuffer(buf, dtype=torch.uint8).view(like.dtype).view(like.shape)
        t.copy_(like)
        return t, "write-combined pinned"
    except Exception:
        return like.detach().cpu().pin_memory(), "pinned"


def build_replicas(case, common, n_envs, n_rep, dev, packed=False):
    from extended_legged_gym_b200 import _lib, synthetic
    from extended_legged_gym_b200.envs import LeggedRobot
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    hf = synthetic.make_height_field(seed=0).to(dev)
    envs = []
    for r in range(n_rep):
        cfg, spec, st = common.make_case_state(case, n_envs, seed=r)
        cfg.env.num_envs = n_envs
        sim = SyntheticSim(cfg, n_envs, dev, spec=spec, height_samples=hf, state=st, packed=packed)
        env = LeggedRobot(cfg, None, sim, dev, True)
        if packed:
            env.actions = sim.actions_in
        env.set_env_state(st)
        env.noise_u = None           # in-kernel Philox noise (0 algorithmic bytes)
        env._sync_native()
        envs.append(env)
    return envs


# ----------------------------------------------------------------------------------------------
def secondary_benchmarks(dev, world, rank, dist, quick=False, only_depth=False, comm=None, K=200):
    """The other kernels of the path at the BASELINE configs 4 and 5 (reported next to the headline, not in `value`):
    depth-camera ray casting (Mrays/s), the main -> rollout clone, the MPPI update (+ its collectives when N > 1)."""
    import numpy as np
    from types import SimpleNamespace
    from extended_legged_gym_b200 import synthetic
    from extended_legged_gym_b200.utils.depth_camera import DepthCameraWarp
    from extended_legged_gym_b200.utils.mppi import mppi_update
    from extended_legged_gym_b200.utils.ray_caster import raycast_mesh
    out = {}

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    # ---- config 4: 1024 cameras x 64 x 48 rays on the default terrain mesh (900 x 900 samples -> 1 616 402 triangles)
    t0 = time.perf_counter()
    hf = synthetic.make_height_field(seed=0)
    v, t = synthetic.heightfield_to_trimesh(hf)
    n_cam = 1024 // world if world > 1 else 1024
    g = torch.Generator().manual_seed(100 + rank)
    pos = torch.stack([torch.rand(n_cam, generator=g) * 36 + 2, torch.rand(n_cam, generator=g) * 36 + 2, torch.zeros(n_cam)], dim=1)
    ix = ((pos[:, 0] + 25.0) / 0.1).long().clamp(0, 898)
    iy = ((pos[:, 1] + 25.0) / 0.1).long().clamp(0, 898)
    pos[:, 2] = hf[ix, iy].float() * 0.005 + 0.55
    yaw = torch.rand(n_cam, generator=g) * 6.2832
    quat = torch.stack([torch.randn(n_cam, generator=g) * 0.03, torch.randn(n_cam, generator=g) * 0.03, torch.sin(yaw / 2), torch.cos(yaw / 2)], dim=1)
    quat = quat / quat.norm(dim=1, keepdim=True)
    ep = torch.full((n_cam,), 5, dtype=torch.int64, device=dev)
    depth = {}
    cam = None
    for far in (2.0, 10.0):
        cfg = SimpleNamespace(camera_type="Warp", original=(64, 48), resized=(64, 48), horizontal_fov=100, buffer_len=2, near_clip=0.0,
                              far_clip=far, dis_noise=0.0, position=[0.5, 0, 0.03], angle=[30, 30])
        if cam is None:
            cam = DepthCameraWarp(cfg, dev, n_cam, v, t)
            build_s = time.perf_counter() - t0
        else:
            mesh = cam.meshes
            cam = DepthCameraWarp(cfg, dev, n_cam, None, None)
            cam.meshes = mesh
        cam.update(0.02, pos.to(dev), quat.to(dev))
        cam.raw_depth = torch.zeros(n_cam, 48, 64, device=dev)
        sec = timed(lambda: cam.update_depth_buffer(None, ep), 3 if quick else 10)
        rays = n_cam * 64 * 48
        hit = float((cam.raw_depth > -far).float().mean())
        cam.raw_depth = None
        depth[f"far_clip_{int(far)}m"] = {"value": rays / sec / 1e6, "unit": "Mrays/s", "ms_per_frame_batch": sec * 1e3, "hit_fraction": hit}
    # API-level cast of the same rays (24 B in, 13 B out per ray; what raycast_mesh callers see)
    d = cam.ray_directions[:, :].reshape(-1, 3)
    o = cam.camera_pos.unsqueeze(1).expand(-1, 64 * 48, -1).reshape(-1, 3).contiguous()
    dirs = torch.nn.functional.normalize(torch.randn(o.shape[0], 3, device=dev) * torch.tensor([1.0, 1.0, 0.3], device=dev) -
                                         torch.tensor([0.0, 0.0, 0.5], device=dev), dim=1)
    sec = timed(lambda: raycast_mesh(o, dirs, 10.0, cam.meshes["terrain"]), 3 if quick else 10)
    out["depth_raycast"] = {"workload": f"{n_cam} cameras x 64x48 rays per GPU, default terrain mesh ({len(t)} triangles), fused depth kernel",
                            "cameras_per_gpu": n_cam, "rays_per_camera": 64 * 48, "triangles": int(len(t)), "bvh_build_s": build_s, **depth,
                            "raycast_mesh_incoherent_10m": {"value": o.shape[0] / sec / 1e6, "unit": "Mrays/s"}}
    del cam
    if only_depth:
        return out
    import ctypes as C
    from extended_legged_gym_b200 import _lib as L
    lib = L.load()
    peak, _ = peaks()
    progress("depth done; config 3")

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- config 3: a1 rough, 65 536 envs IN TOTAL, strong-scaled over the ranks (contiguous env blocks).  One timed iteration =
    # K steps through the public API's sync-free form (torques, command resampling, fused step, in-kernel reset -- whose episode
    # statistics accumulate on the device) + ONE ncclAllReduce of those statistics on the same stream; all of it one CUDA graph.
    import common
    from extended_legged_gym_b200.envs import LeggedRobot
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    from extended_legged_gym_b200.utils.distributed import ShardedEpisodeStats
    total3 = 65536
    n3 = total3 // world
    Ks = min(K, 50)
    hf_dev = hf.to(dev)
    rd3, wr3 = algorithmic_bytes_per_env(12, 4, 8, 1, 187, 235, 10, 4, False)
    bytes3 = (rd3 + wr3) * n3 + hf.numel() * 2
    n_rep3 = max(1, -(-2 * L2_BYTES // bytes3) + 1) if bytes3 < 2 * L2_BYTES else 3
    stats3 = ShardedEpisodeStats(dev, comm=comm)
    envs3 = []
    for r_ in range(n_rep3):
        cfg3, spec3, st3 = common.make_case_state("a1_rough", n3, seed=100 + r_ + 17 * rank)
        cfg3.env.num_envs = n3
        cfg3.domain_rand.push_robots = False
        e3 = LeggedRobot(cfg3, None, SyntheticSim(cfg3, n3, dev, spec=spec3, height_samples=hf_dev, state=st3), dev, True)
        e3.set_env_state(st3)
        e3.noise_u = None
        e3.episode_stats = stats3
        e3._obs_clip_for_step = 100.0
        envs3.append(e3)

    def step3(i):
        e = envs3[i % n_rep3]
        e.torques = e._compute_torques(e.actions).view(e.torques.shape)
        e.post_physics_step()                      # resample -> fused step -> in-kernel reset (sync-free)
    gs3 = torch.cuda.Stream(device=dev)
    g3 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gs3):
        for i in range(n_rep3):
            step3(i)
        gs3.synchronize()
        with torch.cuda.graph(g3, stream=gs3, capture_error_mode="thread_local"):
            for i in range(Ks):
                step3(i)
            if comm is not None:
                L.check(lib.elg_episode_stats_allreduce(stats3.buf.data_ptr(), stats3.buf.numel(), comm.handle, torch.cuda.current_stream(dev).cuda_stream))
        g3.replay()
        gs3.synchronize()
        ts = []
        for _ in range(5):
            if dist:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(gs3); g3.replay(); e1.record(gs3)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        # what the all-reduce delivers: the number of resets of ONE timed graph, summed over the ranks (the running totals are
        # zeroed first: replaying the in-place all-reduce without the reduce() / zero cycle of real use sums the sums)
        stats3.buf.zero_()
        gs3.synchronize()
        if dist:
            dist.barrier()
        g3.replay()
        gs3.synchronize()
    t3 = max_over_ranks(sorted(ts)[2])
    resets3 = float(stats3.buf[L.NUM_REWARD_TERMS])
    out["config3_strong"] = {"workload": f"a1 rough, {total3} envs in total = {n3} per GPU x {world}: {Ks} x (torques, resample, fused step, in-kernel reset) "
                                         f"+ 1 episode-statistics all-reduce per timed graph", "scaling": "strong", "num_envs_total": total3, "steps": Ks,
                             "value": total3 * Ks / t3, "unit": UNIT, "us_per_step": t3 / Ks * 1e6, "kernels_per_step": 4,
                             "per_gpu_algorithmic_gbs": bytes3 / (t3 / Ks) / 1e9, "frac_of_hbm_peak_per_gpu": bytes3 / (t3 / Ks) / 1e9 / peak,
                             "l2_policy": f"{n_rep3} state replicas x {bytes3 / 1e6:.1f} MB rotated per step", "timing": "CUDA events, max over ranks, median of 5 replays",
                             "collectives_in_timed_region": 1 if comm is not None else 0, "resets_in_one_timed_graph_all_ranks": resets3}
    del envs3, g3, stats3
    torch.cuda.empty_cache()
    # ---- config 1 (BASELINE configs[0], the reference's CPU-runnable case): anymal_c_flat, 4096 envs, no height scan -- the same
    # two kernels per step as the headline, every step on a different state replica
    n1 = 4096
    rd1, wr1 = algorithmic_bytes_per_env(12, 4, 17, 4, 0, 48, 9, 4, False)
    bytes1 = (rd1 + wr1) * n1
    n_rep1 = max(2, -(-2 * L2_BYTES // bytes1) + 1)
    envs1 = []
    for r_ in range(n_rep1):
        cfg1, spec1, st1 = common.make_case_state("anymal_c_flat", n1, seed=300 + r_)
        cfg1.env.num_envs = n1
        e1 = LeggedRobot(cfg1, None, SyntheticSim(cfg1, n1, dev, spec=spec1, height_samples=hf_dev, state=st1), dev, True)
        e1.set_env_state(st1)
        e1.noise_u = None
        e1._sync_native()
        envs1.append(e1)

    def step1(i):
        e = envs1[i % n_rep1]
        e.torques = e._compute_torques(e.actions).view(e.torques.shape)
        e._launch(L.PHASE_FUSED, 100.0, noise_step=i)
    gs1 = torch.cuda.Stream(device=dev)
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gs1):
        for i in range(n_rep1):
            step1(i)
        gs1.synchronize()
        with torch.cuda.graph(g1, stream=gs1):
            for i in range(200):
                step1(i)
        g1.replay()
        gs1.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(gs1)
        for _ in range(5):
            g1.replay()
        eb.record(gs1)
        gs1.synchronize()
    t1 = ea.elapsed_time(eb) * 1e-3 / 1000
    out["config1_flat"] = {"workload": f"anymal_c_flat post-physics step (torques + fused step, no height scan, 48 observations), {n1} envs per GPU",
                           "us_per_step": t1 * 1e6, "env_steps_per_s": n1 / t1, "algorithmic_bytes_per_step": bytes1,
                           "achieved_gbs": bytes1 / t1 / 1e9, "frac_of_hbm_peak": bytes1 / t1 / 1e9 / peak,
                           "l2_policy": f"{n_rep1} state replicas x {bytes1 / 1e6:.1f} MB rotated per step", "timing": "CUDA graph of 200 steps, 5 replays"}
    del envs1, g1
    torch.cuda.empty_cache()
    progress("config 3 done; clone / rollout / MPPI")

    # ---- config 5: 64 mains x 512 rollouts: state clone, then the cost-weighted update over a 20-step horizon
    import common
    from extended_legged_gym_b200.envs import RobotBatchRollout
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    mains, rollouts = 64, 512 // world if world > 1 else 512
    n = mains * (1 + rollouts)
    cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=3)
    cfg.env.num_envs, cfg.env.rollout_envs = mains, rollouts
    env = RobotBatchRollout(cfg, None, SyntheticSim(cfg, n, dev, spec=spec, height_samples=hf, state=st), dev, True)
    env.set_env_state(st)
    sec_api = timed(env._sync_main_to_rollout, 20 if quick else 200)
    # kernel time alone: 50 syncs replayed from one CUDA graph (the Python call costs more host time than the kernel runs)
    gs = torch.cuda.Stream(device=dev)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gs):
        env._sync_main_to_rollout()
        gs.synchronize()
        with torch.cuda.graph(gr, stream=gs):
            for _ in range(50):
                env._sync_main_to_rollout()
        gr.replay()
        gs.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(gs)
        for _ in range(4):
            gr.replay()
        e1.record(gs)
        gs.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / 200
    row = 4 * (13 + 24 + 12 * 3 + 6 + 9 + 8) + 4
    peak, _ = peaks()
    out["clone"] = {"workload": f"{mains} mains x {rollouts} rollouts per GPU, _sync_main_to_rollout", "us_per_sync_api": sec_api * 1e6,
                    "us_per_sync_kernel": sec * 1e6, "bytes_written": row * mains * rollouts, "achieved_gbs": row * mains * rollouts / sec / 1e9,
                    "frac_of_hbm_peak": row * mains * rollouts / sec / 1e9 / peak,
                    "kernel": "elg_clone_bulk_kernel (replicated shared-memory tiles, cp.async.bulk stores)",
                    "note": "kernel figure from a CUDA graph of 50 syncs; the 12.7 MB working set is L2 resident"}
    # the rollout-mode step over all 64 x (1 + rollouts) rows (post_physics_step_rollout, one launch per horizon step of the MPPI loop)
    env.noise_u = None
    gs = torch.cuda.Stream(device=dev)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gs):
        env.post_physics_step_rollout()
        gs.synchronize()
        with torch.cuda.graph(gr, stream=gs):
            for _ in range(20):
                env.post_physics_step_rollout()
        gr.replay()
        gs.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(gs)
        for _ in range(4):
            gr.replay()
        e1.record(gs)
        gs.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / 80
    out["rollout_step"] = {"workload": f"post_physics_step_rollout over {n} envs ({mains} mains x (1 + {rollouts}) rows), anymal_c_rough",
                           "us_per_step": sec * 1e6, "env_steps_per_s": n / sec,
                           "note": "lean kernel in rollout mode (measured heights are an input, no termination / episode sums); CUDA graph of 20 steps"}
    # one MPPI iteration of BASELINE config 5: 64 mains x 512 rollouts IN TOTAL, the rollouts of every main env split over the
    # ranks.  rollout_batch (sync, 20 x [actions, 4 x torques, rollout-mode step, restore], sync) is ONE CUDA graph; the
    # cost-weighted update is one extension call (costs -> ncclAllGather -> weights / partial sums -> ncclAllReduce -> means).
    from extended_legged_gym_b200.envs import RobotTrajGradSampling
    horizon, nodes = 20, 5
    del env
    cfg5, spec5, st5 = common.make_case_state("anymal_c_rough", n, seed=3)
    cfg5.env.num_envs, cfg5.env.rollout_envs = mains, rollouts
    env = RobotTrajGradSampling(cfg5, None, SyntheticSim(cfg5, n, dev, spec=spec5, height_samples=hf, state=st5), dev, True)
    env.set_env_state(st5)
    env.comm = comm
    env._cache_main_env_states()          # what step() leaves behind: the main rows the rollouts restore after every horizon step
    g5 = torch.Generator().manual_seed(7)
    us_full = torch.randn(mains, 512, horizon, 12, generator=g5) * 0.3          # the same on every rank; each takes its share
    smp_full = torch.randn(mains, 512, nodes, 12, generator=g5)
    lo5 = rank * rollouts
    all_us = us_full[:, lo5:lo5 + rollouts].reshape(mains * rollouts, horizon, 12).to(dev).contiguous()
    samples = smp_full[:, lo5:lo5 + rollouts].to(dev).contiguous()

    def mppi_iteration():
        rew = env.rollout_batch(all_us)
        return mppi_update(rew.view(mains, rollouts, horizon), samples, 0.05, comm=comm)
    for _ in range(2):
        mppi_iteration()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3 if quick else 7):
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); mppi_iteration(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    sec = max_over_ranks(sorted(ts)[len(ts) // 2])
    # correctness of the sharded update on the same data: every rank also runs the single-rank update on the FULL sample set
    rew_full = torch.randn(mains, 512, horizon, generator=g5).to(dev)
    smp_dev = smp_full.to(dev)
    from extended_legged_gym_b200.utils.mppi import mppi_update_native
    want = mppi_update_native(rew_full, smp_dev, 0.05, comm=None)          # the single-rank update, no collective
    got = mppi_update(rew_full[:, lo5:lo5 + rollouts].contiguous(), samples, 0.05, comm=comm)
    diff = float((got - want).abs().max())
    out["mppi_iteration"] = {"workload": f"{mains} mains x 512 rollouts in total ({rollouts} per main on each of {world} rank{'s' if world > 1 else ''}) x horizon "
                                         f"{horizon}: rollout_batch (one CUDA graph) + elg_mppi_update", "ms_per_iteration": sec * 1e3,
                             "rollout_env_steps_per_s": mains * 512 * horizon / sec, "timing": "CUDA events, max over ranks, median",
                             "collectives_per_iteration": 2 if comm is not None else 0,
                             "sharded_equals_single_rank": bool(torch.allclose(got, want, rtol=1e-4, atol=1e-5)), "max_abs_diff_vs_single_rank": diff}
    del env
    # ---- actuator-network torques (Anymal._compute_torques, the default torque path of the anymal_c configs): 4096 envs x 12 dofs
    from extended_legged_gym_b200.envs import Anymal
    n_act = 4096
    cfg, spec, st = common.make_case_state("anymal_c_rough", n_act, seed=4)
    cfg.env.num_envs = n_act
    cfg.control.actuator_net_weights = os.path.join(ROOT, "tests", "golden", "actuator_net.npz")
    aenv = Anymal(cfg, None, SyntheticSim(cfg, n_act, dev, spec=spec, height_samples=hf, state=st), dev, True)
    aenv.set_env_state(st)
    gs = torch.cuda.Stream(device=dev)

    def time_act():
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(gs):
            aenv._compute_torques(aenv.actions)
            gs.synchronize()
            with torch.cuda.graph(gr, stream=gs):
                for _ in range(50):
                    aenv._compute_torques(aenv.actions)
            gr.replay()
            gs.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(gs)
            for _ in range(4):
                gr.replay()
            e1.record(gs)
            gs.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / 200
    sec = time_act()                       # default: one thread per row, shared-memory weights staged before the grid-dependency wait
    lib.elg_set_actuator_tuning(2)
    sec_smem = time_act()                  # the same, weights staged after the wait
    lib.elg_set_actuator_tuning(5)
    sec_const = time_act()                 # weights as constant-bank operands (the default until the activations went branch-free)
    lib.elg_set_actuator_tuning(3)
    sec_split = time_act()                 # four warps per 32 rows (two hidden units per warp), weights from the constant bank
    lib.elg_set_actuator_tuning(0)
    act_bytes = n_act * 12 * (2 * 2 * 8 * 4 * 2 + 8 + 4 + 4)        # 4 state planes of 8 floats in + out, dof pos/vel, action, torque
    out["actuator_net"] = {"workload": f"{n_act} envs x 12 dofs, LSTMsea (2 -> 8 x 2 layers -> 1) per row, state in place", "us_per_call": sec * 1e6,
                           "bytes_per_call": act_bytes, "achieved_gbs": act_bytes / sec / 1e9, "frac_of_hbm_peak": act_bytes / sec / 1e9 / peak,
                           "us_per_call_weights_staged_after_the_wait": sec_smem * 1e6,
                           "us_per_call_constant_bank_weights": sec_const * 1e6,
                           "us_per_call_unit_split_form": sec_split * 1e6,
                           "note": "CUDA graph of 50 calls on one state (13.6 MB, L2 resident)"}
    del aenv
    # caller-side fusion: EmpiricalNormalization (training) + write into the rollout-storage slot + reward / done columns
    from extended_legged_gym_b200.utils.normalizer import EmpiricalNormalization
    n_obs_rows, n_obs = 4096, 235
    norm = EmpiricalNormalization(shape=[n_obs], until=int(1e8)).to(dev)
    norm.train()
    xo = torch.randn(n_obs_rows, n_obs, device=dev)
    slot = torch.empty(24, n_obs_rows, n_obs, device=dev)
    rw, dn = torch.randn(n_obs_rows, device=dev), torch.zeros(n_obs_rows, device=dev, dtype=torch.bool)
    rws, dns = torch.empty(24, n_obs_rows, 1, device=dev), torch.empty(24, n_obs_rows, 1, device=dev, dtype=torch.uint8)
    def time_norm():
        with torch.cuda.stream(gs):
            norm.forward_into(xo, slot[0], rw, rws[0], dn, dns[0])
            gs.synchronize()
            gn = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gn, stream=gs):
                for i in range(48):
                    norm.forward_into(xo, slot[i % 24], rw, rws[i % 24], dn, dns[i % 24])
            gn.replay()
            gs.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(gs)
            for _ in range(4):
                gn.replay()
            e1.record(gs)
            gs.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / 192
    sec = time_norm()                      # default: ONE launch (a CTA owns four columns and all their rows, which stay in registers)
    lib.elg_set_normalizer_tuning(1)
    sec2 = time_norm()                     # the statistics + apply pair that larger batches take
    lib.elg_set_normalizer_tuning(0)
    nb = n_obs_rows * n_obs * 4 * 2
    out["obs_normalize_store"] = {"workload": f"{n_obs_rows} x {n_obs} observations: running mean / var update + normalise + write to the storage slot "
                                              "(+ reward / done columns), 1 launch", "us_per_call": sec * 1e6, "bytes_per_call": nb,
                                  "achieved_gbs": nb / sec / 1e9, "frac_of_hbm_peak": nb / sec / 1e9 / peak,
                                  "us_per_call_two_launch_form": sec2 * 1e6,
                                  "note": "CUDA graph of 48 calls cycling over 24 storage slots (92 MB of destinations)"}
    # navigation commands and kinematic state integration over the 64 x (1 + 512) rows of config 5 (one launch each)
    from extended_legged_gym_b200.envs import KinematicStateIntegration
    n_all = mains * (1 + rollouts)
    root = torch.randn(n_all, 13, device=dev)
    root[:, 3:7] /= root[:, 3:7].norm(dim=1, keepdim=True)
    goals = torch.randn(mains, 3, device=dev)
    cmds, prevc = torch.zeros(n_all, 4, device=dev), torch.zeros(n_all, 3, device=dev)
    reached, dist_g = torch.zeros(n_all, dtype=torch.bool, device=dev), torch.zeros(n_all, device=dev)
    npar = L.ElgNavParams(1, 1, 4, 1, 1.0, 2.0, 1.0, 1.0, 0.1, 0.9, 0.5)

    class Plan(KinematicStateIntegration):
        pass
    pl = Plan()
    pl.total_num_envs, pl.num_dof, pl.device = n_all, 12, dev
    pl.root_states, pl.dof_state = root.clone(), torch.zeros(n_all * 12, 2, device=dev)
    pl.base_lin_vel, pl.base_ang_vel = torch.zeros(n_all, 3, device=dev), torch.zeros(n_all, 3, device=dev)
    pl._init_planning_settings(SimpleNamespace(max_base_lin_vel=3.0, max_base_ang_vel=2.0, max_joint_vel=10.0, integration_method="euler",
                                               max_integration_step=0.01, enforce_joint_limits=False))
    pl._init_planning_buffers()
    pl._sync_sim_to_integration()
    roll_ids = torch.arange(n_all, device=dev).view(mains, 1 + rollouts)[:, 1:].reshape(-1).contiguous()
    sv = torch.randn(roll_ids.numel(), 18, device=dev)

    def graph_time(fn, reps=50):
        with torch.cuda.stream(gs):
            fn()
            gs.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=gs):
                for _ in range(reps):
                    fn()
            g.replay()
            gs.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(gs)
            for _ in range(4):
                g.replay()
            b.record(gs)
            gs.synchronize()
        return a.elapsed_time(b) * 1e-3 / (4 * reps)

    sec = graph_time(lambda: L.check(lib.elg_nav_commands(mains, rollouts, C.byref(npar), root.data_ptr(), goals.data_ptr(), cmds.data_ptr(),
                                                          prevc.data_ptr(), reached.data_ptr(), dist_g.data_ptr(),
                                                          torch.cuda.current_stream().cuda_stream)))
    out["nav_commands"] = {"workload": f"{n_all} envs ({mains} mains x (1 + {rollouts})): goal-directed commands + goal flags, one launch", "us_per_call": sec * 1e6}
    sec = graph_time(lambda: pl._integrate_state_velocities(sv, 0.02, roll_ids))
    pb = roll_ids.numel() * 4 * (18 + 2 * (7 + 12 + 6 + 12) + 13 + 24 + 6)
    out["plan_integrate"] = {"workload": f"{roll_ids.numel()} rollout envs, 12 dofs: 2 Euler sub-steps + write-through to root / dof state, one launch",
                             "us_per_call": sec * 1e6, "bytes_per_call": pb, "achieved_gbs": pb / sec / 1e9}
    K, D, T = 5, 12, 20
    r = torch.randn(mains, rollouts, T, device=dev)
    u = torch.randn(mains, rollouts, K, D, device=dev)
    sec = timed(lambda: mppi_update(r, u, 0.05, comm=comm), 20 if quick else 100)
    out["mppi_update"] = {"workload": f"{mains} mains x {rollouts * world} samples (x{world} ranks) x horizon {T}, {K} nodes x {D} dof",
                          "us_per_update": max_over_ranks(sec) * 1e6,
                          "collectives": "ncclAllGather(costs) + ncclAllReduce(partials), issued by the extension on the compute stream" if world > 1 else "none (1 rank)"}
    return out


def progress(msg):
    """milestones on stderr (stdout carries only the result line): a multi-rank run that stalls shows where"""
    print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def graph_of(fn, n, stream, warm=3):
    """Capture n calls of fn(i) on `stream` into one CUDA graph (after `warm` eager calls)."""
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for i in range(warm):
            fn(i)
        stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            for i in range(n):
                fn(i)
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="anymal_c_rough")
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config 3 / depth / clone / MPPI measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    from extended_legged_gym_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    numa = pin_to_gpu_numa_node(local_rank)
    dist, comm = None, None
    # stdout carries exactly one JSON line: anything a library (NCCL's version / debug lines) or a host class prints on the way
    # goes to stderr -- fd 1 points at stderr until the result line is due
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        from extended_legged_gym_b200.utils.distributed import ElgComm
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(dev))
        comm = ElgComm(dev)               # the extension's own communicator: collectives on the compute stream, graph-capturable
    from extended_legged_gym_b200.utils.distributed import ShardedEpisodeStats
    W = max(args.warmup, 3)
    K = args.steps
    case, common = case_for(args.config)
    lib = _lib.load()
    cur = lambda: torch.cuda.current_stream(dev).cuda_stream

    n_envs = args.envs
    probe = build_replicas(case, common, n_envs, 1, dev, packed=True)[0]      # replica 0 doubles as the e2e env (packed state block)
    D, F = probe.num_dof, len(probe.feet_indices)
    P, T = len(probe.penalised_contact_indices), len(probe.termination_contact_indices)
    H, O, C_ = probe.num_height_points, probe.num_obs, probe.cfg.commands.num_commands
    R_terms = len(probe.reward_names)
    rd, wr = algorithmic_bytes_per_env(D, F, P, T, H, O, R_terms, C_, probe.cfg.commands.heading_command)
    hf_bytes = probe.height_samples.numel() * 2 if probe.height_samples is not None else 0
    bytes_per_step = (rd + wr) * n_envs + hf_bytes
    n_rep = max(2, -(-2 * L2_BYTES // bytes_per_step) + 1)
    envs = [probe] + build_replicas(case, common, n_envs, n_rep, dev)[1:]
    stats = ShardedEpisodeStats(dev, comm=comm)
    stream = torch.cuda.Stream(device=dev)

    def enqueue(env, step, with_torques=True):
        p = env._params
        p.noise_mode, p.noise_offset, p.clip_observations = _lib.NOISE_PHILOX, step, 100.0
        if with_torques:
            rc = lib.elg_compute_torques(C.byref(env._dims), C.byref(p), env.actions.data_ptr(), env.dof_state.data_ptr(),
                                         env.last_dof_vel.data_ptr(), env.p_gains.data_ptr(), env.d_gains.data_ptr(),
                                         env.torque_limits.data_ptr(), env.default_dof_pos.data_ptr(), env.torques.data_ptr(), None, 0, cur())
            _lib.check(rc)
        _lib.check(lib.elg_post_physics_step(C.byref(env._dims), C.byref(p), C.byref(env._bufs), _lib.PHASE_FUSED, cur()))

    def capture(n_steps, with_torques, reps, with_allreduce):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            for i in range(3):
                enqueue(envs[i % reps], i, with_torques)
            stream.synchronize()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                for i in range(n_steps):
                    enqueue(envs[i % reps], i, with_torques)
                if with_allreduce and comm is not None:
                    # the one genuine reduction of the path, once per K steps, INSIDE the timed region: the running episode
                    # statistics (sums per reward term, reset count) of all ranks, NCCL on the capture stream
                    _lib.check(lib.elg_episode_stats_allreduce(stats._send.data_ptr(), stats._send.numel(), comm.handle, cur()))
        return g

    def timed_replay(g, reps=1):
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(reps):
                g.replay()
            e1.record(stream)
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
        return e0.elapsed_time(e1) * 1e-3 / reps

    progress("replicas built; capturing graphs")
    g_warm = capture(W, True, n_rep, True)
    g_full = capture(K, True, n_rep, True)
    g_step = capture(K, False, n_rep, False)
    g_step_warm = capture(K, False, 1, False)      # steady state: one replica, inputs and outputs stay L2-resident
    timed_replay(g_warm)
    timed_replay(g_full)          # first replay uploads the graph; not a measurement
    timed_replay(g_step)
    timed_replay(g_step_warm)
    with ClockSampler(local_rank) as clk:
        # the sampler (10 ms period) needs a busy region of >= 0.5 s to see clocks UNDER LOAD: the timed graph replayed back to back
        # (the replay count must not depend on a rank's own timing: the graph holds a collective)
        timed_replay(g_full, reps=max(1, min(20000, int(0.6 / (K * 12e-6)))))
        t_fulls = sorted(timed_replay(g_full) for _ in range(7))
        t_steps = sorted(timed_replay(g_step) for _ in range(7))
        t_warms = sorted(timed_replay(g_step_warm) for _ in range(7))
    t_full, t_step_only, t_step_warm = t_fulls[len(t_fulls) // 2], t_steps[len(t_steps) // 2], t_warms[len(t_warms) // 2]
    clocks = clk.summary()
    progress("headline timed")
    if dist:
        tt = torch.tensor([t_full, t_step_only, t_step_warm], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_full, t_step_only, t_step_warm = tt.tolist()
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        sm = [c["sm_mhz"] for c in allc if c["sm_mhz"]]
        clocks = {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": max((c["sm_max_mhz"] or 0) for c in allc) or None,
                  "reasons": sorted(set(r for c in allc for r in c["reasons"])), "samples": sum(c["samples"] for c in allc),
                  "period_ms": 10, "note": "min over ranks of the per-rank median SM clock under load"}
    total_envs = n_envs * world
    value = total_envs * K / t_full
    ms_per_step = t_full / K * 1e3
    peak, peak_src = peaks()
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "step_traffic.json")     # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
    if os.path.exists(tpath) and args.config == "anymal_c_rough" and n_envs == 4096:
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    t_kernel = t_step_only / K
    achieved = bytes_per_step / t_kernel / 1e9
    achieved_warm = bytes_per_step / (t_step_warm / K) / 1e9

    # ---- launch floor on this box, same launch shape (csrc/elg_probe.cu): what one launch / the bytes alone cost
    floor = None
    if hasattr(lib, "elg_probe_empty"):
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        epc = -(-n_envs // sms)
        epc += (-epc) % 4
        b_in, b_out = (rd * epc + 15) // 16 * 16, (wr * epc + 15) // 16 * 16
        if max(b_in, b_out) <= 200 * 1024:
            n_fl = max(2, -(-2 * L2_BYTES // (sms * (b_in + b_out))) + 1)
            src = [torch.zeros(sms * b_in, dtype=torch.uint8, device=dev) for _ in range(n_fl)]
            dst = [torch.zeros(sms * b_out, dtype=torch.uint8, device=dev) for _ in range(n_fl)]
            g_e = graph_of(lambda i: _lib.check(lib.elg_probe_empty(sms, 1024, 90 * 1024, 1, cur())), 200, stream)
            g_c = graph_of(lambda i: _lib.check(lib.elg_probe_roundtrip(src[i % n_fl].data_ptr(), dst[i % n_fl].data_ptr(), b_in, b_out, sms, 1, cur())), 200, stream)
            g_w = graph_of(lambda i: _lib.check(lib.elg_probe_roundtrip(src[0].data_ptr(), dst[0].data_ptr(), b_in, b_out, sms, 1, cur())), 200, stream)
            for g in (g_e, g_c, g_w):
                timed_replay(g)
            med = lambda g: sorted(timed_replay(g) for _ in range(5))[2] / 200 * 1e6
            floor = {"empty_pdl_kernel_us": med(g_e), "bulk_roundtrip_cold_us": med(g_c), "bulk_roundtrip_warm_us": med(g_w),
                     "bytes_per_launch": sms * (b_in + b_out),
                     "what": f"{sms} CTAs x 1024 threads, 90 KB smem, PDL, CUDA graph: an empty kernel; one cp.async.bulk of the step's input bytes in and "
                             "one of its output bytes out per CTA with no arithmetic (cold: rotating buffers > 2x L2; warm: one buffer)"}
            del src, dst, g_e, g_c, g_w

    progress("floor probes done; e2e")
    # ---- e2e: public Python API, PhysX state in pinned host memory, ONE packed H2D + ONE packed D2H every step
    env = envs[0]
    sim = env.sim
    host_in, host_in_kind = write_combined_host_block(sim.state_block.detach().cpu())   # root / dof / contact / rigid-body state + actions, one block
    # outputs the caller reads back, re-pointed at ONE device block: obs [N, O] | rew [N] | reset flags [N] (bytes, padded); two such
    # blocks alternate, so that the kernels of step i + 1 need not wait for the copy-out of step i
    n_out_words = n_envs * O + n_envs + (n_envs + 3) // 4
    out_blocks = [torch.zeros(n_out_words, dtype=torch.float, device=dev) for _ in range(2)]

    def point_outputs_at(blk):
        env.obs_buf = blk[:n_envs * O].view(n_envs, O)
        env.rew_buf = blk[n_envs * O:n_envs * O + n_envs]
        env._reset_bool = blk[n_envs * O + n_envs:].view(torch.uint8)[:n_envs].view(torch.bool)
    point_outputs_at(out_blocks[0])
    # G consecutive steps form one CUDA graph; the copy-out of step i (second stream) overlaps the copy-in of step i + 1 (PCIe is
    # full duplex); the host launches the graph, waits and reads every step's result.
    G = max(2, min(40, args.e2e_steps))
    G -= G % 2                             # (the two staging / output blocks alternate: an even number of steps per graph)
    host_out = [torch.empty(n_out_words).pin_memory() for _ in range(G)]
    h2d = host_in.numel() * 4
    d2h = n_out_words * 4
    env.cfg.domain_rand.push_robots = False
    env._obs_clip_for_step = 100.0
    s_cap, s_out, s_in = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    checksum = [0.0]
    # The inbound block lands in one of two device staging blocks on its own stream, so that the copy-in of step i + 1 runs UNDER the
    # kernels of step i (and next to the copy-out of step i: PCIe is full duplex); the compute stream moves the staged block into the
    # simulator's state block with one device-to-device copy (5.26 MB, a few us) right before the step's kernels.
    # (measured on B200, profiles/README.md r2: one copy engine moves this block at 52.7 GB/s alone and at 42.7 GB/s while the
    #  copy-out runs; an SM-issued copy from mapped host memory -- elg_stage_block, scripts/stage_probe.py -- reaches 49.9 GB/s, no
    #  better; 2 / 4 concurrent chunked copies on side streams are slower; write-combined host memory changes nothing at this size)
    stage = [torch.empty_like(sim.state_block) for _ in range(2)]
    stage_free = [None, None]      # event: the device-to-device copy out of this staging block has run
    out_free = [None, None]        # event: the copy-out of this output block has run

    def one_step(slot):
        cur = torch.cuda.current_stream()
        b = slot & 1
        landed = torch.cuda.Event()
        with torch.cuda.stream(s_in):
            if stage_free[b] is not None:
                s_in.wait_event(stage_free[b])
            stage[b].copy_(host_in, non_blocking=True)
            landed.record(s_in)
        cur.wait_event(landed)
        sim.state_block.copy_(stage[b], non_blocking=True)
        stage_free[b] = torch.cuda.Event()
        stage_free[b].record(cur)
        point_outputs_at(out_blocks[b])
        if out_free[b] is not None:
            cur.wait_event(out_free[b])       # the results of two steps ago have left this output block
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        done = torch.cuda.Event()
        done.record(cur)
        s_out.wait_event(done)
        out_free[b] = torch.cuda.Event()
        with torch.cuda.stream(s_out):
            host_out[slot].copy_(out_blocks[b], non_blocking=True)
            out_free[b].record(s_out)

    def join_side_streams():
        cur = torch.cuda.current_stream()
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)

    with torch.cuda.stream(s_cap):
        s_in.wait_stream(s_cap)
        s_out.wait_stream(s_cap)
        one_step(0)                           # eager, both output blocks once: everything lazily created exists before the capture
        one_step(1)
        join_side_streams()
    s_cap.synchronize()
    stage_free, out_free = [None, None], [None, None]
    e2e_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(e2e_graph, stream=s_cap):
        s_in.wait_stream(s_cap)                # fork the copy streams into the capture
        s_out.wait_stream(s_cap)
        for slot in range(G):
            one_step(slot)
        join_side_streams()

    def e2e_block():
        e2e_graph.replay()
        s_cap.synchronize()
        for slot in range(G):                 # the host reads every step's result (first reward of the step)
            checksum[0] += float(host_out[slot][n_envs * O])

    n_blocks = max(1, args.e2e_steps // G)
    e2e_steps = n_blocks * G
    for _ in range(max(1, W // G)):
        e2e_block()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_cap):
        e0.record(s_cap)
        for _ in range(n_blocks):
            e2e_block()
        e1.record(s_cap)
    torch.cuda.synchronize()
    t_e2e = e0.elapsed_time(e1) * 1e-3
    if dist:
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = tt.item()
    e2e_value = total_envs * e2e_steps / t_e2e

    cfg_line = workload_config(args.config, n_envs, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg_line,
        "workload_detail": {"height_points": H, "num_obs": O, "reward_terms": R_terms,
                            "l2_policy": f"inputs larger than L2: {n_rep} state replicas x {bytes_per_step / 1e6:.1f} MB rotated per step",
                            "noise": "in-kernel Philox4x32-10",
                            "launch": "CUDA graph of K steps, 2 kernels per step (programmatic dependent launch)" +
                                      (", + 1 ncclAllReduce of the episode statistics (extension, capture stream) as the last node" if comm else ""),
                            "numa": numa},
        "gpu_launches": 2 * K,
        "collectives_in_timed_region": 1 if comm is not None else 0,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": t_e2e / e2e_steps * 1e3,
                "api": "LeggedRobot._compute_torques + post_physics_step (in-kernel reset path included), simulator state in pinned host memory: "
                       f"every step ONE copy of the packed state + actions block ({host_in_kind} host memory) in and ONE copy of the packed obs / rew / reset block out; "
                       f"{G} steps per CUDA graph; the copy-in of step i + 1 (own stream, into one of two device staging blocks; one device-to-device "
                       "copy hands it to the state block) runs under the kernels and the copy-out (third stream, two alternating output blocks) "
                       "of step i; the host waits for the graph and reads every step's result"},
        "roofline": {"bound": "hbm", "kernel": "elg_step_fast_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_step, "bytes_per_env": rd + wr, "us_per_launch": t_kernel * 1e6,
                     "regime": "cold L2 (every launch on a different state replica)",
                     "steady_state": {"us_per_launch": t_step_warm / K * 1e6, "achieved": achieved_warm, "frac": achieved_warm / peak,
                                      "regime": "one state replica: inputs and outputs L2-resident, as when the simulator has just written the state"},
                     "floor": floor,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "clocks": clocks,
    }
    if not args.no_secondary:
        del envs, probe, env, e2e_graph, g_full, g_step, g_step_warm, g_warm
        torch.cuda.empty_cache()
        progress("e2e done; secondary")
        sec = secondary_benchmarks(dev, world, rank, dist, quick=args.steps < 200, comm=comm, K=K)
        progress("secondary done")
        if dist:
            sec["note"] = "per-GPU figures measured on rank 0 unless the entry says max over ranks (each rank runs its own share)"
        line["secondary"] = sec
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, med, tot = cpu_reference_run(case, common, n_envs, 20, 3, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"full workload: {n_envs} envs x 20 steps (median step {med * 1e3:.1f} ms) of oracle/legged_oracle.py hot_step"}
    if dist:
        dist.barrier()
        if comm is not None:
            comm.close()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
