"""Sharding of independent envs across the GPUs of one box and the two genuine reductions of the path.

The reference env is single-GPU (no collective anywhere on the env path); envs are independent, so
they shard with no data-path collective.  What does reduce: the episode statistics of ``reset_idx``
(envs/base/legged_robot.py:200-213: means of ``episode_sums`` over the envs that reset) and the MPPI update
(``utils/mppi.py``).  One process per GPU.  On the GPU both reductions are NCCL calls issued by the extension on the
compute stream (``csrc/elg_nccl.cu``: graph-capturable, no host synchronisation) through an ``ElgComm``;
``torch.distributed`` only carries the 128-byte ncclUniqueId at start-up (and is the whole transport of the gloo CPU tests).
"""
import ctypes as C

import torch
import torch.distributed as dist

from .. import _lib


def shard_range(total, rank, world):
    """Contiguous block of ``total`` units owned by ``rank`` (first ``total % world`` ranks get one more)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ElgComm:
    """One ncclComm_t per rank, owned by the extension (include/elg_b200.h ``ElgComm``).  Collective over the default
    ``torch.distributed`` group (or ``group``): rank 0 draws the ncclUniqueId, a broadcast hands it to the others."""

    def __init__(self, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.ElgError("ElgComm needs an initialised torch.distributed process group (it carries the ncclUniqueId)")
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = self._lib = _lib.load()
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(lib.elg_comm_unique_id(uid), "elg_comm_unique_id")
        on_cuda = dist.get_backend(group) == "nccl"
        t = torch.tensor(list(uid), dtype=torch.uint8, device=self.device if on_cuda else "cpu")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = (C.c_uint8 * 128)(*t.cpu().tolist())
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.elg_comm_init(uid, self.rank, self.world, C.byref(handle)), "elg_comm_init")
            self.handle = handle
            # NCCL connects its peers lazily, at the first collective of each kind (allocations, IPC handles, proxy threads):
            # that must not happen inside a CUDA-graph capture, so both collectives the path uses run once here, eagerly
            warm = torch.zeros(8 * self.world, dtype=torch.float64, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.elg_episode_stats_allreduce(warm.data_ptr(), 8, handle, stream), "elg_episode_stats_allreduce (warm-up)")
            if hasattr(lib, "elg_comm_warmup"):
                _lib.check(lib.elg_comm_warmup(handle, warm.data_ptr(), 8, stream), "elg_comm_warmup")
            torch.cuda.synchronize(self.device)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.elg_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedEpisodeStats:
    """Running (sum, count) of the per-term episode returns of the envs that reset, on the device with no host sync:
    ``buf`` = float64 [NUM_REWARD_TERMS + 1] in registry-id order.  The fused reset kernel adds to it directly
    (``ElgResetBuffers.stats_accum``); the host-driven reset path calls ``accumulate``.  ``reduce()`` -- called by every
    rank at the same point, e.g. once per K steps -- is ONE all-reduce of that vector (+ two words for the terrain level)
    and returns the means ``extras['episode']`` would hold on a single GPU owning all envs, as device tensors."""

    def __init__(self, device, comm=None, group=None, use_torch_distributed=False):
        """comm: an ``ElgComm`` (NCCL on the compute stream); without one the totals stay local unless
        ``use_torch_distributed`` routes the all-reduce through ``torch.distributed`` (``group``; the gloo CPU tests)."""
        self.device, self.comm, self.group, self.use_dist = torch.device(device), comm, group, bool(use_torch_distributed)
        self.buf = torch.zeros(_lib.NUM_REWARD_TERMS + 1, dtype=torch.float64, device=device)
        self._send = torch.zeros(_lib.NUM_REWARD_TERMS + 3, dtype=torch.float64, device=device)

    def accumulate(self, episode_sums_all, env_ids):
        """episode_sums_all: the [NUM_REWARD_TERMS, N] SoA tensor of the env; env_ids: the rows that reset."""
        if len(env_ids) == 0:
            return
        nt = _lib.NUM_REWARD_TERMS
        self.buf[:nt] += episode_sums_all[:, env_ids].sum(dim=1, dtype=torch.float64)
        self.buf[nt] += len(env_ids)

    def reduce(self, max_episode_length_s, terrain_levels=None, names=None):
        """All-reduce and clear.  Returns {"rew_<name>": 0-d tensor, ..., "num_resets": 0-d tensor[, "terrain_level"]};
        terms of an iteration without any reset come back as NaN (0 / 0), like an empty ``torch.mean``."""
        nt = _lib.NUM_REWARD_TERMS
        v = self._send
        v[:nt + 1] = self.buf
        if terrain_levels is not None:
            v[nt + 1] = terrain_levels.sum(dtype=torch.float64)
            v[nt + 2] = terrain_levels.numel()
        else:
            v[nt + 1:] = 0
        if self.comm is not None and self.comm.world > 1:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.load().elg_episode_stats_allreduce(v.data_ptr(), v.numel(), self.comm.handle, stream), "elg_episode_stats_allreduce")
        elif self.use_dist and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)         # gloo (CPU tests) / no ElgComm
        means = (v[:nt] / v[nt] / max_episode_length_s).to(torch.float)
        out = {"rew_" + name: means[_lib.TERM_ID[name]] for name in (names if names is not None else _lib.REWARD_TERMS)}
        out["num_resets"] = v[nt].clone()
        if terrain_levels is not None:
            out["terrain_level"] = (v[nt + 1] / v[nt + 2]).to(torch.float)
        self.buf.zero_()
        return out
