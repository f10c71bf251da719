"""``EmpiricalNormalization`` -- running mean / variance normalisation of the observations, host side.

Mirrors rsl_rl/modules/normalizer.py:14-79 of the reference (same constructor, buffers ``_mean / _var / _std / count``, so
``state_dict()`` / ``load_state_dict()`` interchange with the reference's checkpoints, ``mean`` / ``std`` properties,
``forward`` / ``update`` / ``inverse``).  ``forward`` in training mode is one call of ``elg_normalize_observations`` (two
launches) instead of ~15 ATen ops and -- with ``until`` set, as the runner does (on_policy_runner.py:282) -- a device->host read
of ``count`` per call; ``forward_into`` additionally lets the caller name the destination (the rollout-storage slot) and pass
the reward / done columns to be stored with it (rollout_storage.py:95-100), so the step's outputs reach the policy's buffers in
one pass.
"""
import torch
from torch import nn

from .. import _lib


class EmpiricalNormalization(nn.Module):
    def __init__(self, shape, eps=1e-2, until=None):
        super().__init__()
        self.eps = eps
        self.until = until
        self.register_buffer("_mean", torch.zeros(shape).unsqueeze(0))
        self.register_buffer("_var", torch.ones(shape).unsqueeze(0))
        self.register_buffer("_std", torch.ones(shape).unsqueeze(0))
        self.register_buffer("count", torch.tensor(0, dtype=torch.long))
        self._scratch = None

    @property
    def mean(self):
        return self._mean.squeeze(0).clone()

    @property
    def std(self):
        return self._std.squeeze(0).clone()

    def _launch(self, x, out, training, rew=None, rew_out=None, dones=None, dones_out=None):
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2):
            raise ValueError("EmpiricalNormalization needs a contiguous float32 CUDA tensor [N, num_obs]")
        n, o = x.shape
        if o != self._mean.shape[1]:
            raise ValueError(f"Expected {self._mean.shape[1]} observation columns, got {o}")
        for t in (self._mean, self._var, self._std, self.count):
            if t.device != x.device:
                raise ValueError("normalizer state and input are on different devices (call .to(device) first)")
        for t, dt in ((out, torch.float32), (rew, torch.float32), (rew_out, torch.float32)):
            if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == dt):
                raise ValueError("destinations / rewards must be contiguous float32 CUDA tensors")
        for t in (dones, dones_out):
            if t is not None and not (t.is_cuda and t.is_contiguous() and t.element_size() == 1):
                raise ValueError("dones must be contiguous bool / uint8 CUDA tensors")
        lib = _lib.load()
        scratch = None
        if training:
            need = lib.elg_normalizer_scratch_bytes(n, o)
            if self._scratch is None or self._scratch.numel() < need or self._scratch.device != x.device:
                self._scratch = torch.zeros(need, dtype=torch.uint8, device=x.device)       # tickets start at zero
            scratch = self._scratch
        until = -1 if self.until is None else int(self.until)
        _lib.check(lib.elg_normalize_observations(n, o, x.data_ptr(), self._mean.data_ptr(), self._var.data_ptr(), self._std.data_ptr(),
                                                  self.count.data_ptr(), float(self.eps), until, int(training), _lib.ptr(out), _lib.ptr(scratch),
                                                  _lib.ptr(rew), _lib.ptr(rew_out), _lib.ptr(dones), _lib.ptr(dones_out),
                                                  torch.cuda.current_stream(x.device).cuda_stream), "elg_normalize_observations")

    def forward(self, x):
        """normalizer.py:43-56"""
        out = torch.empty_like(x)
        self._launch(x, out, self.training)
        return out

    def forward_into(self, x, out, rewards=None, rewards_out=None, dones=None, dones_out=None):
        """``forward`` with the result written to ``out`` (may be ``x`` itself, or e.g. ``storage.observations[step]``);
        ``rewards -> rewards_out`` and ``dones -> dones_out`` ([N] or [N, 1]) are copied by the same launch."""
        self._launch(x, out, self.training, rewards, rewards_out, dones, dones_out)
        return out

    @torch.jit.unused
    def update(self, x):
        """normalizer.py:58-75 -- learn without producing the output"""
        self._launch(x, None, True)

    @torch.jit.unused
    def inverse(self, y):
        return y * (self._std + self.eps) + self._mean
