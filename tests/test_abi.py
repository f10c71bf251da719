"""CPU-only checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/elg_b200.h declares, its struct sizes agree with the Python mirrors, and
argument validation fails loudly (no compute is launched here -- there is no GPU)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from extended_legged_gym_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "elg_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(elg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    names = declared_functions()
    assert "elg_post_physics_step" in names and "elg_compute_torques" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in elg_b200.h but not exported"


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_struct_mirrors_and_registry_order():
    lib = _lib.load()
    assert lib.elg_abi_version() == 4
    assert lib.elg_sizeof_dims() == C.sizeof(_lib.ElgDims)
    assert lib.elg_sizeof_step_params() == C.sizeof(_lib.ElgStepParams)
    assert lib.elg_sizeof_step_buffers() == C.sizeof(_lib.ElgStepBuffers)
    # ids are the alphabetical order class_to_dict yields for the reference registry
    assert _lib.REWARD_TERMS == sorted(_lib.REWARD_TERMS)
    hdr = open(HEADER).read()
    enum = re.search(r"typedef enum ElgRewardTerm \{(.*?)\} ElgRewardTerm;", hdr, re.S).group(1)
    ids = [m.lower()[len("elg_rew_"):] for m in re.findall(r"ELG_REW_[A-Z0-9_]+", enum)]
    assert ids == _lib.REWARD_TERMS
    assert lib.elg_reward_term_name(len(ids)) is None


def test_argument_validation_fails_loudly():
    lib = _lib.load()
    d, p, b = _lib.ElgDims(), _lib.ElgStepParams(), _lib.ElgStepBuffers()
    assert lib.elg_post_physics_step(None, C.byref(p), C.byref(b), 31, None) == -4
    d.num_envs, d.num_dof = 4, 0
    assert lib.elg_post_physics_step(C.byref(d), C.byref(p), C.byref(b), 31, None) == -1
    assert b"num_dof" in lib.elg_last_error()
    d.num_dof, d.num_obs, d.num_commands = 12, 48, 4
    assert lib.elg_post_physics_step(C.byref(d), C.byref(p), C.byref(b), 0, None) == -1
    assert lib.elg_post_physics_step(C.byref(d), C.byref(p), C.byref(b), 31, None) == -4     # NULL state pointers
    p.control_type = 7
    rc = lib.elg_compute_torques(C.byref(d), C.byref(p), 1, 1, 1, 1, 1, 1, 1, 1, None, 0, None)
    assert rc == -1
    with pytest.raises(NameError):
        _lib.check(rc)


@pytest.mark.gpu
def test_stage_block_probe_copies_both_ways_in_both_modes():
    """elg_stage_block (the SM-issued host <-> device copy measured against the copy engine): exact bytes, ragged tail, both modes"""
    import torch
    from extended_legged_gym_b200 import _lib
    lib = _lib.load()
    n = 16 * 40001                                     # not a multiple of the 16 KB pieces of mode 1
    host = torch.arange(n, dtype=torch.int64).to(torch.uint8).pin_memory()
    for mode in (0, 1):
        dev = torch.zeros(n, dtype=torch.uint8, device="cuda:0")
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.elg_stage_block(dev.data_ptr(), host.data_ptr(), n, mode, 0, s), "elg_stage_block")
        torch.cuda.synchronize()
        assert torch.equal(dev.cpu(), host)
        back = torch.zeros(n, dtype=torch.uint8).pin_memory()
        _lib.check(lib.elg_stage_block(back.data_ptr(), dev.data_ptr(), n, mode, 7, s), "elg_stage_block")
        torch.cuda.synchronize()
        assert torch.equal(back, host)
    assert lib.elg_stage_block(None, host.data_ptr(), 16, 0, 0, None) == -4
    assert lib.elg_stage_block(host.data_ptr(), host.data_ptr(), 24, 0, 0, None) == -1
