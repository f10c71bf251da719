"""Generate tests/golden/actuator_net.npz from the UNMODIFIED reference (container only):
the TorchScript actuator network resources/actuator_nets/anydrive_v3_lstm.pt driven exactly as
Anymal._compute_torques does (envs/anymal_c/anymal.py:93-105), three consecutive calls with the hidden state carried.
The network's parameters travel in the fixture: they are the inputs of the computation under test.

    python tests/golden/make_actuator_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PT = "/root/reference/legged_gym/resources/actuator_nets/anydrive_v3_lstm.pt"


def main():
    net = torch.jit.load(PT, map_location="cpu")
    out = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    out["in_scale"] = net.in_scale.detach().reshape(2).numpy()
    out["out_scale"] = net.out_scale.detach().reshape(1).numpy()
    n, d, scale = 64, 12, 0.5
    g = torch.Generator().manual_seed(0)
    q0 = torch.randn(1, d, generator=g) * 0.4
    sea_input = torch.zeros(n * d, 1, 2)
    hidden = torch.zeros(2, n * d, 8)
    cell = torch.zeros(2, n * d, 8)
    out["default_dof_pos"], out["action_scale"] = q0.numpy(), np.array([scale], dtype=np.float32)
    for step in range(3):
        actions = torch.randn(n, d, generator=g)
        dof_pos = q0 + torch.randn(n, d, generator=g) * 0.3
        dof_vel = torch.randn(n, d, generator=g) * 2.0
        with torch.inference_mode():      # anymal.py:96-103, verbatim data flow
            sea_input[:, 0, 0] = (actions * scale + q0 - dof_pos).flatten()
            sea_input[:, 0, 1] = dof_vel.flatten()
            torques, (hidden[:], cell[:]) = net(sea_input, (hidden, cell))
        out[f"s{step}__actions"], out[f"s{step}__dof_pos"], out[f"s{step}__dof_vel"] = actions.numpy(), dof_pos.numpy(), dof_vel.numpy()
        out[f"s{step}__torques"] = torques.clone().numpy().reshape(n, d)
        out[f"s{step}__hidden"], out[f"s{step}__cell"] = hidden.clone().numpy(), cell.clone().numpy()
    path = os.path.join(ROOT, "tests", "golden", "actuator_net.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
