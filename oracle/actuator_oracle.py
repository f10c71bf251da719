"""CPU oracle of the actuator-network torque path -- TEST INFRASTRUCTURE, never on the product path.

Restates ``Anymal._compute_torques`` with ``use_actuator_network`` (envs/anymal_c/anymal.py:93-105 in
/root/reference/legged_gym/legged_gym) and the TorchScript module it calls,
``resources/actuator_nets/anydrive_v3_lstm.pt`` (class ``LSTMsea``: ``x * in_scale`` -> ``nn.LSTM(2, 8, num_layers=2,
batch_first=True)`` on one time step -> ``nn.Linear(8, 1)`` -> ``* out_scale``; structure read off ``module.code`` and the
parameter shapes).  Pinned: ``tests/test_actuator_net.py`` compares this file with the UNMODIFIED TorchScript module in the
container (``/root/reference`` present) and with ``tests/golden/actuator_net.npz`` generated from that module by
``tests/golden/make_actuator_golden.py`` (three consecutive ``_compute_torques`` calls, hidden state carried).
"""
import torch

WEIGHT_KEYS = ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
               "lstm.weight_ih_l1", "lstm.weight_hh_l1", "lstm.bias_ih_l1", "lstm.bias_hh_l1", "linear.weight", "linear.bias")


def _cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    # torch.nn.LSTM cell, gate order (i, f, g, o)
    g = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i, f, gg, o = g.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


class ActuatorNetOracle:
    def __init__(self, weights):
        """weights: dict with WEIGHT_KEYS + 'in_scale' [2] + 'out_scale' [1] (float32 tensors)"""
        self.w = {k: torch.as_tensor(v, dtype=torch.float).clone() for k, v in weights.items()}

    def forward(self, x, hidden, cell):
        """x [R,2]; hidden / cell [2,R,8] -> (torques [R], hidden', cell')  (LSTMsea.forward on a [R,1,2] batch)"""
        w = self.w
        x0 = x * w["in_scale"].view(1, 2)
        h0, c0 = _cell(x0, hidden[0], cell[0], w["lstm.weight_ih_l0"], w["lstm.weight_hh_l0"], w["lstm.bias_ih_l0"], w["lstm.bias_hh_l0"])
        h1, c1 = _cell(h0, hidden[1], cell[1], w["lstm.weight_ih_l1"], w["lstm.weight_hh_l1"], w["lstm.bias_ih_l1"], w["lstm.bias_hh_l1"])
        y = (h1 @ w["linear.weight"].t() + w["linear.bias"]).squeeze(1)
        return w["out_scale"] * y, torch.stack([h0, h1]), torch.stack([c0, c1])

    def compute_torques(self, actions, action_scale, default_dof_pos, dof_pos, dof_vel, hidden, cell):
        """anymal.py:96-103: sea_input[:,0,0] = (a * scale + q0 - q).flatten(); sea_input[:,0,1] = qd.flatten()"""
        x = torch.stack([(actions * action_scale + default_dof_pos - dof_pos).flatten(), dof_vel.flatten()], dim=1)
        t, h, c = self.forward(x, hidden, cell)
        return t.view(actions.shape), h, c
