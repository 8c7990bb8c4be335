"""A torch module with the same parameter names and math as the reference actor, for users who
do not have the reference tree on the path (weights initialisation, checkpoints, training).  It is
NOT the acting path - that is ``FusedActor`` - and it is what the fp32 kernel is compared with.

Structure per rls/model/ac_network_multi_gumbel.py:41-50 (and ac_network_model_multi_gumbel.py:49):
Linear(D,64) -> ReLU -> BiLSTM(64 -> 2x32) over the agent axis -> ReLU -> Linear(64, A) head(s).
"""
import torch
import torch.nn as nn


class _PerAgent(nn.Module):
    """Applies ``module`` to every (batch, agent) row; parameters live under ``.module`` so the
    state_dict keys match the reference's ``TimeDistributed`` wrapper (``dense1.module.weight``)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        return self.module(x.reshape(-1, x.shape[-1])).reshape(*x.shape[:-1], -1)


class ActorNetwork(nn.Module):
    def __init__(self, input_dim, out_dim, model_head=False):
        super().__init__()
        self.out_dim = out_dim
        self.dense1 = _PerAgent(nn.Linear(input_dim, 64))
        self.bilstm = nn.LSTM(64, 32, num_layers=1, batch_first=True, bidirectional=True)
        if isinstance(out_dim, (list, tuple)):
            self.dense2_1 = _PerAgent(nn.Linear(64, out_dim[0]))
            self.dense2_2 = _PerAgent(nn.Linear(64, out_dim[1]))
        else:
            self.dense2 = _PerAgent(nn.Linear(64, out_dim))
        if model_head:
            self.dense3 = _PerAgent(nn.Linear(64, input_dim))
        self.model_head = model_head

    def forward(self, obs):
        hid = torch.relu(self.dense1(obs))
        hid, _ = self.bilstm(hid, None)
        hid = torch.relu(hid)
        if isinstance(self.out_dim, (list, tuple)):
            policy = [self.dense2_1(hid), self.dense2_2(hid)]
        else:
            policy = self.dense2(hid)
        if self.model_head:
            return policy, self.dense3(hid)
        return policy


def random_state_dict(D, A, seed, model_head=False):
    """Random actor weights with torch's default-init distribution (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) under the
    reference's state_dict keys (rls/model/ac_network_multi_gumbel.py:24-50), drawn from numpy so that benchmarks and
    profiling targets get the same weights on every box without touching torch's RNG.  A: int or [A0, A1]."""
    import numpy as np
    hid, h = 64, 32
    rng = np.random.RandomState(seed)

    def U(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    sd = {'dense1.module.weight': U((hid, D), D), 'dense1.module.bias': U((hid,), D)}
    for sfx in ('', '_reverse'):
        sd['bilstm.weight_ih_l0' + sfx] = U((4 * h, hid), h)
        sd['bilstm.weight_hh_l0' + sfx] = U((4 * h, h), h)
        sd['bilstm.bias_ih_l0' + sfx] = U((4 * h,), h)
        sd['bilstm.bias_hh_l0' + sfx] = U((4 * h,), h)
    if isinstance(A, (list, tuple)):
        sd['dense2_1.module.weight'] = U((A[0], hid), hid)
        sd['dense2_1.module.bias'] = U((A[0],), hid)
        sd['dense2_2.module.weight'] = U((A[1], hid), hid)
        sd['dense2_2.module.bias'] = U((A[1],), hid)
    else:
        sd['dense2.module.weight'] = U((A, hid), hid)
        sd['dense2.module.bias'] = U((A,), hid)
    if model_head:
        sd['dense3.module.weight'] = U((D, hid), hid)
        sd['dense3.module.bias'] = U((D,), hid)
    return sd
