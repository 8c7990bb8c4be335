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
