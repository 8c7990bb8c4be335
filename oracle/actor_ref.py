"""numpy restatement of the reference actor forward and its hard Gumbel sampling.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * rls/model/ac_network_multi_gumbel.py:7-21 (TimeDistributed), :41-50 (layers),
    :52-67 (forward: relu(dense1) -> BiLSTM over the AGENT axis -> relu -> dense2
    / dense2_1 + dense2_2);
  * rls/model/ac_network_model_multi_gumbel.py:49,65 (extra ``dense3`` head);
  * rls/agent/multiagent/ddpg_gumbel_fix.py:59-61 (process_obs: float32 cast),
    :86-107 (get_exploration_action), :109-116 (gumbel_softmax, hard=True).
PINNED: tests/golden/actor_*.npz holds outputs of the reference's own
``ActorNetwork`` (imported from /root/reference by oracle/gen_golden.py in the
authoring container); tests/test_oracle.py checks this file against them.

LSTM conventions (torch.nn.LSTM): gate rows [0:H]=i, [H:2H]=f, [2H:3H]=g,
[3H:4H]=o; both biases added; zero initial (h, c); the reverse direction scans
agents N-1..0 and its output for agent t is stored at index t.
"""
import numpy as np

HID = 64
H = 32


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """x [B,N,64] -> out [B,N,H]."""
    B, N, _ = x.shape
    h = np.zeros((B, H), dtype=x.dtype)
    c = np.zeros((B, H), dtype=x.dtype)
    out = np.zeros((B, N, H), dtype=x.dtype)
    order = range(N - 1, -1, -1) if reverse else range(N)
    for t in order:
        g = x[:, t, :] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i_g = _sigmoid(g[:, 0:H])
        f_g = _sigmoid(g[:, H:2 * H])
        g_g = np.tanh(g[:, 2 * H:3 * H])
        o_g = _sigmoid(g[:, 3 * H:4 * H])
        c = f_g * c + i_g * g_g
        h = o_g * np.tanh(c)
        out[:, t, :] = h
    return out


def forward(sd, obs, dtype=np.float64):
    """sd: dict name -> array with the reference's state_dict keys
    (``dense1.module.weight`` ...).  obs [B,N,D].  Returns dict with
    ``logits`` (list of one or two [B,N,A_k] arrays) and, if ``dense3`` is
    present, ``next_state`` [B,N,D]."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    x = np.asarray(obs, dtype=np.float32).astype(dtype)  # process_obs casts to float32 first
    h1 = np.maximum(x @ p['dense1.module.weight'].T + p['dense1.module.bias'], 0.0)
    fwd = _lstm_dir(h1, p['bilstm.weight_ih_l0'], p['bilstm.weight_hh_l0'],
                    p['bilstm.bias_ih_l0'], p['bilstm.bias_hh_l0'], False)
    rev = _lstm_dir(h1, p['bilstm.weight_ih_l0_reverse'], p['bilstm.weight_hh_l0_reverse'],
                    p['bilstm.bias_ih_l0_reverse'], p['bilstm.bias_hh_l0_reverse'], True)
    hid = np.maximum(np.concatenate([fwd, rev], axis=-1), 0.0)
    out = {}
    if 'dense2.module.weight' in p:
        out['logits'] = [hid @ p['dense2.module.weight'].T + p['dense2.module.bias']]
    else:
        out['logits'] = [hid @ p['dense2_1.module.weight'].T + p['dense2_1.module.bias'],
                         hid @ p['dense2_2.module.weight'].T + p['dense2_2.module.bias']]
    if 'dense3.module.weight' in p:
        out['next_state'] = hid @ p['dense3.module.weight'].T + p['dense3.module.bias']
    return out


def sample_hard(logits, gumbel):
    """Index chosen by F.gumbel_softmax(hard=True): argmax(softmax(logits + g))
    == argmax(logits + g) (first maximum on ties).  logits/gumbel [B,N,A]."""
    return np.argmax(np.asarray(logits) + np.asarray(gumbel), axis=-1)


def top2_gap(logits, gumbel):
    z = np.sort(np.asarray(logits, dtype=np.float64) + np.asarray(gumbel, dtype=np.float64), axis=-1)
    return z[..., -1] - z[..., -2]


def init_state_dict(D, A, seed, model_head=False):
    """Random weights with torch's default-init *distribution* (U(-1/sqrt(fan_in), ..)) but drawn
    from numpy so the GPU box needs no reference import.  A: int or [A0, A1]."""
    rng = np.random.RandomState(seed)

    def U(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    sd = {'dense1.module.weight': U((HID, D), D), 'dense1.module.bias': U((HID,), D)}
    for sfx in ('', '_reverse'):
        sd['bilstm.weight_ih_l0' + sfx] = U((4 * H, HID), H)
        sd['bilstm.weight_hh_l0' + sfx] = U((4 * H, H), H)
        sd['bilstm.bias_ih_l0' + sfx] = U((4 * H,), H)
        sd['bilstm.bias_hh_l0' + sfx] = U((4 * H,), H)
    if isinstance(A, (list, tuple)):
        sd['dense2_1.module.weight'] = U((A[0], HID), HID)
        sd['dense2_1.module.bias'] = U((A[0],), HID)
        sd['dense2_2.module.weight'] = U((A[1], HID), HID)
        sd['dense2_2.module.bias'] = U((A[1],), HID)
    else:
        sd['dense2.module.weight'] = U((A, HID), HID)
        sd['dense2.module.bias'] = U((A,), HID)
    if model_head:
        sd['dense3.module.weight'] = U((D, HID), HID)
        sd['dense3.module.bias'] = U((D,), HID)
    return sd
