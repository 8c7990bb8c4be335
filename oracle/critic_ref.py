"""numpy restatement of the reference critics' forward pass.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * rls/model/ac_network_multi_gumbel.py:70-148 (CriticNetwork of the main.py path):
      cat(obs, action) -> relu(dense1) -> LSTM(64 -> 64) over the AGENT axis -> dot-product attention of every
      step's output with the final hidden state (:103-110) -> softmax over agents -> weighted sum -> relu (:141)
      -> dense2;
  * rls/model/ac_network_model_multi_gumbel.py:69-143 (the "+model" critic of main_scalability_*):
      the same up to the attention output, NO relu after it (:139), two heads Q = dense2, r = dense3 (:140-141).
PINNED: tests/golden/critic_*.npz holds outputs of the reference's own CriticNetwork classes (oracle/gen_golden.py,
authoring container); tests/test_oracle.py checks this file against them.

torch.nn.LSTM conventions: gate rows [0:H]=i, [H:2H]=f, [2H:3H]=g, [3H:4H]=o; both biases added; zero initial (h, c).
"""
import numpy as np

H = 64


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def forward(sd, obs, action, dtype=np.float64):
    """sd: the reference's state_dict (``dense1.module.weight``, ``lstm.weight_ih_l0``, ``dense2.weight``, optional
    ``dense3.weight``).  obs [B,N,D], action [B,N,A] (or a list of such, concatenated like critic.forward does).
    Returns {'q': [B,out]} plus {'r': [B,out]} for the model critic."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    acts = action if isinstance(action, (list, tuple)) else [action]
    x = np.concatenate([np.asarray(obs, dtype=np.float32)] + [np.asarray(a, dtype=np.float32) for a in acts], -1).astype(dtype)
    B, N, _ = x.shape
    h1 = np.maximum(x @ p['dense1.module.weight'].T + p['dense1.module.bias'], 0.0)
    h = np.zeros((B, H), dtype=dtype)
    c = np.zeros((B, H), dtype=dtype)
    out = np.zeros((B, N, H), dtype=dtype)
    for t in range(N):
        g = h1[:, t] @ p['lstm.weight_ih_l0'].T + p['lstm.bias_ih_l0'] + h @ p['lstm.weight_hh_l0'].T + p['lstm.bias_hh_l0']
        i_g, f_g = _sigmoid(g[:, 0:H]), _sigmoid(g[:, H:2 * H])
        g_g, o_g = np.tanh(g[:, 2 * H:3 * H]), _sigmoid(g[:, 3 * H:4 * H])
        c = f_g * c + i_g * g_g
        h = o_g * np.tanh(c)
        out[:, t] = h
    attn = np.einsum('bth,bh->bt', out, h)                      # bmm(lstm_output, final_hidden)
    attn = np.exp(attn - attn.max(axis=1, keepdims=True))
    attn = attn / attn.sum(axis=1, keepdims=True)               # softmax over the agent axis
    new_h = np.einsum('bth,bt->bh', out, attn)
    model = 'dense3.weight' in p
    if not model:
        new_h = np.maximum(new_h, 0.0)
    res = {'q': new_h @ p['dense2.weight'].T + p['dense2.bias']}
    if model:
        res['r'] = new_h @ p['dense3.weight'].T + p['dense3.bias']
    return res
