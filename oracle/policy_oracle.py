"""TEST INFRASTRUCTURE (oracle): CPU restatement of the rollout's per-step policy work.

Follows, for a Box action space:
  * pufferlib/models.py:41-98  Default.forward_eval: hidden = GELU(Linear(obs)) (nn.GELU() = exact erf form),
    mean = decoder_mean(hidden), logstd = decoder_logstd.expand_as(mean), value = value(hidden)
  * pufferlib/pytorch.py:189-199  sample_logits(Normal): action = loc + scale * eps,
    log_prob = sum_i( -(a_i - loc_i)^2 / (2 scale_i^2) - log scale_i - log sqrt(2 pi) )   (torch.distributions.Normal)
  * pufferlib/pufferl.py:260,281,292-294  reward clamp to [-1, 1], terminals as float, action clip to the space.
Arithmetic is float64 (the device kernel is float32 with FMA; tests state the tolerance).  The noise
`eps` is the device's counter-based stream restated here: Philox4x32-10 (Salmon et al., SC'11; vectorised
below and checked against oracle/drone_oracle.c:orc_philox4x32_10 in tests/test_policy_cpu.py), counter
(global row, call number, 'POLI', 0), key = 64-bit seed, 23-bit uniforms, Box-Muller.

Only tests/ may import this module.
"""
import numpy as np
from scipy import special

POLI = 0x504F4C49


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; inputs uint32 arrays (broadcastable), returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def noise(rows, call, seed, row_id_base=0):
    """[rows, 4] standard normals of policy call number `call` (float64 Box-Muller of the 23-bit uniforms)."""
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row_id_base)).astype(np.uint32)
    w = philox4x32_10(r, np.uint32(call), np.uint32(POLI), np.uint32(0), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [((x >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23 for x in w]
    r0, r1 = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
    return np.stack([r0 * np.cos(2 * np.pi * u[1]), r0 * np.sin(2 * np.pi * u[1]),
                     r1 * np.cos(2 * np.pi * u[3]), r1 * np.sin(2 * np.pi * u[3])], axis=1)


def tf32(x):
    """float32 -> TF32 (10-bit mantissa), round to nearest, ties away from zero (PTX cvt.rna.tf32.f32): the operand
    rounding of a TF32 tensor-core GEMM, which is what torch.set_float32_matmul_precision('high') (pufferl.py:55)
    selects for the reference's float32 Linear layers on a GPU."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_truncate(x):
    """float32 -> TF32 by dropping the low 13 mantissa bits: what a tensor core does with a 32-bit operand that was
    not rounded first (the second GEMM of csrc/rollout_kernels.cuh reads the GELU activations this way)."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return (b & np.uint32(0xFFFFE000)).view(np.float32)


def forward_eval(w, obs, tf32_gemm=False, hidden="round"):
    """models.Default.forward_eval for a Box action space; w = dict of numpy arrays in nn.Linear layouts.
    tf32_gemm: round the operands of both Linear layers to TF32 first (products and sums stay exact/float64);
    hidden = "round" | "truncate": how the activations become TF32 operands of the second layer."""
    rnd = tf32 if tf32_gemm else (lambda z: np.asarray(z, dtype=np.float32))
    x = rnd(obs).astype(np.float64) @ rnd(w["encoder_weight"]).astype(np.float64).T + w["encoder_bias"].astype(np.float64)
    hidden_act = 0.5 * x * (1.0 + special.erf(x / np.sqrt(2.0)))
    if tf32_gemm:
        hidden_act = (tf32 if hidden == "round" else tf32_truncate)(hidden_act.astype(np.float32)).astype(np.float64)
    hidden = hidden_act
    mean = hidden @ rnd(w["decoder_mean_weight"]).astype(np.float64).T + w["decoder_mean_bias"].astype(np.float64)
    value = hidden @ rnd(w["value_weight"]).astype(np.float64).T + w["value_bias"].astype(np.float64)
    logstd = np.broadcast_to(w["decoder_logstd"].astype(np.float64).reshape(1, -1), mean.shape)
    return mean, logstd, value[:, 0]


def policy_act(w, obs, rewards, terminals, call, seed, row_id_base=0, deterministic=False, tf32_gemm=False, hidden="round"):
    """One policy step: returns dict(actions, logprobs, values, rewards, terminals, env_actions)."""
    mean, logstd, value = forward_eval(w, obs, tf32_gemm, hidden)
    std = np.exp(logstd)
    eps = np.zeros_like(mean) if deterministic else noise(obs.shape[0], call, seed, row_id_base)
    action = mean + std * eps
    logp = (-((action - mean) ** 2) / (2.0 * std * std) - logstd - np.log(np.sqrt(2.0 * np.pi))).sum(1)
    return dict(actions=action, logprobs=logp, values=value, rewards=np.clip(rewards.astype(np.float64), -1.0, 1.0),
                terminals=terminals.astype(np.float64), env_actions=np.clip(action, -1.0, 1.0), mean=mean, noise=eps)
