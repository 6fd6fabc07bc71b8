"""The rollout's per-step policy work as one CUDA kernel (b2d_policy_act, csrc/policy_kernels.cuh).

`FusedPolicyStep` binds a `DronePolicy` (= pufferlib.models.Default for a Box action space,
models.py:41-98) and a vec env's contract buffers; `act(k-th experience row)` does what
PuffeRL.evaluate does between recv() and send() (pufferl.py:229-296): forward_eval, Normal sample +
summed log-prob (pytorch.py:189-199), reward clamp, the experience stores, action clip -- without
materialising the hidden layer in HBM.  The weights are read in place from the module's parameters
(float32, contiguous), so optimiser updates are seen by later calls and by CUDA-graph replays.
There is no CPU path.
"""
import ctypes as C

import torch

from . import capi


def _ptr(t):
    return t.data_ptr() if t is not None else None


class FusedPolicyStep:
    def __init__(self, policy, observations, rewards, terminals, env_actions, noise_seed=0, row_id_base=0,
                 deterministic=False, precision="tf32"):
        """precision: "tf32" = both Linear layers as TF32 tensor-core GEMMs with float32 accumulation, which is what
        the reference runs on a GPU (torch.set_float32_matmul_precision('high'), pufferl.py:55); "fp32" = float32 FMAs."""
        if precision not in ("tf32", "fp32"):
            raise ValueError("precision must be 'tf32' or 'fp32'")
        enc, mean, value = policy.encoder[0], policy.decoder_mean, policy.value
        params = (enc.weight, enc.bias, mean.weight, mean.bias, policy.decoder_logstd, value.weight, value.bias)
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("policy parameters must be contiguous float32 CUDA tensors")
        rows, obs_dim = observations.shape
        if enc.weight.shape[1] != obs_dim or mean.weight.shape[0] != 4 or value.weight.shape[0] != 1:
            raise ValueError("policy shape does not match the env (obs_dim inputs, 4 action means, 1 value)")
        if observations.dtype != torch.float32 or rewards.dtype != torch.float32 or env_actions.dtype != torch.float32:
            raise ValueError("observations / rewards / actions must be float32")
        if terminals.dtype not in (torch.uint8, torch.bool):
            raise ValueError("terminals must be uint8 or bool")
        for t in (observations, rewards, terminals, env_actions):
            if not t.is_cuda or not t.is_contiguous() or t.device != enc.weight.device:
                raise ValueError("env buffers must be contiguous CUDA tensors on the policy's device")
        self.device = enc.weight.device
        self._keep = (policy, observations, rewards, terminals, env_actions)
        self.weights = capi.PolicyWeights(*[p.data_ptr() for p in params], int(enc.weight.shape[0]),
                                          capi.POLICY_TF32 if precision == "tf32" else capi.POLICY_FP32)
        self.precision = precision
        self.rows, self.obs_dim = int(rows), int(obs_dim)
        self.noise_seed, self.row_id_base, self.deterministic = int(noise_seed), int(row_id_base), bool(deterministic)
        self.counter = torch.zeros(2, dtype=torch.int32, device=self.device)  # [calls completed, CTA arrivals]
        self.launches = 0

    @property
    def calls(self):
        return int(self.counter[0].item())

    def seek(self, call):
        """Set the call number the next act() draws its noise with."""
        self.counter.copy_(torch.tensor([int(call), 0], dtype=torch.int32))

    def act(self, observations=None, actions=None, logprobs=None, rewards=None, terminals=None, values=None, stream=None):
        """One policy step; the keyword tensors are the experience row to store (each optional)."""
        _, obs, rew, term, env_act = self._keep
        for name, t, shape in (("observations", observations, (self.rows, self.obs_dim)), ("actions", actions, (self.rows, 4)),
                               ("logprobs", logprobs, (self.rows,)), ("rewards", rewards, (self.rows,)),
                               ("terminals", terminals, (self.rows,)), ("values", values, (self.rows,))):
            if t is not None and (t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous() or t.device != self.device):
                raise ValueError(f"experience row `{name}` must be a contiguous float32 tensor of shape {shape} on {self.device}")
        io = capi.PolicyIO(obs.data_ptr(), rew.data_ptr(), term.data_ptr(), env_act.data_ptr(),
                           _ptr(observations), _ptr(actions), _ptr(logprobs), _ptr(rewards), _ptr(terminals), _ptr(values),
                           self.rows, self.obs_dim, self.row_id_base)
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        with torch.cuda.device(self.device):
            capi.check(capi.lib().b2d_policy_act(C.byref(self.weights), C.byref(io), C.c_uint64(self.noise_seed),
                                                 C.c_void_p(self.counter.data_ptr()), int(self.deterministic),
                                                 C.c_void_p(st.cuda_stream)))
        self.launches += 1
