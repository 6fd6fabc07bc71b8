"""On-device rollout: policy forward, action sampling and the env step alternate on one stream
with no host synchronisation; a K-step collection is captured once in a CUDA graph and replayed.

This is the device-resident counterpart of PuffeRL.evaluate (pufferlib/pufferl.py:214-314) for
envs whose buffers already live on the GPU: what that loop does per step -- recv (obs, reward,
done), `policy.forward_eval`, `sample_logits` on a Normal (pufferlib/pytorch.py:189-199), store
(obs, action, logprob, reward clamped to [-1, 1], done, value), clip the action to the action
space, send -- happens here without the per-step `.to(device)` / `.cpu().numpy()` round trips
(pufferl.py:240-243,292) that cap any GPU env at ~1e6 steps/s.  Three forms:

  policy_impl="rollout_kernel"  the WHOLE K-step loop as one kernel (b2d_race_rollout,
      csrc/rollout_kernels.cuh): 128 envs per CTA stay in registers for all K steps, both Linear
      layers run as tcgen05 tensor-core GEMMs with TMEM accumulators, only the experience row is
      written per step.  Race envs + DronePolicy(hidden 128); what "auto" picks when it applies.
  policy_impl="fused"  two kernels per step: the fused policy step (drone_b200.policy /
      csrc/policy_kernels.cuh: MLP, sampling, log-prob, experience stores and action clip in one
      launch) and the env step behind `vec.step()`; K steps captured in one CUDA graph.
  policy_impl="torch"  the same loop on plain torch ops (library GEMMs) -- the form any other
      policy class uses, and the cross-check of the other two in the tests.

Experience is stored time-major, [horizon, num_agents, ...], so every store is one contiguous
write; `segments()` returns the reference's [num_agents, horizon, ...] views.
"""
import ctypes as C
import math

import torch
from torch import nn

from . import capi


class DronePolicy(nn.Module):
    """pufferlib.models.Default for a Box action space (models.py:41-63,86-98): Linear(obs,128) + GELU,
    mean head (std 0.01 init), state-independent log-std parameter, value head."""

    def __init__(self, obs_dim=29, act_dim=4, hidden_size=128):
        super().__init__()
        self.encoder = nn.Sequential(nn.Linear(obs_dim, hidden_size), nn.GELU())
        self.decoder_mean = nn.Linear(hidden_size, act_dim)
        self.decoder_logstd = nn.Parameter(torch.zeros(1, act_dim))
        self.value = nn.Linear(hidden_size, 1)
        nn.init.orthogonal_(self.encoder[0].weight, math.sqrt(2))
        nn.init.orthogonal_(self.decoder_mean.weight, 0.01)
        nn.init.orthogonal_(self.value.weight, 1.0)
        for lin in (self.encoder[0], self.decoder_mean, self.value):
            nn.init.constant_(lin.bias, 0.0)

    def forward_eval(self, observations, state=None):
        hidden = self.encoder(observations.float())
        mean = self.decoder_mean(hidden)
        return mean, self.decoder_logstd.expand_as(mean), self.value(hidden)

    forward = forward_eval


class DeviceRollout:
    def __init__(self, vec, policy, horizon=128, use_graph=True, deterministic=False, autocast=None,
                 policy_impl="auto", noise_seed=0, precision="tf32"):
        self.vec, self.policy, self.horizon = vec, policy, int(horizon)
        self.deterministic, self.autocast = deterministic, autocast
        kernel_ok = (isinstance(policy, DronePolicy) and autocast is None and getattr(vec, "obs_dim", 0) == 29 and
                     policy.encoder[0].weight.shape == (128, 29) and precision == "tf32")
        if policy_impl == "auto":
            policy_impl = "rollout_kernel" if kernel_ok else ("fused" if (isinstance(policy, DronePolicy) and autocast is None) else "torch")
        if policy_impl not in ("rollout_kernel", "fused", "torch"):
            raise ValueError("policy_impl must be 'auto', 'rollout_kernel', 'fused' or 'torch'")
        if policy_impl == "rollout_kernel" and not kernel_ok:
            raise ValueError("policy_impl='rollout_kernel' needs a race vec, a DronePolicy with hidden_size=128, precision='tf32' and no autocast")
        self.policy_impl = policy_impl
        self.noise_seed = int(noise_seed)
        self.fused = None
        self._kernel_launches = 0
        if policy_impl == "rollout_kernel":
            enc, mean, value = policy.encoder[0], policy.decoder_mean, policy.value
            params = (enc.weight, enc.bias, mean.weight, mean.bias, policy.decoder_logstd, value.weight, value.bias)
            for p in params:
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or p.device != vec.device:
                    raise ValueError("policy parameters must be contiguous float32 CUDA tensors on the env's device")
            self._weights = capi.PolicyWeights(*[p.data_ptr() for p in params], 128, capi.POLICY_TF32)
            self.counter = torch.zeros(2, dtype=torch.int32, device=vec.device)  # [policy calls completed, CTA arrivals]
        if policy_impl == "fused":
            from .policy import FusedPolicyStep
            self.fused = FusedPolicyStep(policy, vec.observations, vec.rewards, vec.terminals, vec.actions,
                                         noise_seed=noise_seed, row_id_base=getattr(vec, "row_id_base", 0),
                                         deterministic=deterministic, precision=precision)
        n, dev = vec.num_agents, vec.device
        k = self.horizon
        self.observations = torch.zeros((k, n, vec.obs_dim), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((k, n, 4), dtype=torch.float32, device=dev)
        self.logprobs = torch.zeros((k, n), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((k, n), dtype=torch.float32, device=dev)
        self.terminals = torch.zeros((k, n), dtype=torch.float32, device=dev)
        self.values = torch.zeros((k, n), dtype=torch.float32, device=dev)
        self.graph = None
        self.use_graph = use_graph
        self.env_steps = 0
        self._graph_launches_per_replay = 0
        self._replays = 0

    @property
    def kernel_name(self):
        if self.policy_impl == "rollout_kernel":
            return "race_rollout_kernel (tcgen05 policy GEMMs + env step, K steps per launch)"
        if self.policy_impl == "fused":
            return f"policy_act_{self.fused.precision}_kernel + race_step_kernel (two launches per step)"
        return "torch policy ops + race_step_kernel"

    @property
    def kernel_launches(self):
        """Kernels of this package launched for the rollout so far (graph replays re-launch the captured ones)."""
        eager = self.vec.kernel_launches + (self.fused.launches if self.fused is not None else 0)
        return eager + self._graph_launches_per_replay * max(0, self._replays - 1)

    def _launch_rollout_kernel(self):
        vec = self.vec
        store = capi.RolloutStore(self.observations.data_ptr(), self.actions.data_ptr(), self.logprobs.data_ptr(),
                                  self.rewards.data_ptr(), self.terminals.data_ptr(), self.values.data_ptr())
        with torch.cuda.device(vec.device):
            st = torch.cuda.current_stream(vec.device)
            capi.check(capi.lib().b2d_race_rollout(vec.h, C.byref(self._weights), C.byref(store), self.horizon,
                                                   C.c_uint64(self.noise_seed), C.c_void_p(self.counter.data_ptr()),
                                                   int(self.deterministic), C.c_void_p(st.cuda_stream)))

    @torch.no_grad()
    def _one_step(self, k):
        vec = self.vec
        if self.fused is not None:
            self.fused.act(self.observations[k], self.actions[k], self.logprobs[k], self.rewards[k], self.terminals[k],
                           self.values[k])
            vec.step()
            return
        obs = vec.observations
        self.observations[k].copy_(obs)
        torch.clamp(vec.rewards, -1.0, 1.0, out=self.rewards[k])          # pufferl.py:260
        self.terminals[k].copy_(vec.terminals)                              # d.float(), pufferl.py:281
        if self.autocast is not None:
            with torch.autocast("cuda", dtype=self.autocast):
                mean, logstd, value = self.policy.forward_eval(obs)
            mean, value = mean.float(), value.float()
        else:
            mean, logstd, value = self.policy.forward_eval(obs)
        if self.deterministic:
            action = mean
            self.logprobs[k].copy_((-logstd - 0.5 * math.log(2.0 * math.pi)).sum(1))
        else:
            noise = torch.randn_like(mean)
            action = torch.addcmul(mean, logstd.exp(), noise)               # Normal(mean, std).sample()
            self.logprobs[k].copy_((-0.5 * noise * noise - logstd - 0.5 * math.log(2.0 * math.pi)).sum(1))
        self.actions[k].copy_(action)
        self.values[k].copy_(value.flatten())
        torch.clamp(action, -1.0, 1.0, out=vec.actions)                     # np.clip to the action space, pufferl.py:293-294
        vec.step()

    @torch.no_grad()
    def _collect_eager(self):
        for k in range(self.horizon):
            self._one_step(k)

    @torch.no_grad()
    def collect(self):
        """One horizon of experience.  Asynchronous: returns after enqueueing (graph replay)."""
        if self.policy_impl == "rollout_kernel":
            # one launch is the whole K-step loop: nothing for a graph to save; state that must advance between
            # collections (noise call counter, step counter, episode numbers) lives on the device
            self._launch_rollout_kernel()
            self.env_steps += self.horizon
            return self
        if not self.use_graph:
            self._collect_eager()
        else:
            if self.graph is None:
                side = torch.cuda.Stream(device=self.vec.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):  # warm-up outside capture (cuBLAS workspaces, lazy inits)
                    for _ in range(2):
                        self._one_step(0)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self.env_steps += 2
                self.graph = torch.cuda.CUDAGraph()
                before = self.vec.kernel_launches + (self.fused.launches if self.fused is not None else 0)
                with torch.cuda.graph(self.graph):
                    self._collect_eager()
                self._graph_launches_per_replay = (self.vec.kernel_launches +
                                                   (self.fused.launches if self.fused is not None else 0)) - before
            self.graph.replay()
            self._replays += 1
        self.env_steps += self.horizon
        return self

    def segments(self):
        """The reference's experience layout [segments = num_agents, horizon, ...] (views)."""
        t = lambda x: x.transpose(0, 1)  # noqa: E731
        return dict(observations=t(self.observations), actions=t(self.actions), logprobs=t(self.logprobs),
                    rewards=t(self.rewards), terminals=t(self.terminals), values=t(self.values))
