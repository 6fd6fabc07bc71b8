"""compute_puff_advantage on the GPU through the C ABI (b2d_puff_advantage).

Same call shape as pufferlib.pufferl.compute_puff_advantage (pufferl.py:639-659): float32 CUDA
tensors values / rewards / terminals / ratio / advantages, in place, returns `advantages`.
Row-major [segments, horizon] tensors (the reference's layout) and the time-major
[horizon, num_agents] experience of drone_b200.rollout.DeviceRollout are both accepted:
pass `time_major=True` for the latter.  There is no CPU path.
"""
import ctypes as C

import torch

from . import capi


def compute_puff_advantage(values, rewards, terminals, ratio, advantages, gamma, gae_lambda, vtrace_rho_clip,
                           vtrace_c_clip, time_major=False, priority=None, math="fast", stream=None):
    ts = (values, rewards, terminals, ratio, advantages)
    for t in ts:
        if t.dim() != 2:
            raise ValueError("Tensor must be 2D")
        if t.dtype != torch.float32:
            raise ValueError("All tensors must be float32")
        if not t.is_cuda or t.device != values.device:
            raise ValueError("All tensors must be on the same CUDA device")
        if t.shape != values.shape:
            raise ValueError("All tensors must have the same shape")
        if not t.is_contiguous():
            raise ValueError("All tensors must be contiguous")
    if time_major:
        horizon, rows = values.shape
        row_stride, t_stride = 1, rows
    else:
        rows, horizon = values.shape
        row_stride, t_stride = horizon, 1
    if priority is not None and (priority.dtype != torch.float32 or priority.numel() != rows or not priority.is_cuda):
        raise ValueError("priority must be a float32 CUDA tensor with one element per row")
    st = torch.cuda.current_stream(values.device) if stream is None else stream
    with torch.cuda.device(values.device):
        capi.check(capi.lib().b2d_puff_advantage(
            C.c_void_p(values.data_ptr()), C.c_void_p(rewards.data_ptr()), C.c_void_p(terminals.data_ptr()),
            C.c_void_p(ratio.data_ptr()), C.c_void_p(advantages.data_ptr()),
            C.c_void_p(priority.data_ptr()) if priority is not None else None,
            int(rows), int(horizon), int(row_stride), int(t_stride), float(gamma), float(gae_lambda),
            float(vtrace_rho_clip), float(vtrace_c_clip), {"fast": 0, "strict": 1}[math], C.c_void_p(st.cuda_stream)))
    return advantages
