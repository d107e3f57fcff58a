"""d_inner-channel-sharded mixer for single very large images (BASELINE.json configs[4]: FastVim-T,
2048x2048, 16384 tokens, one image over 8 GPUs).

The reference has no counterpart (it is data-parallel only); the partition follows from
the structure of ``Mamba.forward`` (``mamba_ssm/modules/mamba_simple_faster.py:181-457``): the depthwise
conv, the pooling, the scan recurrence, the D skip and the z gate are per-channel
(``csrc/selective_scan/selective_scan_fwd_kernel.cuh:97-98`` launches grid (batch, dim)); the only
couplings across channels are ``x_proj`` (:321-323), the LayerNorm over d_inner (:437) and ``out_proj``
(:442-444).  With G ranks, rank r owns channels [r D/G, (r+1) D/G):

    in_proj       column-sharded (rows r of the x half and of the z half), input replicated  -> no comm
    K1 conv+pool  local channels
    x_proj        partial product over local channels -> all-reduce (2, B*Lp, R+2N) fp32   (11 KB at 2048^2)
    K2a scan      local channels
    K2b epilogue  writes the pre-norm value + per-token (sum, sum of squares) of the local channels
                  -> all-reduce (B, L, 2) fp32 (131 KB) -> fv_norm_gate_apply
    out_proj      "gather": all-gather the gated y (what BASELINE.json names), replicated GEMM; or
                  "reduce": row-sharded out_proj on the local channels + all-reduce of the (B, L, d_model)
                  output (half the bytes)

Everything outside the mixer (patch embed, add+norm, head) is replicated.  Inference only.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .mixer import Mamba


def shard_range(D: int, rank: int, world: int):
    if D % (4 * world) != 0:
        raise ValueError(f"d_inner={D} must be a multiple of 4*world={4 * world} to shard channels")
    n = D // world
    return rank * n, (rank + 1) * n


def shard_mixer_params(mixer: Mamba, rank: int, world: int, act_dtype: torch.dtype) -> Dict[str, Optional[torch.Tensor]]:
    """Kernel-ready parameter slices of rank `rank` (host logic; runs on any device)."""
    D = mixer.d_inner
    lo, hi = shard_range(D, rank, world)
    f32 = torch.float32
    with torch.no_grad():
        iw = mixer.in_proj.weight
        ib = mixer.in_proj.bias
        p = {
            "in_w": torch.cat([iw[lo:hi], iw[D + lo:D + hi]]).to(act_dtype).contiguous(),
            "in_b": None if ib is None else torch.cat([ib[lo:hi], ib[D + lo:D + hi]]).to(act_dtype).contiguous(),
            "conv_w": torch.stack([mixer.conv1d.weight[lo:hi, 0], mixer.conv1d_b.weight[lo:hi, 0]]).to(f32).contiguous(),
            "conv_b": None if mixer.conv1d.bias is None else
            torch.stack([mixer.conv1d.bias[lo:hi], mixer.conv1d_b.bias[lo:hi]]).to(f32).contiguous(),
            # x_proj contracts over channels: keep the local columns, (2, D_loc, R+2N)
            "x_w_t": torch.stack([mixer.x_proj.weight[:, lo:hi].t(), mixer.x_proj_b.weight[:, lo:hi].t()]).to(act_dtype).contiguous(),
            "dt_w": torch.stack([mixer.dt_proj.weight[lo:hi], mixer.dt_proj_b.weight[lo:hi]]).to(f32).contiguous(),
            "dt_b": torch.stack([mixer.dt_proj.bias[lo:hi], mixer.dt_proj_b.bias[lo:hi]]).to(f32).contiguous(),
            "A_log": torch.stack([mixer.A_log[lo:hi], mixer.A_b_log[lo:hi]]).to(f32).contiguous(),
            "D": torch.stack([mixer.D[lo:hi], mixer.D_b[lo:hi]]).to(f32).contiguous(),
            "ln_w": mixer.layernorm.weight[lo:hi].to(f32).contiguous() if mixer.use_norm_after_ssm else None,
            "ln_b": mixer.layernorm.bias[lo:hi].to(f32).contiguous() if mixer.use_norm_after_ssm else None,
            "out_w": mixer.out_proj.weight.to(act_dtype).contiguous(),
            "out_w_loc": mixer.out_proj.weight[:, lo:hi].to(act_dtype).contiguous(),
            "out_b": None if mixer.out_proj.bias is None else mixer.out_proj.bias.to(act_dtype),
        }
    return p


class ChannelShardedMamba(torch.nn.Module):
    """Wraps a (replicated) ``Mamba`` and runs its forward with d_inner split over the process group."""

    def __init__(self, mixer: Mamba, group=None, out_mode: str = "gather"):
        super().__init__()
        assert out_mode in ("gather", "reduce")
        self.mixer, self.group, self.out_mode = mixer, group, out_mode
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._cache = {}

    def _params(self, act_dtype):
        # keyed like Mamba._packed: any parameter update / load_state_dict / .to() re-shards
        key = (act_dtype,) + tuple((p.data_ptr(), p._version) for p in self.mixer.parameters())
        if self._cache.get("k") != key:
            self._cache = {"k": key, "v": shard_mixer_params(self.mixer, self.rank, self.world, act_dtype)}
        return self._cache["v"]

    def invalidate(self):
        """Drop the sharded copies (needed after parameter updates that do not bump ``_version``, e.g. optimizer steps
        replayed from a CUDA graph)."""
        self._cache = {}

    @torch.no_grad()
    def forward(self, hidden_states, inference_params=None, rotated: bool = False):
        m = self.mixer
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else hidden_states.dtype
        h = hidden_states.to(act_dtype)
        pk = self._params(act_dtype)
        geom = m.geometry(rotated)
        B, L, _ = h.shape
        D, G = m.d_inner, self.world
        Dl = D // G
        xz = F.linear(h, pk["in_w"], pk["in_b"])                          # (B, L, 2*Dl): local x | local z
        x, z = xz[..., :Dl], xz[..., Dl:]
        u = ops.conv_pool_fwd(x, geom, pk["conv_w"], pk["conv_b"], float(m.scaling_factor), m.collapse_method)
        xdbl = torch.bmm(u.view(2, B * geom.Lp, Dl).float(), pk["x_w_t"].float())   # partial over local channels
        dist.all_reduce(xdbl, group=self.group)
        xdbl = xdbl.to(act_dtype)
        s = ops.scan_fwd(u, xdbl, geom, m.dt_rank, m.d_state, pk["dt_w"], pk["dt_b"], pk["A_log"], a_is_log=True)
        eps = m.layernorm.eps if m.use_norm_after_ssm else 1e-5
        if m.use_norm_after_ssm:
            stats = torch.empty((B, L, 2), device=h.device, dtype=torch.float32)
            y = ops.gate_fwd(x, z, s, geom, pk["conv_w"], pk["conv_b"], pk["D"], pk["ln_w"], pk["ln_b"], eps, stats=stats)
            dist.all_reduce(stats, group=self.group)
            ops.norm_gate_apply(y, z, stats, geom, D, pk["ln_w"], pk["ln_b"], eps)
        else:
            y = ops.gate_fwd(x, z, s, geom, pk["conv_w"], pk["conv_b"], pk["D"], None, None, eps)
        if self.out_mode == "reduce":
            out = F.linear(y, pk["out_w_loc"])                             # partial over local channels
            dist.all_reduce(out, group=self.group)
            if pk["out_b"] is not None:
                out = out + pk["out_b"]
        else:
            ys = torch.empty((G, B, L, Dl), device=h.device, dtype=act_dtype)
            dist.all_gather_into_tensor(ys, y.contiguous(), group=self.group)
            y_full = ys.permute(1, 2, 0, 3).reshape(B, L, D)
            out = F.linear(y_full, pk["out_w"], pk["out_b"])
        if m.init_layer_scale is not None:
            out = out * m.gamma
        return out


def shard_model_channels(model, group=None, out_mode: str = "gather"):
    """Replaces every block's mixer of a ``fastvim_b200.vision.VisionMamba`` by its channel-sharded form."""
    for blk in model.layers:
        blk.mixer = ChannelShardedMamba(blk.mixer, group, out_mode)
    return model
