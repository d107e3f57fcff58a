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


# =====================================================================================================================
# Hybrid token / channel sharding over NVLink peer memory (round 2): no NCCL on the data path
# =====================================================================================================================
import ctypes as _C

from . import _lib
from .norm import RMSNorm, layer_norm_fn


class PeerBuffer:
    """One symmetric buffer per rank (same layout everywhere), mapped into every process with
    ``torch.distributed._symmetric_memory`` -- torch is the plumbing that allocates and exchanges the handles; the
    exchanges themselves are ``fv_peer_sum_f32`` / ``fv_peer_copy2d`` (csrc/peer.cu): barrier + direct NVLink reads in one
    kernel each."""

    def __init__(self, nbytes: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.header = int(_lib.lib().fv_peer_header_bytes())
        self.nbytes = self.header + int(nbytes)
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(self.group)      # every rank's flag words are zero before anyone signals
        self.ptrs = (_C.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        self._cursor = self.header

    def carve(self, shape, dtype) -> "tuple[torch.Tensor, int]":
        """A (1024-byte aligned) region of the local buffer as a tensor, and its byte offset (the same on every rank)."""
        n = 1
        for d in shape:
            n *= int(d)
        nb = n * torch.empty((), dtype=dtype).element_size()
        off = (self._cursor + 1023) // 1024 * 1024
        if off + nb > self.nbytes:
            raise RuntimeError("PeerBuffer: out of symmetric memory")
        self._cursor = off + nb
        return self.buf[off:off + nb].view(dtype).view(*shape), off

    def sum_f32(self, off: int, n: int, out32=None, out16=None):
        stream = _C.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)
        _lib.call("fv_peer_sum_f32", self.world, self.rank, self.ptrs, int(off), int(n),
                  None if out32 is None else _C.c_void_p(out32.data_ptr()),
                  None if out16 is None else _C.c_void_p(out16.data_ptr()), stream)

    def copy2d(self, nparts: int, rows: int, row_bytes: int, src_off, src_ld: int, dst_off, dst_ld: int, dst: torch.Tensor):
        so = (_C.c_int64 * (2 * self.world))(*[int(v) for v in src_off])
        do = (_C.c_int64 * (2 * self.world))(*[int(v) for v in dst_off])
        stream = _C.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)
        _lib.call("fv_peer_copy2d", self.world, self.rank, self.ptrs, int(nparts), int(rows), int(row_bytes), so, int(src_ld),
                  do, int(dst_ld), _C.c_void_p(dst.data_ptr()), stream)

    def error(self) -> int:
        return int(_lib.lib().fv_peer_error(_C.c_void_p(self.buf.data_ptr())))


def hybrid_supported(model, world: int, img_hw) -> bool:
    """Token bands must be whole token rows and the channel shards 16-byte multiples."""
    gh, gw = model.token_size
    L = gh * gw
    D = model.layers[0].mixer.d_inner
    return (world <= 8 and L % world == 0 and (L // world) % gw == 0 and D % (8 * world) == 0
            and model.final_pool_type == "mean" and model.if_abs_pos_embed)


class HybridShardedVisionMamba(torch.nn.Module):
    """FastVim forward of ONE large image over G GPUs (BASELINE.json configs[4]) with every heavy op sharded:

        token-sharded   (each rank: L/G tokens = a band of token rows)   patch embed, add + RMSNorm, in_proj, out_proj
        channel-sharded (each rank: d_inner/G channels, all tokens)      conv + pool, x_proj partial, scan, gate

    joined by peer-memory exchanges (csrc/peer.cu), three per block:
        A  all-to-all  token -> channel of the x half of the in_proj output   (L x d_inner/G per rank, bf16: 1.6 MB at 2048^2)
        B  sum of the x_proj partial products                                  (2 x Lp x (R + 2N) fp32: 45 KB)
        D  all-to-all  channel -> token of the pre-norm merged value v         (L/G x d_inner per rank: 1.6 MB)
    z stays on the token side and LayerNorm + gate run there (fv_ln_gate_fwd), where a rank holds every channel of its
    tokens, so the LayerNorm statistics never cross GPUs.
    Nothing is replicated and no NCCL collective runs on the data path (round 1: three NCCL collectives per block and
    add_norm / out_proj / patch embed replicated on every rank; slower than one GPU).  Follows the structure of
    ``Mamba.forward`` (mamba_simple_faster.py:181-457): conv, pool, scan recurrence, D skip and gate are per channel,
    x_proj (:321-323) and the LayerNorm (:437) couple channels, everything else is per token.  Inference only."""

    def __init__(self, model, group=None):
        super().__init__()
        self.model, self.group = model, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._state = {}
        model._shard_desc = (f"hybrid token/channel sharding x{self.world}: token-sharded patch embed / add+norm / in_proj / "
                             f"LayerNorm+gate / out_proj, channel-sharded conv+pool / scan / D skip; 3 peer-memory exchanges per block "
                             f"(fv_peer_copy2d, fv_peer_sum_f32 over NVLink), no NCCL on the data path")

    # ---- per-dtype state: sharded parameters + symmetric buffer ------------------------------------------------
    def _prepare(self, act_dtype, device):
        key = (act_dtype,) + tuple((p.data_ptr(), p._version) for p in self.model.parameters())
        st = self._state.get("v")
        if st is not None and self._state.get("k") == key:
            return st
        m, G, r = self.model, self.world, self.rank
        gh, gw = m.token_size
        L = gh * gw
        Lt = L // G
        mix0 = m.layers[0].mixer
        D, R, N, dm = mix0.d_inner, mix0.dt_rank, mix0.d_state, mix0.d_model
        Dl = D // G
        ncols = R + 2 * N
        es = torch.empty((), dtype=act_dtype).element_size()
        Lp_max = max(gh, gw)
        need = (Lt * 2 * D * es + 2 * Lp_max * ((ncols + 3) // 4 * 4) * 4 + L * 2 * 4 + L * Dl * es + 4 * ((dm + 3) // 4 * 4)
                + 8 * 1024)
        old = self._state.get("pb")
        pb = old if (old is not None and old.nbytes >= need + old.header and self._state.get("dt") == act_dtype) else \
            PeerBuffer(need, device, self.group)
        pb._cursor = pb.header
        st = {"pb": pb, "L": L, "Lt": Lt, "D": D, "Dl": Dl, "R": R, "N": N, "dm": dm, "ncols": ncols, "es": es}
        st["xz_tok"], st["off_xz"] = pb.carve((Lt, 2 * D), act_dtype)
        st["ncp"] = (ncols + 3) // 4 * 4                       # fp32 partial rows padded to a float4
        st["xpart"], st["off_xp"] = pb.carve((2, Lp_max, st["ncp"]), torch.float32)
        st["stats"], st["off_st"] = pb.carve((1, L, 2), torch.float32)
        st["y_ch"], st["off_y"] = pb.carve((1, L, Dl), act_dtype)
        st["msum"], st["off_ms"] = pb.carve(((dm + 3) // 4 * 4,), torch.float32)
        st["xpart"].zero_(); st["msum"].zero_()
        st["layers"] = [shard_mixer_params(blk.mixer, r, G, act_dtype) for blk in m.layers]
        with torch.no_grad():
            for blk, sp in zip(m.layers, st["layers"]):
                sp["in_w_full"] = blk.mixer.in_proj.weight.to(act_dtype).contiguous()
                sp["in_b_full"] = None if blk.mixer.in_proj.bias is None else blk.mixer.in_proj.bias.to(act_dtype)
                # x_proj weights padded to ncp columns so the partial product lands float4-aligned
                xw = sp["x_w_t"].float()
                sp["x_w_t32"] = torch.nn.functional.pad(xw, (0, st["ncp"] - ncols)).contiguous()
                mx = blk.mixer
                sp["ln_w_full"] = mx.layernorm.weight.float().contiguous() if mx.use_norm_after_ssm else None
                sp["ln_b_full"] = mx.layernorm.bias.float().contiguous() if mx.use_norm_after_ssm else None
                sp["ln_w_one"] = torch.ones(Dl, device=device, dtype=torch.float32)   # selects fv_gate_fwd's pre-norm mode
        self._state = {"k": key, "v": st, "pb": pb, "dt": act_dtype}
        torch.cuda.synchronize(device)
        dist.barrier(self.group)
        return st

    # ---- one mixer on the token shard ------------------------------------------------------------------------
    def _mixer(self, blk, sp, st, h_tok, rotated):
        from .mixer import linear as _linear

        m, pb, G, r = blk.mixer, st["pb"], self.world, self.rank
        L, Lt, D, Dl, es = st["L"], st["Lt"], st["D"], st["Dl"], st["es"]
        act = h_tok.dtype
        geom = m.geometry(rotated)
        Lp = geom.Lp
        # in_proj on my tokens, all 2 d_inner columns -> symmetric xz_tok
        h2 = h_tok.reshape(Lt, -1)
        xz_tok = st["xz_tok"]
        if (act == torch.bfloat16 and sp["in_b_full"] is None and ops.gemm_supported(Lt, 2 * D, h2.shape[1])
                and h2.stride(0) % 8 == 0 and h2.data_ptr() % 16 == 0):
            ops.gemm_bf16_tn(h2, sp["in_w_full"], out=xz_tok)
        else:
            xz_tok.copy_(F.linear(h2, sp["in_w_full"], sp["in_b_full"]))
        # A: all-to-all token -> channel of the x half only: my channel columns of every rank's tokens.  z never travels --
        # the gate is applied on the token side, where this rank already holds z for all channels of its tokens.
        x_ch = torch.empty((1, L, Dl), device=h_tok.device, dtype=act)
        lo = r * Dl
        src_off, dst_off = [], []
        for q in range(G):
            src_off += [st["off_xz"] + lo * es, 0]
            dst_off += [q * Lt * Dl * es, 0]
        pb.copy2d(1, Lt, Dl * es, src_off, 2 * D * es, dst_off, Dl * es, x_ch)
        x = x_ch
        u = ops.conv_pool_fwd(x, geom, sp["conv_w"], sp["conv_b"], float(m.scaling_factor), m.collapse_method)
        # B: x_proj partial over my channels -> symmetric buffer -> rank-ordered sum on every rank
        xpart = st["xpart"][:, :Lp]
        torch.bmm(u.view(2, Lp, Dl).float(), sp["x_w_t32"], out=xpart) if xpart.is_contiguous() else \
            xpart.copy_(torch.bmm(u.view(2, Lp, Dl).float(), sp["x_w_t32"]))
        xdbl = torch.empty((2, st["xpart"].shape[1], st["ncp"]), device=h_tok.device, dtype=act)
        n_xp = st["xpart"].numel()
        if act == torch.float32:
            pb.sum_f32(st["off_xp"], n_xp, out32=xdbl)
        else:
            pb.sum_f32(st["off_xp"], n_xp, out16=xdbl)
        xdbl = xdbl[:, :Lp]
        if not xdbl.is_contiguous():
            xdbl = xdbl.contiguous()
        s = ops.scan_fwd(u, xdbl, geom, m.dt_rank, m.d_state, sp["dt_w"], sp["dt_b"], sp["A_log"], a_is_log=True)
        eps = m.layernorm.eps if m.use_norm_after_ssm else 1e-5
        # channel side ends with the PRE-norm merged value v = (s_f + s_b + D_f xc_f + D_b xc_b) / 2 (fv_gate_fwd's
        # statistics mode writes exactly that; its per-shard sums are not needed any more)
        v_ch = st["y_ch"]
        ops.gate_fwd(x, x, s, geom, sp["conv_w"], sp["conv_b"], sp["D"], sp["ln_w_one"], None, eps, out=v_ch,
                     stats=st["stats"])
        # D: all-to-all channel -> token: all d_inner channels of my tokens
        v_tok = torch.empty((Lt, D), device=h_tok.device, dtype=act)
        src_off, dst_off = [], []
        for q in range(G):
            src_off += [st["off_y"] + r * Lt * Dl * es, 0]
            dst_off += [q * Dl * es, 0]
        pb.copy2d(1, Lt, Dl * es, src_off, Dl * es, dst_off, D * es, v_tok)
        # token side: LayerNorm over d_inner is local now; gate with the z this rank computed for its own tokens
        y_tok = ops.ln_gate_fwd(v_tok, xz_tok[:, D:], sp["ln_w_full"], sp["ln_b_full"], eps)
        out = _linear(y_tok, sp["out_w"], sp["out_b"]).view(1, Lt, -1)
        if m.init_layer_scale is not None:
            out = out * m.gamma
        return out

    @torch.no_grad()
    def forward(self, imgs):
        m, G, r = self.model, self.world, self.rank
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else imgs.dtype
        if imgs.shape[0] != 1:
            raise ValueError("HybridShardedVisionMamba shards ONE image over the ranks (batch 1)")
        if not hybrid_supported(m, G, imgs.shape[-2:]):
            raise ValueError("token bands must be whole token rows and d_inner / world a multiple of 8")
        st = self._prepare(act_dtype, imgs.device)
        gh, gw = m.token_size
        Lt, L = st["Lt"], st["L"]
        rows = Lt // gw
        ps = m.patch_size
        band = imgs[:, :, r * rows * ps:(r + 1) * rows * ps, :].contiguous()
        x = m.patch_embed(band)                                             # (1, Lt, dm): my band of token rows
        x = x + m.pos_embed[:, r * Lt:(r + 1) * Lt].to(x.dtype)
        residual, hidden = None, x
        for i, blk in enumerate(m.layers):
            is_rms = isinstance(blk.norm, RMSNorm)
            hidden, residual = layer_norm_fn(hidden, blk.norm.weight, blk.norm.bias, residual=residual, prenorm=True,
                                             residual_in_fp32=blk.residual_in_fp32, eps=blk.norm.eps, is_rms_norm=is_rms)
            rotated = blk.rotate_every_block is True and blk.layer_idx % 2 != 0
            hidden = self._mixer(blk, st["layers"][i], st, hidden.to(act_dtype), rotated)
        hs = layer_norm_fn(hidden, m.norm_f.weight, m.norm_f.bias, eps=m.norm_f.eps, residual=residual, prenorm=False,
                           residual_in_fp32=m.residual_in_fp32, is_rms_norm=isinstance(m.norm_f, RMSNorm))
        # mean over ALL tokens: my partial sum -> symmetric buffer -> peer sum
        dm = st["dm"]
        st["msum"][:dm].copy_(hs.float().sum(dim=1).reshape(-1) / float(L))
        feat32 = torch.empty_like(st["msum"])
        st["pb"].sum_f32(st["off_ms"], st["msum"].numel(), out32=feat32)
        feat = feat32[:dm].to(hs.dtype).view(1, dm)
        if m.num_classes > 0:
            feat = F.linear(feat, m.head.weight.to(feat.dtype), m.head.bias.to(feat.dtype))
        return feat


def shard_model_hybrid(model, group=None):
    """-> a module with the model's parameters whose forward runs one image over the process group (peer memory)."""
    return HybridShardedVisionMamba(model, group)
