"""FastChannelVim backbone -- host-side mirror of the reference
``models/channel_wise_tokenization/models_channel_mamba_faster.py`` module interface.

``PatchEmbedPerChannel`` (:22-203), ``Block`` (:206-336), ``create_block`` (:339-408), ``VisionMamba`` (:458-683)
and the S/16 factory (:686-706) keep the reference's constructor keywords, ``forward`` signatures and parameter
names, so reference checkpoints load unchanged.  Every image channel is tokenised on its own (one shared 16 x 16
patch projection + a per-channel embedding), giving ``rows * cols * tokens_per_patch`` tokens, and the scan runs
over the sequence pooled along the patch columns (``mixer_channel.Mamba``: Channel-First ``(rows, cols, tpp)`` or
Spatial-First ``(tpp*rows, cols, 1)`` geometry on the same kernels as FastVim).

What differs from the reference underneath: add + RMSNorm is the CUDA kernel behind ``fastvim_b200.norm`` (no
Triton), the patch projection is a GEMM over unfolded patches (tcgen05 when the shape qualifies), the mixer is
``fastvim_b200.mixer_channel.Mamba``.  The odd-layer token transposition (:304-331) is materialised as in the
reference (the Spatial-First transposition is not a strided walk of the ``(outer, pool, inner)`` geometry).
"""
from __future__ import annotations

import random
from functools import partial
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import autograd as fv_autograd
from . import ops
from .mixer import linear as _linear
from .mixer_channel import Mamba
from .norm import RMSNorm, layer_norm_fn
from .vision import _init_weights, _segm_init_weights, _to_2tuple


class PatchEmbedPerChannel(nn.Module):
    """Per-channel patch embedding with hierarchical channel sampling (reference :22-203)."""

    def __init__(self, img_size=224, patch_size=16, stride=16, in_chans=8, embed_dim=768, hcs=True,
                 scan_order="Channel-First", sort_channels=True, flatten=True, scanpath_type="rowwise"):
        super().__init__()
        self.img_size, self.patch_size = _to_2tuple(img_size), _to_2tuple(patch_size)
        gh, gw = self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1]
        self.grid_size = (gw, gh) if scanpath_type == "colwise" else (gh, gw)
        self.num_patches = gh * gw
        self.scanpath_type, self.flatten = scanpath_type, flatten
        self.hcs, self.scan_order, self.sort_channels = hcs, scan_order, sort_channels
        if stride != patch_size:
            raise NotImplementedError("fastvim_b200: non-overlapping patches (stride == patch_size) only")
        # all channels share the same filter weights: the channel axis is the conv's depth axis (:113-121)
        self.proj = nn.Conv3d(1, embed_dim, kernel_size=(1, patch_size, patch_size), stride=(1, stride, stride))
        self.channel_embed = nn.Embedding(in_chans, embed_dim)

    def forward(self, x: Tensor, input_channel_order: Optional[Tensor] = None):
        B, num_channels, h, w = x.shape
        if input_channel_order is None:
            ids = torch.arange(0, num_channels, device=x.device).repeat(B, 1)
        else:
            ids = input_channel_order
        channel_embed = self.channel_embed(ids)                                  # (B, C, E)
        if self.training and self.hcs:                                           # :164-178
            C_new = random.randint(1, num_channels)
            channels = random.sample(range(num_channels), k=C_new)
            if self.sort_channels is True:
                channels.sort()
            num_channels = C_new
            x = x[:, channels, :, :]
            channel_embed = channel_embed[:, channels, :]
        else:
            channels = random.sample(range(num_channels), k=num_channels)
            channels.sort()
        # shared non-overlapping projection == GEMM over unfolded patches: (B*C*gh*gw, p*p) x (p*p, E)
        p0, p1 = self.patch_size
        gh, gw = h // p0, w // p1
        wt = self.proj.weight
        autocast = torch.is_autocast_enabled("cuda")
        act_dtype = torch.get_autocast_dtype("cuda") if autocast else x.dtype
        no_grad = not (torch.is_grad_enabled() and (wt.requires_grad or x.requires_grad))
        native_in = (act_dtype == torch.bfloat16 and p0 == p1 and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16)
                     and ops.patchify_supported(x.contiguous(), p0))
        if (native_in and not no_grad and fv_autograd.NATIVE_PATCH_TRAIN and not x.requires_grad
                and ops.gemm_supported(B * num_channels * gh * gw, wt.shape[0], p0 * p1) and wt.shape[0] % 8 == 0):
            # training: fv_patchify + tcgen05 GEMM under autograd, weight gradient on the general tcgen05 GEMM
            out = fv_autograd.PatchEmbedFn.apply(x.contiguous(), wt, self.proj.bias, p0, True)
        else:
            if native_in and no_grad:
                cols = ops.patchify(x.contiguous(), p0, per_channel=True)     # one pass image -> bf16 patches (fv_patchify)
            else:
                cols = x.reshape(B, num_channels, gh, p0, gw, p1).permute(0, 1, 2, 4, 3, 5).reshape(-1, p0 * p1).to(act_dtype)
            wmat = wt.reshape(wt.shape[0], -1).to(cols.dtype)
            bias = None if self.proj.bias is None else self.proj.bias.to(cols.dtype)
            if cols.is_cuda and no_grad:
                out = _linear(cols, wmat, bias)
            else:
                out = F.linear(cols, wmat, bias)
        out = out.reshape(B, num_channels, gh, gw, -1)                            # (B, C, H/ps, W/ps, E)
        out = out + channel_embed.to(out.dtype)[:, :, None, None, :]              # channel-specific offsets (:184)
        if self.scanpath_type == "colwise":
            out = out.transpose(2, 3)
        if self.scan_order == "Channel-First":
            out = out.permute(0, 2, 3, 1, 4)                                      # (B, H/ps, W/ps, C, E)
        if self.flatten:
            out = out.reshape(B, -1, out.shape[-1])                               # (B, L, E)
        else:                                                                     # reference layout (B, E, ...)
            out = out.permute(0, 4, 1, 2, 3)
        return out, num_channels, h, w, channels


class Block(nn.Module):
    """Add -> norm -> channel mixer (reference :206-336).  ``forward`` returns (hidden_states, residual)."""

    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False,
                 drop_path=0.0, rotate_every_block=True, layer_idx=None, token_size=None, scan_order=None,
                 max_tokens_per_patch=None):
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.mixer, self.norm = mixer_cls(dim), norm_cls(dim)
        self.rotate_every_block, self.layer_idx, self.token_size = rotate_every_block, layer_idx, token_size
        self.scan_order = scan_order
        self.drop_path_rate = drop_path

    def drop_path(self, x):
        if self.drop_path_rate == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_path_rate
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep

    def forward(self, hidden_states: Tensor, tokens_per_patch: int, residual: Optional[Tensor] = None,
                inference_params=None):
        hidden_states, residual = layer_norm_fn(
            hidden_states if residual is None else self.drop_path(hidden_states), self.norm.weight,
            self.norm.bias, residual=residual, prenorm=True, residual_in_fp32=self.residual_in_fp32,
            eps=self.norm.eps, is_rms_norm=isinstance(self.norm, RMSNorm))
        odd = self.rotate_every_block is True and self.layer_idx % 2 != 0
        t0, t1 = self.token_size
        if odd:                                                                   # :304-316
            B, M, _ = hidden_states.shape
            if self.scan_order == "Spatial-First":
                hidden_states = hidden_states.reshape(B, tokens_per_patch, t0, t1, -1).transpose(2, 3).reshape(B, M, -1)
            elif self.scan_order == "Channel-First":
                hidden_states = hidden_states.reshape(B, t0, t1, tokens_per_patch, -1).transpose(1, 2).reshape(B, M, -1)
        hidden_states = self.mixer(hidden_states, tokens_per_patch, inference_params=inference_params)
        if odd:                                                                   # :322-331
            if self.scan_order == "Spatial-First":
                hidden_states = hidden_states.reshape(B, tokens_per_patch, t1, t0, -1).transpose(2, 3).reshape(B, M, -1)
            elif self.scan_order == "Channel-First":
                hidden_states = hidden_states.reshape(B, t1, t0, tokens_per_patch, -1).transpose(1, 2).reshape(B, M, -1)
        return hidden_states, residual


def create_block(d_model, ssm_cfg=None, norm_epsilon=1e-5, drop_path=0.0, rms_norm=False, residual_in_fp32=False,
                 fused_add_norm=False, layer_idx=None, device=None, dtype=None, init_layer_scale=None,
                 scanpath_type="rowwise", use_norm_after_ssm=True, rotate_every_block=True, collapse_method="mean",
                 token_size=None, scan_order=None, max_tokens_per_patch=None):
    ssm_cfg = ssm_cfg or {}
    factory_kwargs = {"device": device, "dtype": dtype}
    odd = rotate_every_block is True and layer_idx % 2 != 0
    mixer_cls = partial(Mamba, layer_idx=layer_idx, init_layer_scale=init_layer_scale, scanpath_type=scanpath_type,
                        use_norm_after_ssm=use_norm_after_ssm,
                        token_size=[token_size[1], token_size[0]] if odd else token_size,   # reference :363-388
                        collapse_method=collapse_method, scan_order=scan_order, **ssm_cfg, **factory_kwargs)
    norm_cls = partial(nn.LayerNorm if not rms_norm else RMSNorm, eps=norm_epsilon, **factory_kwargs)
    block = Block(d_model, mixer_cls, norm_cls=norm_cls, drop_path=drop_path, fused_add_norm=fused_add_norm,
                  residual_in_fp32=residual_in_fp32, rotate_every_block=rotate_every_block, layer_idx=layer_idx,
                  token_size=token_size, scan_order=scan_order, max_tokens_per_patch=max_tokens_per_patch)
    block.layer_idx = layer_idx
    return block


class VisionMamba(nn.Module):
    """FastChannelVim classifier (reference :458-683)."""

    def __init__(self, img_size=224, patch_size=16, stride=16, depth=24, embed_dim=192, channels=3, num_classes=1000,
                 ssm_cfg=None, drop_rate=0.0, drop_path_rate=0.1, norm_epsilon: float = 1e-5, rms_norm: bool = False,
                 initializer_cfg=None, fused_add_norm=False, residual_in_fp32=False, device=None, dtype=None,
                 final_pool_type="mean", if_abs_pos_embed=True, init_layer_scale=None, scan_order="Channel-First",
                 hcs=True, sort_channels=True, scanpath_type="rowwise", use_norm_after_ssm=True,
                 rotate_every_block=True, collapse_method="mean", **kwargs):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.final_pool_type, self.if_abs_pos_embed = final_pool_type, if_abs_pos_embed
        self.rotate_every_block, self.channels = rotate_every_block, channels
        self.num_classes = num_classes
        self.d_model = self.num_features = self.embed_dim = embed_dim
        self.scan_order, self.patch_size = scan_order, patch_size
        self.patch_embed = PatchEmbedPerChannel(img_size=img_size, patch_size=patch_size, stride=stride,
                                                in_chans=channels, embed_dim=embed_dim, hcs=hcs, scan_order=scan_order,
                                                sort_channels=sort_channels, scanpath_type=scanpath_type)
        self.num_patches = self.patch_embed.num_patches
        self.token_size = self.patch_embed.grid_size
        if if_abs_pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, self.embed_dim))
            self.pos_drop = nn.Dropout(p=drop_rate)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        inter_dpr = [0.0] + dpr
        self.drop_path_rate = drop_path_rate
        self.layers = nn.ModuleList([
            create_block(embed_dim, ssm_cfg=ssm_cfg, norm_epsilon=norm_epsilon, rms_norm=rms_norm,
                         residual_in_fp32=residual_in_fp32, fused_add_norm=fused_add_norm, layer_idx=i,
                         drop_path=inter_dpr[i], init_layer_scale=init_layer_scale, scanpath_type=scanpath_type,
                         use_norm_after_ssm=use_norm_after_ssm, rotate_every_block=rotate_every_block,
                         collapse_method=collapse_method, token_size=self.token_size, scan_order=self.scan_order,
                         max_tokens_per_patch=self.channels, **factory_kwargs)
            for i in range(depth)])
        self.norm_f = (nn.LayerNorm if not rms_norm else RMSNorm)(embed_dim, eps=norm_epsilon, **factory_kwargs)
        self.patch_embed.apply(_segm_init_weights)   # names Conv2d only: the Conv3d keeps torch's default (:443-455)
        self.head.apply(_segm_init_weights)
        if if_abs_pos_embed:
            nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-2.0, b=2.0)
        self.apply(partial(_init_weights, n_layer=depth, **(initializer_cfg if initializer_cfg is not None else {})))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "pos_embed_obj", "cls_token", "dist_token", "cls_token_head", "cls_token_tail"}

    def _drop_path(self, x):
        if self.drop_path_rate == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_path_rate
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep

    def forward_features(self, x, inference_params=None):
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        if x.dtype == torch.uint8:
            x = x.float()
        x, tokens_per_patch, h, w, channels_list = self.patch_embed(x)
        if self.if_abs_pos_embed:                                                 # :621-630
            pe = self.pos_embed.to(x.dtype)
            if self.scan_order == "Spatial-First":
                x = x + pe.expand(tokens_per_patch, -1, self.embed_dim).reshape(1, -1, self.embed_dim)
            elif self.scan_order == "Channel-First":
                x = x + torch.repeat_interleave(pe, tokens_per_patch, 1)
            x = self.pos_drop(x)
        residual, hidden_states = None, x
        for layer in self.layers:
            hidden_states, residual = layer(hidden_states, tokens_per_patch, residual, inference_params=inference_params)
        hidden_states = layer_norm_fn(self._drop_path(hidden_states), self.norm_f.weight, self.norm_f.bias,
                                      eps=self.norm_f.eps, residual=residual, prenorm=False,
                                      residual_in_fp32=self.residual_in_fp32,
                                      is_rms_norm=isinstance(self.norm_f, RMSNorm))
        if self.final_pool_type == "none":
            return hidden_states[:, -1, :]
        if self.final_pool_type == "mean":
            return hidden_states.mean(dim=1)
        if self.final_pool_type in ("max", "all"):
            return hidden_states
        raise NotImplementedError

    def forward(self, x, return_features=False, inference_params=None):
        x = self.forward_features(x, inference_params)
        if return_features:
            return x
        x = F.linear(x, self.head.weight.to(x.dtype), self.head.bias.to(x.dtype)) if self.num_classes > 0 else x
        if self.final_pool_type == "max":
            x = x.max(dim=1)[0]
        return x


def channelvim_small_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2(
        pretrained=False, patch_size=16, stride=16, if_abs_pos_embed=True, **kwargs):
    """FastChannelVim-S/16 (reference :686-706)."""
    if pretrained:
        raise NotImplementedError("no pretrained FastChannelVim checkpoints are published (reference url: 'to.do')")
    return VisionMamba(patch_size=patch_size, stride=stride, if_abs_pos_embed=if_abs_pos_embed, embed_dim=384,
                       depth=24, rms_norm=True, residual_in_fp32=True, fused_add_norm=True, **kwargs)
