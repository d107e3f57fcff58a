"""FastMaskVim encoder blocks -- host-side mirror of the reference ``models/mae/models_mamba_faster_mae_vimdecoder_v2.py``:
``Block_masked`` (:279-402), ``create_block_masked`` (:405-466), ``random_masking`` (:740-774) and the encoder half of
``MaskedAutoencoderViM`` (``forward_encoder`` :776-819).  Same constructor keywords, ``forward`` signatures and parameter
names.  The MAE decoder is a stack of plain bidirectional Vim blocks (``mamba_simple.Mamba``), the reference's baseline
architecture, and is not part of the FastVim hot path.

The encoder sees only the kept tokens of a masked image.  ``Block_masked`` carries their ORIGINAL token ids
(``ids_keep``); on odd layers it maps the ids through the (h, w) -> (w, h) rotation, re-sorts the tokens by rotated id,
runs the mixer built with the swapped ``token_size`` and restores the order (:376-396).  Add + norm is the CUDA kernel of
``fastvim_b200.norm``; the mixer is ``fastvim_b200.mixer_masked.Mamba_masked``.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .mixer_masked import Mamba_masked
from .norm import RMSNorm, layer_norm_fn
from .vision import PatchEmbed, _init_weights, _segm_init_weights


class Block_masked(nn.Module):
    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False,
                 rotate_every_block=True, layer_idx=None, token_size=None):
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.mixer, self.norm = mixer_cls(dim), norm_cls(dim)
        self.rotate_every_block, self.layer_idx, self.token_size = rotate_every_block, layer_idx, token_size
        self.rotate_indices = self.compute_rotate_indices(token_size[0], token_size[1])

    @staticmethod
    def compute_rotate_indices(H, W):
        """rotate_indices[i*W + j] = j*H + i (:320-328), as one index expression."""
        idx = torch.arange(H * W)
        return (idx % W) * H + idx // W

    def forward(self, hidden_states: Tensor, residual: Optional[Tensor] = None, ids_keep=None, inference_params=None):
        hidden_states, residual = layer_norm_fn(
            hidden_states, self.norm.weight, self.norm.bias, residual=residual, prenorm=True,
            residual_in_fp32=self.residual_in_fp32, eps=self.norm.eps, is_rms_norm=isinstance(self.norm, RMSNorm))
        odd = self.rotate_every_block is True and self.layer_idx % 2 != 0
        if odd:                                                                   # :376-386
            ids_keep = self.rotate_indices.to(hidden_states.device)[ids_keep]
            rotated_ids = torch.argsort(ids_keep)
            ids_keep = torch.gather(ids_keep, 1, rotated_ids)
            rotated_ids = rotated_ids.unsqueeze(-1)
            hidden_states = torch.gather(hidden_states, 1, rotated_ids.repeat(1, 1, hidden_states.shape[-1]))
        hidden_states = self.mixer(hidden_states, ids_keep, inference_params=inference_params)
        if odd:                                                                   # :392-396
            hidden_states = torch.gather(hidden_states, 1,
                                         torch.argsort(rotated_ids, -2).repeat(1, 1, hidden_states.shape[-1]))
        return hidden_states, residual


def create_block_masked(d_model, ssm_cfg=None, norm_epsilon=1e-5, rms_norm=False, residual_in_fp32=False,
                        fused_add_norm=False, layer_idx=None, device=None, dtype=None, init_layer_scale=None,
                        scanpath_type="rowwise", use_norm_after_ssm=True, rotate_every_block=True, collapse_method="mean",
                        token_size=None):
    ssm_cfg = ssm_cfg or {}
    factory_kwargs = {"device": device, "dtype": dtype}
    odd = rotate_every_block is True and layer_idx % 2 != 0
    mixer_cls = partial(Mamba_masked, layer_idx=layer_idx, init_layer_scale=init_layer_scale, scanpath_type=scanpath_type,
                        use_norm_after_ssm=use_norm_after_ssm,
                        token_size=[token_size[1], token_size[0]] if odd else token_size,   # :423-446
                        collapse_method=collapse_method, **ssm_cfg, **factory_kwargs)
    norm_cls = partial(nn.LayerNorm if not rms_norm else RMSNorm, eps=norm_epsilon, **factory_kwargs)
    block = Block_masked(d_model, mixer_cls, norm_cls=norm_cls, fused_add_norm=fused_add_norm,
                         residual_in_fp32=residual_in_fp32, rotate_every_block=rotate_every_block, layer_idx=layer_idx,
                         token_size=token_size)
    block.layer_idx = layer_idx
    return block


def random_masking(x: Tensor, mask_ratio: float):
    """Per-sample random masking with the kept ids SORTED, "for Mamba since sequential" (:740-774).
    x (N, L, D) -> x_masked (N, len_keep, D), mask (N, L) [0 keep, 1 remove], ids_restore, ids_keep."""
    N, L, D = x.shape
    len_keep = int(L * (1 - mask_ratio))
    noise = torch.rand(N, L, device=x.device)
    ids_shuffle = torch.argsort(noise, dim=1)
    ids_shuffle[:, :len_keep] = ids_shuffle[:, :len_keep].sort().values
    ids_shuffle = ids_shuffle.contiguous()
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    ids_keep = ids_shuffle[:, :len_keep]
    x_masked = torch.gather(x, dim=1, index=ids_keep.unsqueeze(-1).repeat(1, 1, D))
    mask = torch.ones([N, L], device=x.device)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, dim=1, index=ids_restore)
    return x_masked, mask, ids_restore, ids_keep


class MaskedEncoder(nn.Module):
    """The encoder half of the reference's ``MaskedAutoencoderViM`` (ctor :513-640, ``forward_encoder`` :776-819):
    patch embedding + positional embedding + random masking + ``depth`` ``Block_masked`` + final norm.  Parameter names
    (``patch_embed``, ``pos_embed``, ``layers``, ``norm_f``) follow the reference, so its encoder weights load with
    ``strict=False`` (the decoder's are ignored)."""

    def __init__(self, img_size=224, patch_size=16, depth=24, embed_dim=192, channels=3, norm_epsilon=1e-5,
                 rms_norm=True, fused_add_norm=True, residual_in_fp32=True, ssm_cfg=None, init_layer_scale=None,
                 scanpath_type="rowwise", use_norm_after_ssm=True, rotate_every_block=True, collapse_method="mean",
                 initializer_cfg=None, device=None, dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=channels, embed_dim=embed_dim,
                                      scanpath_type=scanpath_type)
        self.token_size = self.patch_embed.grid_size
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
        self.layers = nn.ModuleList([
            create_block_masked(embed_dim, ssm_cfg=ssm_cfg, norm_epsilon=norm_epsilon, rms_norm=rms_norm,
                                residual_in_fp32=residual_in_fp32, fused_add_norm=fused_add_norm, layer_idx=i,
                                init_layer_scale=init_layer_scale, scanpath_type=scanpath_type,
                                use_norm_after_ssm=use_norm_after_ssm, rotate_every_block=rotate_every_block,
                                collapse_method=collapse_method, token_size=self.token_size, **factory_kwargs)
            for i in range(depth)])
        self.norm_f = (nn.LayerNorm if not rms_norm else RMSNorm)(embed_dim, eps=norm_epsilon, **factory_kwargs)
        self.patch_embed.apply(_segm_init_weights)
        nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-2.0, b=2.0)
        self.apply(partial(_init_weights, n_layer=depth, **(initializer_cfg if initializer_cfg is not None else {})))

    def forward(self, x, mask_ratio, inference_params=None, ids_keep=None):
        """-> (latent (N, len_keep, E), mask, ids_restore).  ``ids_keep`` (sorted, (N, len_keep)) overrides the random
        draw (tests); then mask / ids_restore are returned as None."""
        act_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        x = self.patch_embed(x.to(act_dtype))
        x = x + self.pos_embed.to(x.dtype)
        if ids_keep is None:
            x, mask, ids_restore, ids_keep = random_masking(x, mask_ratio)
        else:
            mask = ids_restore = None
            x = torch.gather(x, dim=1, index=ids_keep.unsqueeze(-1).repeat(1, 1, x.shape[-1]))
        residual, hidden_states = None, x
        for layer in self.layers:
            hidden_states, residual = layer(hidden_states, residual, ids_keep.clone(), inference_params=inference_params)
        hidden_states = layer_norm_fn(hidden_states, self.norm_f.weight, self.norm_f.bias, eps=self.norm_f.eps,
                                      residual=residual, prenorm=False, residual_in_fp32=self.residual_in_fp32,
                                      is_rms_norm=isinstance(self.norm_f, RMSNorm))
        return hidden_states, mask, ids_restore
