"""FastVim backbone -- host-side mirror of the reference ``models/fastvim.py`` module interface.

``PatchEmbed`` (:25-103), ``Block`` (:106-217), ``create_block`` (:220-291), ``VisionMamba``
(:342-557) and the T/S/B factories (:695-819) keep the reference's constructor keywords,
``forward`` signatures and parameter names, so reference checkpoints load unchanged.  What
differs is what runs underneath:

* add + RMSNorm is the CUDA kernel behind ``fastvim_b200.norm`` (the reference uses Triton);
* the odd-layer token rotation (:192-210) is NOT materialised: the residual stream stays in
  row-major token order for the whole network and the mixer is told ``rotated=True``, which
  only changes the row strides its kernels walk;
* the mixer is ``fastvim_b200.mixer.Mamba``.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import autograd as fv_autograd
from . import ops
from .mixer import Mamba
from .mixer import linear as _linear
from .norm import RMSNorm, layer_norm_fn, rms_norm_fn


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class PatchEmbed(nn.Module):
    """2D image to patch embedding (reference :25-103)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True, strict_img_size=True, dynamic_img_pad=False, scanpath_type="rowwise"):
        super().__init__()
        self.img_size, self.patch_size = _to_2tuple(img_size), _to_2tuple(patch_size)
        gh, gw = self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1]
        self.grid_size = (gw, gh) if scanpath_type == "colwise" else (gh, gw)
        self.num_patches = gh * gw
        self.scanpath_type, self.flatten = scanpath_type, flatten
        self.strict_img_size, self.dynamic_img_pad = strict_img_size, dynamic_img_pad
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    # ---- uint8 pipelines: (x / 255 - mean) / std folded into the projection ---------------------------------
    def set_input_normalization(self, mean=None, std=None):
        """Declare how uint8 images map to the float images the model was trained on: ``(x / 255 - mean) / std`` per
        input channel (``None`` = plain ``x.float()``).  The fast path folds this affine map into the projection's weights
        and bias, so a uint8 image crosses PCIe and HBM at one byte per pixel and is converted exactly (0..255 are
        bf16-representable)."""
        C = self.proj.weight.shape[1]
        if mean is None and std is None:
            self._in_norm = None
        else:
            mean = torch.as_tensor(0.0 if mean is None else mean, dtype=torch.float32).reshape(-1).expand(C).clone()
            std = torch.as_tensor(1.0 if std is None else std, dtype=torch.float32).reshape(-1).expand(C).clone()
            self._in_norm = (mean, std)
        self._fold_cache = None

    def _uint8_to_float(self, x):
        x = x.float()
        norm = getattr(self, "_in_norm", None)
        if norm is not None:
            mean, std = (t.to(x.device)[None, :, None, None] for t in norm)
            x = (x / 255.0 - mean) / std
        return x

    def _folded(self, dtype):
        """Projection weights (E, C*p*p) / fp32 bias for uint8 input: W' = W / (255 std_c), b' = b - sum W mean_c / std_c."""
        w, b = self.proj.weight, self.proj.bias
        norm = getattr(self, "_in_norm", None)
        key = (w.data_ptr(), w._version, None if b is None else b._version, dtype, id(norm))
        hit = getattr(self, "_fold_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            w32 = w.float()
            b32 = torch.zeros(w.shape[0], device=w.device) if b is None else b.float()
            if norm is not None:
                mean, std = (t.to(w.device) for t in norm)
                b32 = b32 - (w32 * (mean / std)[None, :, None, None]).sum(dim=(1, 2, 3))
                w32 = w32 / (255.0 * std)[None, :, None, None]
            wmat = w32.reshape(w.shape[0], -1).to(dtype).contiguous()
        self._fold_cache = (key, wmat, b32.contiguous())
        return wmat, b32

    def forward(self, x):
        B, C, H, W = x.shape
        if self.strict_img_size:
            assert H == self.img_size[0] and W == self.img_size[1], "input size doesn't match model"
        if self.dynamic_img_pad:
            ph = (self.patch_size[0] - H % self.patch_size[0]) % self.patch_size[0]
            pw = (self.patch_size[1] - W % self.patch_size[1]) % self.patch_size[1]
            if ph or pw:
                if x.dtype == torch.uint8:
                    x = self._uint8_to_float(x)   # zero padding is applied to the normalised image
                x = F.pad(x, (0, pw, 0, ph))
        # Non-overlapping conv == GEMM over unfolded patches: (B*L, C*p*p) x (C*p*p, E).
        p0, p1 = self.patch_size
        B, C, H, W = x.shape
        gh, gw = H // p0, W // p1
        w = self.proj.weight
        autocast = torch.is_autocast_enabled("cuda")
        act_dtype = torch.get_autocast_dtype("cuda") if autocast else (torch.float32 if x.dtype == torch.uint8 else x.dtype)
        no_grad = not (torch.is_grad_enabled() and (w.requires_grad or x.requires_grad))
        if (no_grad and act_dtype == torch.bfloat16 and p0 == p1 and x.is_cuda and ops.patchify_supported(x.contiguous(), p0)
                and ops.gemm_supported(B * gh * gw, w.shape[0], C * p0 * p1)):
            # inference fast path: ONE pass image -> bf16 patches (fv_patchify: fp32 / bf16 / uint8 in), then the tcgen05 GEMM
            cols = ops.patchify(x.contiguous(), p0)
            if x.dtype == torch.uint8:
                wmat, bias32 = self._folded(torch.bfloat16)
            else:
                wmat, bias32 = self._folded_plain()
            out = ops.gemm_bf16_tn(cols, wmat, bias=bias32)
        elif (not no_grad and fv_autograd.NATIVE_PATCH_TRAIN and not x.requires_grad and act_dtype == torch.bfloat16
              and p0 == p1 and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16)
              and ops.patchify_supported(x.contiguous(), p0) and ops.gemm_supported(B * gh * gw, w.shape[0], C * p0 * p1)
              and C * p0 * p1 % 8 == 0 and w.shape[0] % 8 == 0):
            # training: same two kernels under autograd, weight gradient on the general tcgen05 GEMM
            out = fv_autograd.PatchEmbedFn.apply(x.contiguous(), w, self.proj.bias, p0)
        else:
            if x.dtype == torch.uint8:
                x = self._uint8_to_float(x)
            cols = x.reshape(B, C, gh, p0, gw, p1).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * p0 * p1)
            cols = cols.to(act_dtype)
            wmat = w.reshape(w.shape[0], -1).to(cols.dtype)
            bias = None if self.proj.bias is None else self.proj.bias.to(cols.dtype)
            if cols.is_cuda and no_grad:
                out = _linear(cols, wmat, bias)     # tcgen05 GEMM (bf16) / cuBLAS
            else:
                out = F.linear(cols, wmat, bias)
        out = out.reshape(B, gh, gw, -1)
        if self.scanpath_type == "colwise":
            out = out.transpose(1, 2)
        if self.flatten:
            out = out.reshape(B, gh * gw, -1)
        else:
            out = out.permute(0, 3, 1, 2)
        return self.norm(out)

    def _folded_plain(self):
        w, b = self.proj.weight, self.proj.bias
        key = (w.data_ptr(), w._version, None if b is None else b._version)
        hit = getattr(self, "_plain_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            wmat = w.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
            # the reference's autocast conv adds the bias rounded to bf16; keep that rounding, in an fp32 container
            b32 = None if b is None else b.to(torch.bfloat16).float().contiguous()
        self._plain_cache = (key, wmat, b32)
        return wmat, b32


class Block(nn.Module):
    """Add -> norm -> mixer (reference :106-217).  ``forward`` returns (hidden_states, residual)."""

    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False,
                 residual_in_fp32=False, drop_path=0.0, rotate_every_block=True, layer_idx=None,
                 token_size=None):
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.mixer, self.norm = mixer_cls(dim), norm_cls(dim)
        self.rotate_every_block, self.layer_idx, self.token_size = rotate_every_block, layer_idx, token_size
        self.drop_path_rate = drop_path

    def drop_path(self, x):
        if self.drop_path_rate == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_path_rate
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep

    def forward(self, hidden_states: Tensor, residual: Optional[Tensor] = None, inference_params=None):
        is_rms = isinstance(self.norm, RMSNorm)
        hidden_states, residual = layer_norm_fn(
            hidden_states if residual is None else self.drop_path(hidden_states), self.norm.weight,
            self.norm.bias, residual=residual, prenorm=True, residual_in_fp32=self.residual_in_fp32,
            eps=self.norm.eps, is_rms_norm=is_rms)
        rotated = self.rotate_every_block is True and self.layer_idx % 2 != 0
        # reference :192-210 permutes tokens before and after the mixer; here the permutation is
        # an index map inside the mixer's kernels (see fastvim_b200.ops.Geometry.grid).
        hidden_states = self.mixer(hidden_states, inference_params=inference_params, rotated=rotated)
        return hidden_states, residual


def create_block(d_model, ssm_cfg=None, norm_epsilon=1e-5, drop_path=0.0, rms_norm=False,
                 residual_in_fp32=False, fused_add_norm=False, layer_idx=None, device=None, dtype=None,
                 init_layer_scale=None, scanpath_type="rowwise", use_norm_after_ssm=True,
                 rotate_every_block=True, collapse_method="mean", token_size=None,
                 use_our_selective_scan=False, scaling_factor=1):
    ssm_cfg = ssm_cfg or {}
    factory_kwargs = {"device": device, "dtype": dtype}
    odd = rotate_every_block is True and layer_idx % 2 != 0
    mixer_cls = partial(Mamba, layer_idx=layer_idx, init_layer_scale=init_layer_scale,
                        scanpath_type=scanpath_type, use_norm_after_ssm=use_norm_after_ssm,
                        token_size=[token_size[1], token_size[0]] if odd else token_size,  # reference :244-260
                        collapse_method=collapse_method, use_our_selective_scan=use_our_selective_scan,
                        scaling_factor=scaling_factor, **ssm_cfg, **factory_kwargs)
    norm_cls = partial(nn.LayerNorm if not rms_norm else RMSNorm, eps=norm_epsilon, **factory_kwargs)
    block = Block(d_model, mixer_cls, norm_cls=norm_cls, drop_path=drop_path, fused_add_norm=fused_add_norm,
                  residual_in_fp32=residual_in_fp32, rotate_every_block=rotate_every_block,
                  layer_idx=layer_idx, token_size=token_size)
    block.layer_idx = layer_idx
    return block


def _init_weights(module, n_layer, initializer_range=0.02, rescale_prenorm_residual=True,
                  n_residuals_per_layer=1):
    """Reference :295-324."""
    if isinstance(module, nn.Linear):
        if module.bias is not None and not getattr(module.bias, "_no_reinit", False):
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Embedding):
        nn.init.normal_(module.weight, std=initializer_range)
    if rescale_prenorm_residual:
        for name, p in module.named_parameters():
            if name in ["out_proj.weight", "fc2.weight"]:
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                with torch.no_grad():
                    p /= math.sqrt(n_residuals_per_layer * n_layer)


def _segm_init_weights(m):
    """Reference :327-339."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02, a=-2.0, b=2.0)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Conv2d):
        fan_in = nn.init._calculate_fan_in_and_fan_out(m.weight)[0]
        nn.init.trunc_normal_(m.weight, std=math.sqrt(1.0 / fan_in) / 0.87962566103423978)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d)):
        nn.init.zeros_(m.bias)
        nn.init.ones_(m.weight)


class VisionMamba(nn.Module):
    """Reference ``VisionMamba`` (:342-557) on the B200 kernels."""

    def __init__(self, img_size=224, patch_size=16, stride=16, depth=24, embed_dim=192, channels=3,
                 num_classes=1000, ssm_cfg=None, drop_rate=0.0, drop_path_rate=0.1,
                 norm_epsilon: float = 1e-5, rms_norm: bool = True, initializer_cfg=None,
                 fused_add_norm=False, residual_in_fp32=False, device=None, dtype=None,
                 final_pool_type="none", if_abs_pos_embed=True, init_layer_scale=None,
                 embed_layer=PatchEmbed, scanpath_type="rowwise", use_norm_after_ssm=True,
                 rotate_every_block=True, collapse_method="mean", use_our_selective_scan=False,
                 scaling_factor=1, **kwargs):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.final_pool_type, self.if_abs_pos_embed = final_pool_type, if_abs_pos_embed
        self.rotate_every_block = rotate_every_block
        self.num_classes = num_classes
        self.d_model = self.num_features = self.embed_dim = embed_dim
        self.patch_size = patch_size
        self.depth = depth
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=channels,
                                       embed_dim=embed_dim, strict_img_size=False, dynamic_img_pad=True,
                                       scanpath_type=scanpath_type)
        self.num_patches = self.patch_embed.num_patches
        self.token_size = self.patch_embed.grid_size
        if if_abs_pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, embed_dim))
            self.pos_drop = nn.Dropout(p=drop_rate)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        inter_dpr = [0.0] + dpr
        self.drop_path_rate = drop_path_rate
        self.layers = nn.ModuleList([
            create_block(embed_dim, ssm_cfg=ssm_cfg, norm_epsilon=norm_epsilon, rms_norm=rms_norm,
                         residual_in_fp32=residual_in_fp32, fused_add_norm=fused_add_norm, layer_idx=i,
                         drop_path=inter_dpr[i], init_layer_scale=init_layer_scale,
                         scanpath_type=scanpath_type, use_norm_after_ssm=use_norm_after_ssm,
                         rotate_every_block=rotate_every_block, collapse_method=collapse_method,
                         token_size=self.token_size, use_our_selective_scan=use_our_selective_scan,
                         scaling_factor=scaling_factor, **factory_kwargs)
            for i in range(depth)])
        self.norm_f = (nn.LayerNorm if not rms_norm else RMSNorm)(embed_dim, eps=norm_epsilon, **factory_kwargs)
        self.patch_embed.apply(_segm_init_weights)
        self.head.apply(_segm_init_weights)
        if if_abs_pos_embed:
            nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-2.0, b=2.0)
        self.apply(partial(_init_weights, n_layer=depth, **(initializer_cfg or {})))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed"}

    def set_input_normalization(self, mean=None, std=None):
        """uint8 images given to ``forward`` mean ``(x / 255 - mean) / std`` (see ``PatchEmbed.set_input_normalization``)."""
        self.patch_embed.set_input_normalization(mean, std)

    def forward_features(self, x, inference_params=None, out_indices=None):
        B, _, H, W = x.shape
        # images go to the patch embedding in their host dtype (fp32, bf16 or uint8): the unfold kernel casts on the fly
        x = self.patch_embed(x)
        if self.if_abs_pos_embed:
            gh, gw = math.ceil(H / self.patch_size), math.ceil(W / self.patch_size)
            if gh != self.token_size[0] or gw != self.token_size[1]:
                # the reference's resize branch cannot run (it calls a 5-argument staticmethod with 4,
                # models/fastvim.py:494-496 vs :646): build the model with img_size == input size.
                raise ValueError("build VisionMamba with img_size equal to the input size")
            x = x + self.pos_embed.to(x.dtype)
            x = self.pos_drop(x)
        outs = []
        residual, hidden_states = None, x
        if out_indices is None and inference_params is None and self._out_norm_fusable(x):
            return self._pool(self._forward_blocks_fused_norm(x))
        for layer_idx, layer in enumerate(self.layers):
            hidden_states, residual = layer(hidden_states, residual, inference_params=inference_params)
            if out_indices is not None and layer_idx in out_indices:
                outs.append(hidden_states)
        if out_indices is not None:
            return outs, (gh, gw)
        is_rms = isinstance(self.norm_f, RMSNorm)
        hidden_states = layer_norm_fn(hidden_states, self.norm_f.weight, self.norm_f.bias, eps=self.norm_f.eps,
                                      residual=residual, prenorm=False,
                                      residual_in_fp32=self.residual_in_fp32, is_rms_norm=is_rms)
        return self._pool(hidden_states)

    def _pool(self, hidden_states):
        if self.final_pool_type == "none":
            return hidden_states[:, -1, :]
        if self.final_pool_type == "mean":
            return hidden_states.mean(dim=1)
        if self.final_pool_type in ("max", "all"):
            return hidden_states
        raise NotImplementedError

    def _out_norm_fusable(self, x) -> bool:
        """Inference with bf16 activations, RMSNorm, fp32 residual stream and a d_model that fits one tcgen05 accumulator
        (FastVim-T): every residual add + norm after the first rides in the previous block's out_proj epilogue."""
        if torch.is_grad_enabled() or self.training or not x.is_cuda:
            return False
        act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        if act != torch.bfloat16 or not self.residual_in_fp32 or not isinstance(self.norm_f, RMSNorm):
            return False
        for layer in self.layers:
            if not isinstance(layer.norm, RMSNorm) or layer.norm.bias is not None \
                    or not hasattr(layer.mixer, "out_norm_fusable") or not layer.mixer.out_norm_fusable(x, act):
                return False
        return self.norm_f.bias is None

    def _forward_blocks_fused_norm(self, x):
        """The block stack with add + RMSNorm folded into the preceding out_proj (``fv_gemm_out_norm``): same arithmetic as
        the loop of ``forward_features`` (reference models/fastvim.py:497-527), 24 launches fewer."""
        first = self.layers[0]
        hidden, residual = layer_norm_fn(x, first.norm.weight, None, residual=None, prenorm=True,
                                         residual_in_fp32=True, eps=first.norm.eps, is_rms_norm=True)
        n = len(self.layers)
        # [0] tile counter, [1 + b] completion flag of image b: shared by the n (block kernel, out_proj GEMM) pairs of this
        # forward, zeroed here once (fv_block_fwd_signal / fv_gemm_out_norm_flow)
        sync = torch.zeros(x.shape[0] + 1, dtype=torch.int32, device=x.device)
        for i, layer in enumerate(self.layers):
            nxt = self.layers[i + 1].norm if i + 1 < n else self.norm_f
            rotated = layer.rotate_every_block is True and layer.layer_idx % 2 != 0
            hidden, residual = layer.mixer.forward_out_norm(hidden, rotated, residual, nxt.weight, nxt.eps,
                                                            want_residual=i + 1 < n, flow=(sync, i))
        return hidden

    def forward(self, x, return_features=False, inference_params=None):
        x = self.forward_features(x, inference_params)
        if return_features:
            return x
        x = F.linear(x, self.head.weight.to(x.dtype), self.head.bias.to(x.dtype)) if self.num_classes > 0 else x
        if self.final_pool_type == "max":
            x = x.max(dim=1)[0]
        return x


def _factory(embed_dim):
    def make(pretrained=False, img_size=224, patch_size=16, stride=16, **kwargs):
        if pretrained:
            raise ValueError("no pretrained weights are bundled (the reference's URL is 'to.do')")
        return VisionMamba(img_size=img_size, patch_size=patch_size, stride=stride, embed_dim=embed_dim,
                           depth=24, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
                           final_pool_type="mean", if_abs_pos_embed=True, **kwargs)
    return make


# reference factories, models/fastvim.py:695-819
vim_tiny_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2 = _factory(192)
vim_small_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2 = _factory(384)
vim_base_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2 = _factory(768)
fastvim_tiny, fastvim_small, fastvim_base = (
    vim_tiny_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2,
    vim_small_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2,
    vim_base_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2)
