"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference Python.

Imports insitro/FastVim's own modules from ``/root/reference`` (read-only, only
present in the build container, never on the GPU box) so that
``oracle/gen_golden.py`` can (a) pin the CPU oracle in ``oracle/fastvim_oracle.py``
against the reference itself and (b) write golden vectors to ``tests/golden``.

The reference imports CUDA-only / un-vendored packages at module import time
(``mamba_ssm/ops/selective_scan_interface.py:3-4,7``; ``models/fastvim.py:10-22``).
We pre-insert tiny stand-ins for them:

* ``causal_conv1d`` / ``causal_conv1d_cuda`` (causal-conv1d==1.1.3.post1, PyPI, not
  vendored; pinned in the reference ``README.md:43``): replaced by the PyTorch form
  the reference itself uses as its fallback, ``act(conv1d(x)[..., :seqlen])``
  (``mamba_ssm/modules/mamba_simple.py:302-303``).
* ``selective_scan_cuda``: not built; ``selective_scan_fn`` is rebound to the
  reference's own ``selective_scan_ref`` (``selective_scan_interface.py:126-206``).
* ``timm`` / ``mmdet`` / ``mmseg`` helpers that only touch registration and random
  init (``models/fastvim.py:10-12, 21-22``).

Nothing here is imported by the product package, tests marked ``gpu``, ``smoke()`` or
``bench.py``.
"""
from __future__ import annotations

import math
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref", "pyref")     # oracle/build_ref.py stages the reference's .py files here (git-ignored)
REFERENCE_ROOT = os.environ.get("FASTVIM_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/mamba-1p1p1/mamba_ssm") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mamba-1p1p1", "mamba_ssm"))


def _causal_conv1d_fn(x, weight, bias=None, seq_idx=None, activation=None, **_kw):
    """(B, D, L) depthwise causal conv, the reference's own fallback form
    (mamba_simple.py:302-303): conv1d with padding=W-1, truncated to L, then SiLU."""
    if isinstance(seq_idx, str):
        # mamba_inner_ref passes "silu" positionally (selective_scan_interface.py:1778-1780),
        # which lands in ``seq_idx`` for causal-conv1d>=1.1; the fused CUDA path it is tested
        # against applies SiLU (``causal_conv1d_fwd(x, w, b, None, True)``, :250), so honour it.
        seq_idx, activation = None, seq_idx
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError(activation)
    d, w = weight.shape
    out = F.conv1d(x, weight.unsqueeze(1), bias, padding=w - 1, groups=d)[..., : x.shape[-1]]
    return out if activation is None else F.silu(out)


class _DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)


def _lecun_normal_(t):
    fan_in = nn.init._calculate_fan_in_and_fan_out(t)[0]
    return nn.init.trunc_normal_(t, std=math.sqrt(1.0 / fan_in) / 0.87962566103423978)


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class _Registry:
    def register_module(self, *a, **k):
        def deco(f):
            return f

        return deco


def _install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "causal_conv1d" not in sys.modules:
        mod("causal_conv1d", causal_conv1d_fn=_causal_conv1d_fn, causal_conv1d_update=None)
        mod("causal_conv1d_cuda")
    if "selective_scan_cuda" not in sys.modules:
        mod("selective_scan_cuda")
    if "timm" not in sys.modules:
        timm = mod("timm")
        timm.layers = mod(
            "timm.layers",
            DropPath=_DropPath,
            lecun_normal_=_lecun_normal_,
            to_2tuple=_to_2tuple,
            trunc_normal_=_trunc_normal_,
        )
        timm.models = mod("timm.models", register_model=lambda f: f)
        timm.models.layers = mod(
            "timm.models.layers",
            DropPath=_DropPath,
            lecun_normal_=_lecun_normal_,
            to_2tuple=_to_2tuple,
            trunc_normal_=_trunc_normal_,
        )
        timm.models.registry = mod("timm.models.registry", register_model=lambda f: f)
        timm.models.vision_transformer = mod(
            "timm.models.vision_transformer",
            _cfg=lambda **k: dict(k),
            _load_weights=None,
            VisionTransformer=object,
        )
    if "mmdet" not in sys.modules:
        mmdet = mod("mmdet")
        mmdet.registry = mod("mmdet.registry", MODELS=_Registry())
    if "mmseg" not in sys.modules:
        mmseg = mod("mmseg")
        mmseg.models = mod("mmseg.models")
        mmseg.models.builder = mod("mmseg.models.builder", BACKBONES=_Registry())


_loaded = None


def load_reference():
    """Returns a namespace with the reference's own modules (imported, not copied)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_shims()
    for p in (os.path.join(REFERENCE_ROOT, "mamba-1p1p1"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    warnings.filterwarnings("ignore", category=FutureWarning)
    import mamba_ssm.ops.selective_scan_interface as ssi  # noqa: E402

    # CPU: route the CUDA op to the reference's own PyTorch reference.
    ssi.selective_scan_fn = ssi.selective_scan_ref
    import mamba_ssm.modules.mamba_simple_faster as msf  # noqa: E402
    import mamba_ssm.ops.triton.layernorm as ln  # noqa: E402

    # The Triton add+norm kernel needs a GPU; the reference ships its own PyTorch
    # statement of the same op (layernorm.py:18-49) -- route the functional API to it.
    def _rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False,
                     eps=1e-6):
        if residual is not None and residual_in_fp32:
            residual = residual.float()
        return ln.rms_norm_ref(x, weight, bias, residual=residual, eps=eps, prenorm=prenorm,
                               upcast=True)

    def _layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False,
                       residual_in_fp32=False, is_rms_norm=False):
        fn = ln.rms_norm_ref if is_rms_norm else ln.layer_norm_ref
        return fn(x, weight, bias, residual=residual, eps=eps, prenorm=prenorm, upcast=True)

    ln.rms_norm_fn = _rms_norm_fn
    ln.layer_norm_fn = _layer_norm_fn
    import models.fastvim as fastvim  # noqa: E402

    mscf = msmf = msmf2 = None
    try:
        import mamba_ssm.modules.mamba_simple_channel_faster as mscf  # noqa: E402  (FastChannelVim mixer)

        import mamba_ssm.modules.mamba_simple_masked_faster as msmf  # noqa: E402  (FastMaskVim encoder mixer)
        import mamba_ssm.modules.mamba_simple_masked_faster_v2 as msmf2  # noqa: E402
    except ImportError:     # only the FastVim classifier is needed by bench.py's reference arm
        pass

    ns = types.SimpleNamespace(ssi=ssi, msf=msf, mscf=mscf, msmf=msmf, msmf2=msmf2, ln=ln, fastvim=fastvim)
    _loaded = ns
    return ns


def build_reference_fastvim(ref, *, embed_dim=192, depth=24, img_size=224, channels=3,
                            num_classes=1000, seed=0, **kw):
    """FastVim built by the reference's own ``VisionMamba`` on CPU.

    ``fused_add_norm=False`` takes the reference's non-Triton add+norm branch
    (models/fastvim.py:158-166), which is its own CPU-runnable equivalent of
    ``rms_norm_fn(prenorm=True, residual_in_fp32=True)``.
    """
    torch.manual_seed(seed)
    model = ref.fastvim.VisionMamba(
        img_size=img_size, patch_size=16, stride=16, embed_dim=embed_dim, depth=depth,
        channels=channels, num_classes=num_classes, rms_norm=True, residual_in_fp32=True,
        fused_add_norm=False, final_pool_type="mean", if_abs_pos_embed=True,
        drop_path_rate=0.0, **kw,
    )
    return model.eval()
