"""GPU: the UNMODIFIED reference model files running on the B200 kernels through the import shims.

``oracle/build_ref.py`` stages the reference's own ``models/*.py`` (unmodified, git-ignored, under ``oracle/_ref/pyref``)
so that they exist on the GPU box.  With ``fastvim_b200/compat`` first on ``sys.path`` the reference's
``models/fastvim.py`` binds ``mamba_ssm.modules.mamba_simple_faster.Mamba`` / ``mamba_ssm.ops.triton.layernorm`` to this
repo's mixer and norm: its own ``VisionMamba.forward`` -- including the reference ``Block``'s
physical odd-layer token rotation (models/fastvim.py:192-210) -- then runs on ``libfastvim_b200.so``.  Logits are compared
with the CPU oracle."""
import importlib
import os
import sys

import pytest
import torch

import fastvim_oracle as O
from util import TOL, assert_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")


@pytest.fixture()
def shimmed():
    if not os.path.isfile(os.path.join(PYREF, "models", "fastvim.py")):
        pytest.skip("reference model files not staged (python oracle/build_ref.py in the build container)")
    import ref_loader

    def purge():
        for k in [k for k in sys.modules if k.split(".")[0] in ("mamba_ssm", "models")]:
            del sys.modules[k]

    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("mamba_ssm", "models")}
    purge()
    ref_loader._install_shims()
    paths = [os.path.join(ROOT, "fastvim_b200", "compat"), ROOT, PYREF]
    for p in reversed(paths):
        sys.path.insert(0, p)
    try:
        yield
    finally:
        for p in paths:
            sys.path.remove(p)
        purge()
        sys.modules.update(saved)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reference_fastvim_tiny_forward_on_b200_kernels(shimmed, dtype, monkeypatch):
    """The reference's own FastVim-T factory, batch 3, 224 x 224: logits vs the oracle."""
    from fastvim_b200 import _lib, mixer

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)   # the reference's nn.Conv2d patch embedding in true fp32

    fv = importlib.import_module("models.fastvim")
    assert fv.__file__.startswith(PYREF) and fv.Mamba is mixer.Mamba
    torch.manual_seed(0)
    model = fv.vim_tiny_patch16_224_final_pool_mean_abs_pos_embed_with_noclstok_div2(drop_path_rate=0.0).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    imgs = torch.randn(3, 3, 224, 224)
    want = O.fastvim_oracle(imgs, sd, depth=24)
    model = model.cuda()
    _lib.reset_launch_count()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        got = model(imgs.cuda())
    assert _lib.launch_count() >= 24 * 3, "the reference model did not run on the library's kernels"
    assert_close(got, want, TOL[dtype], f"reference VisionMamba over the shims, {dtype}")


def test_reference_fastvim_training_step_on_b200_kernels(shimmed, monkeypatch):
    """Reference model file, 4 blocks, forward + backward through this repo's autograd functions: every gradient vs the
    fp64 oracle."""
    fv = importlib.import_module("models.fastvim")
    torch.manual_seed(0)
    # the reference's own PatchEmbed is an nn.Conv2d: cuDNN would run it in TF32 by default (1e-3 relative), which is the
    # reference's arithmetic, not this repo's -- compare fp32 with fp32
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    model = fv.VisionMamba(img_size=(64, 96), patch_size=16, stride=16, embed_dim=32, depth=4, num_classes=10, rms_norm=True,
                           residual_in_fp32=True, fused_add_norm=True, final_pool_type="mean", if_abs_pos_embed=True,
                           drop_path_rate=0.0)
    sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in model.state_dict().items()}
    imgs = torch.randn(2, 3, 64, 96)
    tgt = torch.tensor([3, 7])
    logits_o = O.fastvim_oracle(imgs.double(), sd, depth=4)
    torch.nn.functional.cross_entropy(logits_o, tgt).backward()
    model = model.cuda().train()
    logits = model(imgs.cuda())
    torch.nn.functional.cross_entropy(logits.float(), tgt.cuda()).backward()
    assert_close(logits, logits_o.detach(), 1e-4, "logits")
    for k, v in model.named_parameters():
        assert v.grad is not None, k
        assert_close(v.grad, sd[k].grad, 2e-4, f"d {k}")


def test_reference_mae_model_imports_and_runs_over_shims(shimmed):
    """models/mae/models_mamba_faster_mae_vimdecoder_v2.py: the encoder blocks bind to Mamba_masked, the decoder blocks to the
    plain Vim mixer (mamba_ssm.modules.mamba_simple.Mamba shim); one forward + backward of the reference's own MAE."""
    from fastvim_b200 import mixer_masked, mixer_plain

    mm = importlib.import_module("models.mae.models_mamba_faster_mae_vimdecoder_v2")
    assert mm.__file__.startswith(PYREF)
    assert mm.Mamba_masked is mixer_masked.Mamba_masked and mm.Mamba is mixer_plain.Mamba
    torch.manual_seed(0)
    cls = mm.MaskedAutoencoderViM
    model = cls(img_size=64, patch_size=16, embed_dim=32, depth=2, decoder_embed_dim=32, decoder_depth=1, rms_norm=True,
                residual_in_fp32=True, fused_add_norm=True).cuda().train()
    imgs = torch.randn(2, 3, 64, 64, device="cuda")
    out = model(imgs, mask_ratio=0.5)
    loss = out[0] if isinstance(out, (tuple, list)) else out
    assert torch.isfinite(loss).all()
    loss.backward()
    n = sum(1 for p in model.parameters() if p.grad is not None)
    assert n > 0 and all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
