"""Workload for ``compute-sanitizer`` (memcheck / racecheck) on the GPU box: short eager runs that launch every kernel
family of the library at BASELINE-shaped sizes (depth cut to 2 blocks) -- the fused FastVim-T chain with the block ->
out_proj dataflow, the cluster kernel (FastVim-S), the four-launch path at 2048², the channel-layout kernels, and one
bf16 training step (forward + backward).  ``python tools/sanitize_target.py [infer|train|all]``; see tools/gpu_sanitize.sh."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvim_b200 import _lib  # noqa: E402
from fastvim_b200.vision import VisionMamba  # noqa: E402

KW = dict(patch_size=16, stride=16, depth=2, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
          final_pool_type="mean", if_abs_pos_embed=True, drop_path_rate=0.0)


def infer():
    torch.manual_seed(0)
    for name, dim, size, batch in (("fastvim_t_224", 192, 224, 150), ("fastvim_s_224", 384, 224, 3),
                                   ("fastvim_b_224", 768, 224, 2), ("fastvim_t_2048", 192, 2048, 1)):
        m = VisionMamba(img_size=size, embed_dim=dim, **KW).eval().cuda()
        x = torch.randn(batch, 3, size, size, device="cuda")
        _lib.reset_launch_count()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all()
        print(f"[sanitize] {name} inference ok, {_lib.launch_count()} launches", flush=True)
    from fastvim_b200.vision_channel import VisionMamba as ChannelVim

    m = ChannelVim(img_size=224, depth=2, embed_dim=384, channels=8, num_classes=161, rms_norm=True, residual_in_fp32=True,
                   fused_add_norm=True, drop_path_rate=0.0, scan_order="Channel-First", hcs=False).eval().cuda()
    x = torch.randn(2, 8, 224, 224, device="cuda")
    _lib.reset_launch_count()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    print(f"[sanitize] fastchannelvim_s inference ok, {_lib.launch_count()} launches", flush=True)


def train():
    torch.manual_seed(1)
    for name, dim, batch in (("fastvim_t_224", 192, 4), ("fastvim_b_224", 768, 2)):
        m = VisionMamba(img_size=224, embed_dim=dim, **KW).train().cuda()
        x = torch.randn(batch, 3, 224, 224, device="cuda")
        _lib.reset_launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
        y.float().logsumexp(-1).mean().backward()
        torch.cuda.synchronize()
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
        print(f"[sanitize] {name} training step ok, {_lib.launch_count()} launches", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("infer", "all"):
        infer()
    if what in ("train", "all"):
        train()
