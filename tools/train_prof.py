"""Kernel-level time table of one training step (torch.profiler / CUPTI; no nsys in the image).
usage: python tools/train_prof.py [--embed 768] [--batch 128] [--out gpurun_out/x/train_prof.txt]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

from fastvim_b200.vision import VisionMamba

ap = argparse.ArgumentParser()
ap.add_argument("--embed", type=int, default=768)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--img", type=int, default=224)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = VisionMamba(img_size=a.img, embed_dim=a.embed, depth=24, rms_norm=True, residual_in_fp32=True,
                    fused_add_norm=True, final_pool_type="mean", drop_path_rate=0.0).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05, fused=True)
x = torch.randn(a.batch, 3, a.img, a.img, device=dev)
t = torch.softmax(torch.randn(a.batch, 1000, device=dev) * 3, -1)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(x)
    loss = torch.sum(-t * F.log_softmax(logits.float(), dim=-1), dim=-1).mean()
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
txt = f"embed {a.embed} batch {a.batch} img {a.img}: {ms:.2f} ms / step un-profiled ({a.batch / ms * 1e3:.0f} img/s); table = 2 profiled steps\n" + tab
print(txt[-9000:])
if a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write(txt)
