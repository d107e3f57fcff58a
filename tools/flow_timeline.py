#!/usr/bin/env python
"""Kernel start / end times (CUPTI, via torch.profiler) of one FastVim-T inference step replayed from a CUDA graph:
shows how far fv_gemm_out_norm_flow overlaps fv_block_fwd_signal (programmatic dependent launch + per-image flags).

    python tools/flow_timeline.py [--batch 256] [--blocks 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastvim_b200.vision import fastvim_tiny  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--blocks", type=int, default=3)
a = ap.parse_args()
torch.manual_seed(0)
model = fastvim_tiny().cuda().eval()
img = torch.randn(a.batch, 3, 224, 224).cuda()


def fwd():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return model(img)


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        fwd()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    fwd()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "fv::" in e.name]
evs.sort(key=lambda e: e.time_range.start)
t0 = None
shown = 0
for e in evs:
    nm = e.name.split("(")[0].replace("void ", "").split("<")[0].replace("fv::", "")
    if t0 is None and nm.startswith("block_fwd"):
        t0 = e.time_range.start
    if t0 is None:
        continue
    print(f"{nm:28s} start {e.time_range.start - t0:8.1f} us   end {e.time_range.end - t0:8.1f} us   dur {e.time_range.end - e.time_range.start:7.1f}")
    if nm.startswith("gemm_tc"):
        shown += 1
        if shown >= a.blocks:
            break
