#!/usr/bin/env python
"""bench.py -- FastVim images/s on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (one rank per GPU)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU path

A "step" is one forward pass of the hot path's model over one batch of synthetic images.
Headline workload at every N: BASELINE.json configs[1] -- FastVim-T (patch16, d=192, 24 blocks) inference,
224x224, bf16 autocast, batch 256 PER GPU (weak scaling: images are independent, no data-path
collective).  The default line also carries ``extra_workloads``: the FastVim-B training step
(configs[2]) and the 2048x2048 single image (configs[4], d_inner-sharded when N > 1) measured at the same N,
so the driver's 1/2/4/8 runs record their curves too.  ``--workload X`` runs one workload alone.

Keys (see DESIGN.md "Measurement"):
  value        images/s, whole job, inputs resident in HBM, CUDA-graph replay of the forward,
               CUDA events on the launching stream, max over ranks.
  e2e          the same metric through the public module API with HOST (pinned) fp32 images: every
               step copies its images host->device and its logits device->host inside the timed
               region (double-buffered on a copy stream).  e2e_bf16_host / e2e_u8: the same loop with
               bf16 / uint8 host images (the model accepts both; uint8 normalisation is folded into
               the patch embedding), with their own byte counts.
  roofline     dominant kernel of ours: algorithmic bytes / its duration INSIDE the replayed graph step
               (CUPTI activity records of the same graph the value was timed on), against MEASURED_PEAKS.json.
  kernels      per-kernel table from the same records (sum <= ms_per_step).
  cpu_baseline the oracle port of the reference's selective_scan_ref / mamba_inner_ref CPU path
               (oracle/fastvim_oracle.py) timed on this box's host cores on a bounded sample.
  gpu_competitor  the reference's own CUDA selective scan (oracle/_ref, built from /root/reference for
               sm_100a) timed beside fv_selective_scan_fwd/bwd on identical tensors.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "fastvim_t_224": dict(embed_dim=192, img=224, batch=256, desc="FastVim-T patch16 d192 24 blocks, 224x224 inference"),
    "fastvim_s_224": dict(embed_dim=384, img=224, batch=256, desc="FastVim-S patch16 d384 24 blocks, 224x224 inference"),
    "fastvim_b_224": dict(embed_dim=768, img=224, batch=128, desc="FastVim-B patch16 d768 24 blocks, 224x224 inference"),
    "fastvim_t_2048": dict(embed_dim=192, img=2048, batch=1, desc="FastVim-T patch16 d192 24 blocks, 2048x2048 inference"),
    # BASELINE.json configs[3]: FastChannelVim-S/16, 8-channel JUMP-CP-shape images, 14 x 14 patches x 8 channels = 1568 tokens
    "fastchannelvim_s_224": dict(embed_dim=384, img=224, batch=32, channels=8, classes=161, model="channel",
                                 desc="FastChannelVim-S/16 d384 24 blocks, 8-channel 224x224 inference, Channel-First"),
    # BASELINE.json configs[2]: supervised training step, batch-sharded DDP, 128 images per GPU
    "fastvim_b_224_train": dict(embed_dim=768, img=224, batch=128, train=True,
                                desc="FastVim-B patch16 d768 24 blocks, 224x224 training step (fwd+bwd+AdamW), bf16 autocast"),
    "fastvim_t_224_train": dict(embed_dim=192, img=224, batch=128, train=True,
                                desc="FastVim-T patch16 d192 24 blocks, 224x224 training step (fwd+bwd+AdamW), bf16 autocast"),
}
EXTRA_OF_DEFAULT = ["fastvim_b_224_train", "fastvim_t_2048"]
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is polled from a thread
    every few milliseconds (the driver's 20-step region lasts ~50 ms: nvidia-smi -lms 100 gave one sample); falls back to
    the nvidia-smi loop when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.rows = index, None, [], []
        self._stop, self.t, self.nv = False, None, None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.nv = nv
            self.mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.perf_counter(), float(sm), float(self.mx), pw,
                                  [n for n, b in bits.items() if rs & b]))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.nv is not None:
            self._stop = True
            self.t.join(timeout=1)
            rows = [r for r in self.rows if t0 is None or t0 <= r[0] <= t1 + 0.01] or self.rows[-3:]
            reasons = sorted({n for r in rows for n in r[4]})
            return {"sm_mhz": statistics.median(r[1] for r in rows) if rows else None,
                    "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                    "power_w_max": round(max(r[3] for r in rows), 1) if rows else None, "samples": len(rows),
                    "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


# ----------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(name: str, B: int, L: int, Lp: int, D: int, d_model: int, R: int, N: int, s: int) -> int:
    """Bytes one launch of kernel `name` must move (DESIGN.md "Kernels")."""
    if name == "fv_block_fwd_signal":           # fv_block_fwd + per-image completion flags: same traffic
        name = "fv_block_fwd"
    name = name.replace("fv_gemm_out_norm_flow[", "fv_gemm_out_norm[")
    if name == "fv_conv_pool_fwd":      # read x, write pooled u for both directions
        return B * L * D * s + 2 * B * Lp * D * s
    if name == "fv_scan_fwd":           # read u, x_dbl (both directions), write fp32 direction sum
        return 2 * B * Lp * D * s + 2 * B * Lp * (R + 2 * N) * s + B * Lp * D * 4
    if name == "fv_gate_fwd":           # read x, z, s; write gated y
        return 3 * B * L * D * s + B * Lp * D * 4
    if name == "fv_add_norm_fwd":       # read x (+ fp32 residual), write y + fp32 residual
        return B * L * d_model * (s + 4) * 2
    if name == "fv_block_fwd":          # fused conv+pool+x_proj+scan+gate: x, z read once, y written once
        return 3 * B * L * D * s
    if name == "fv_patchify":           # fp32 image in, bf16 patches out (3 x 16 x 16 values per token)
        return B * L * 768 * (4 + 2)
    if name.startswith("fv_gemm_bf16_tn["):   # A read, W read, C written (bf16)
        M, N_, K = (int(v) for v in name[name.index("[") + 1:-1].split("x")[:3])
        return (M * K + N_ * K + M * N_) * 2
    if name.startswith("fv_gemm_out_norm["):  # y read (bf16), W, fp32 residual read + written, normalised bf16 written
        M, N_, K = (int(v) for v in name[name.index("[") + 1:-1].split("x")[:3])
        return M * K * 2 + N_ * K * 2 + M * N_ * (4 + 4 + 2)
    return 0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def load_traffic(workload: str, kernel: str):
    """dram bytes per launch of `kernel` in `workload` from a committed ncu --set full capture of THAT workload
    (profiles/traffic.json: {workload: {kernel: bytes, "_source": file}}); None when no matching capture exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return (json.load(f).get(workload) or {}).get(kernel)
    return None


# ----------------------------------------------------------------------------- CPU (reference) arm
def cpu_reference_throughput(workload: str, budget_s: float, steps: int, warmup: int, threads: int):
    """images/s of the reference's CPU path (selective_scan_ref / mamba_inner_ref semantics, fp32) on `threads` host
    threads, on a bounded sample of the workload.  When ``oracle/build_ref.py`` has staged the reference's own Python
    (oracle/_ref/pyref, git-ignored) the model timed IS the reference's ``VisionMamba`` (kind "reference"); otherwise the
    oracle port (kind "port").  Returns (images/s, ms/step, sample batch, kind)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fastvim_oracle as O
    from fastvim_b200.vision import VisionMamba

    w = WORKLOADS[workload]
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = VisionMamba(img_size=w["img"], embed_dim=w["embed_dim"], depth=24, rms_norm=True, residual_in_fp32=True,
                    fused_add_norm=True, final_pool_type="mean", drop_path_rate=0.0).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    run, kind = (lambda x: O.fastvim_oracle(x, sd, depth=24)), "port"
    try:
        import ref_loader

        if ref_loader.reference_available():
            ref = ref_loader.load_reference()
            rm = ref_loader.build_reference_fastvim(ref, embed_dim=w["embed_dim"], depth=24, img_size=w["img"])
            rm.load_state_dict(sd, strict=True)
            run, kind = (lambda x: rm(x)), "reference"
    except Exception as ex:   # the staged reference is optional
        sys.stderr.write(f"[bench] staged reference not usable ({type(ex).__name__}: {ex}); timing the oracle port\n")
    # size the per-step sample so that (steps + warmup) steps fit the budget.  The reference's selective_scan_ref
    # materialises (batch, d_inner, L, d_state) tensors, so its cost per image GROWS with the batch: probe twice.
    probe_b = 1 if w["img"] > 512 else 4
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        sample_b = probe_b
        for _ in range(2):
            x = torch.randn(sample_b, 3, w["img"], w["img"], generator=g)
            if sample_b == probe_b:
                run(x)                      # first call pays one-time costs
            t0 = time.perf_counter()
            run(x)
            per_img = (time.perf_counter() - t0) / sample_b
            nxt = int(max(1, min(w["batch"], 64, budget_s / max(per_img, 1e-6) / (steps + warmup))))
            if nxt <= sample_b:
                sample_b = nxt
                break
            sample_b = nxt
        x = torch.randn(sample_b, 3, w["img"], w["img"], generator=g)
        for _ in range(warmup):
            run(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            run(x)
        dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps * 1e3, sample_b, kind


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    val, ms, sample_b, kind = cpu_reference_throughput(a.workload, a.cpu_budget, a.steps, a.warmup, cores)
    w = WORKLOADS[a.workload]
    what = ("the reference's own VisionMamba (models/fastvim.py, selective_scan_ref path; staged unmodified under "
            "oracle/_ref/pyref)" if kind == "reference" else "oracle port of selective_scan_ref/mamba_inner_ref")
    sample = (f"{w['desc']}, fp32, {what} on {cores} host threads; "
              f"each step = a batch of {sample_b} images (bounded sample of the {w['batch']}-image batch)")
    line = {"impl": "reference", "metric": "FastVim inference throughput", "value": round(val, 3), "unit": "images/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": a.workload, "desc": w["desc"], "batch_per_step": sample_b},
            "cpu_baseline": {"value": round(val, 3), "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(val, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- process context
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # keep NCCL's own prints (version banner at VERSION / WARN level, INFO lines when the caller asks for them)
            # out of stdout: rank 0 prints ONE JSON line there
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
                os.environ["NCCL_DEBUG"] = "NONE"
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fastvim_bench_nccl_%h_%p.log")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.torch.cuda.synchronize()
            self.dist.barrier()
            # a communicator that captured graphs still reference can block in destroy: bound it
            done = threading.Event()

            def _destroy():
                try:
                    self.dist.destroy_process_group()
                finally:
                    done.set()

            th = threading.Thread(target=_destroy, daemon=True)
            th.start()
            if not done.wait(timeout=20):
                sys.stdout.flush(); sys.stderr.flush()
                os._exit(0)


# ----------------------------------------------------------------------------- in-graph kernel table (CUPTI)
def profile_graph_kernels(torch, replay, tags, n=3):
    """Device time of every kernel INSIDE `n` replays of the step graph (torch.profiler = CUPTI activity records, no
    replay under a profiler tool).  Our kernels (namespace fv::) are matched by launch order with `tags`, the list of
    C-ABI calls an eager pass of the same step made; everything else (cuBLAS, torch elementwise) is summed as "other".
    Returns ({tag: [total_us, count]}, other_us, total_us) per step, or None when CUPTI is unavailable."""
    try:
        from torch.profiler import ProfilerActivity, profile

        replay(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(n):
                replay()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    except Exception as ex:   # pragma: no cover
        sys.stderr.write(f"[bench] CUPTI kernel table unavailable: {type(ex).__name__}: {ex}\n")
        return None
    kern = [e for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    kern.sort(key=lambda e: e.time_range.start)
    dur = lambda e: float(e.time_range.end - e.time_range.start)
    ours = [e for e in kern if "fv::" in e.name]
    other_us = sum(dur(e) for e in kern if "fv::" not in e.name) / n
    agg = {}
    if tags and len(ours) == len(tags) * n:
        for i, e in enumerate(ours):
            d = agg.setdefault(tags[i % len(tags)], [0.0, 0])
            d[0] += dur(e); d[1] += 1
    else:       # launch lists differ (e.g. a helper kernel): fall back to the kernel's own name
        for e in ours:
            nm = e.name.split("(")[0].replace("void ", "").split("<")[0]
            d = agg.setdefault(nm, [0.0, 0])
            d[0] += dur(e); d[1] += 1
    for d in agg.values():
        d[0] /= n; d[1] //= n
    return agg, other_us, sum(dur(e) for e in kern) / n


def eager_tags(torch, fwd, _lib):
    """One eager pass of the step with every C-ABI call recorded: the per-launch tag list (calls that launch k kernels
    contribute k entries)."""
    recs = []
    _lib.set_profile(recs)
    try:
        fwd()
    finally:
        _lib.set_profile(None)
    torch.cuda.synchronize()
    tags = []
    for r in recs:
        tags += [r[0]] * max(1, r[3])
    return tags


# ----------------------------------------------------------------------------- inference workloads
def run_infer(a, ctx: Ctx, workload: str, main: bool):
    torch, dist = ctx.torch, ctx.dist
    from fastvim_b200 import _lib
    from fastvim_b200.vision import VisionMamba

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    w = WORKLOADS[workload]
    Bt = (a.batch if main and a.batch else w["batch"])
    img, E = w["img"], w["embed_dim"]
    torch.manual_seed(0)
    C_in, n_cls, tpp = w.get("channels", 3), w.get("classes", 1000), 1
    if w.get("model") == "channel":
        from fastvim_b200.vision_channel import VisionMamba as ChannelVisionMamba
        model = ChannelVisionMamba(img_size=img, embed_dim=E, depth=24, channels=C_in, num_classes=n_cls, rms_norm=True,
                                   residual_in_fp32=True, fused_add_norm=True, final_pool_type="mean", drop_path_rate=0.0,
                                   scan_order="Channel-First", hcs=False).eval().to(dev)
        tpp = C_in
    else:
        model = VisionMamba(img_size=img, embed_dim=E, depth=24, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
                            final_pool_type="mean", drop_path_rate=0.0).eval().to(dev)
        model.set_input_normalization(IMAGENET_MEAN, IMAGENET_STD)     # meaning of uint8 inputs (e2e_u8)
    # one very large image: d_inner channels sharded over the ranks (strong scaling), same image on every rank
    sharded = img >= 1024 and world > 1
    shard_info = None
    runner = model
    if sharded:
        from fastvim_b200 import sharded as fv_sharded
        if a.out_mode == "hybrid" and fv_sharded.hybrid_supported(model, world, (img, img)):
            try:   # peer-memory exchanges (csrc/peer.cu); needs torch symmetric memory for the mapping
                runner = fv_sharded.shard_model_hybrid(model, None)
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    runner(torch.zeros(1, C_in, img, img, device=dev))
                torch.cuda.synchronize()
            except Exception as ex:
                sys.stderr.write(f"[bench] hybrid peer-memory sharding unavailable ({type(ex).__name__}: {ex}); NCCL path\n")
                runner = model
        if runner is model:
            fv_sharded.shard_model_channels(model, None, "gather" if a.out_mode == "hybrid" else a.out_mode)
            model._shard_desc = (f"d_inner channel-sharded x{world}: NCCL all-reduce of x_proj partials + LN sums, "
                                 f"{a.out_mode} around out_proj")
        shard_info = getattr(model, "_shard_desc", None)
    n_img_step = Bt if sharded else Bt * world
    g = torch.Generator(device="cpu").manual_seed(100 + (0 if sharded else rank))
    formats = ["f32"] if (sharded or w.get("model") == "channel" or a.no_e2e_variants) else ["f32", "bf16", "u8"]
    host_imgs = {"f32": [torch.randn(Bt, C_in, img, img, generator=g).pin_memory() for _ in range(2)]}
    if "bf16" in formats:
        host_imgs["bf16"] = [t.bfloat16().pin_memory() for t in host_imgs["f32"]]
        host_imgs["u8"] = [torch.randint(0, 256, (Bt, C_in, img, img), generator=g, dtype=torch.uint8).pin_memory()
                           for _ in range(2)]
    host_out = [torch.empty(Bt, n_cls).pin_memory() for _ in range(2)]

    def fwd(x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return runner(x)

    # ---- static buffers + CUDA graphs (two per host format, for the double-buffered e2e loops)
    static_in = {f: [host_imgs[f][i].to(dev) for i in range(2)] for f in formats}
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in formats:
            for _ in range(2):
                fwd(static_in[f][0])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs, static_out, launches_per_step = {}, {}, 0
    use_graph = not a.no_graph
    if use_graph:
        try:
            pool = None
            for f in formats:
                graphs[f], static_out[f] = [], []
                for i in range(2):
                    gr = torch.cuda.CUDAGraph()
                    _lib.reset_launch_count()
                    with torch.cuda.graph(gr, pool=pool):
                        o = fwd(static_in[f][i])
                    pool = gr.pool()
                    if f == "f32":
                        launches_per_step = _lib.launch_count()
                    graphs[f].append(gr)
                    static_out[f].append(o)
        except Exception as ex:  # launch mode only (e.g. a collective that cannot be captured): run eagerly
            sys.stderr.write(f"[bench] CUDA graph capture failed ({type(ex).__name__}: {ex}); running eagerly\n")
            use_graph, graphs, static_out = False, {}, {}
            torch.cuda.synchronize()
    if not use_graph:
        _lib.reset_launch_count()
        fwd(static_in["f32"][0])
        launches_per_step = _lib.launch_count()

    def step(i=0, f="f32"):
        if use_graph:
            graphs[f][i].replay()
            return static_out[f][i]
        return fwd(static_in[f][i])

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        step(0)
    sampler = ClockSampler(ctx.local)
    if rank == 0 and main:
        sampler.start()
        time.sleep(0.05)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        step(0)
    e1.record()
    ctx.barrier()
    tw1 = time.perf_counter()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / a.steps
    value = n_img_step * a.steps / (ms_total * 1e-3)
    clocks = sampler.stop(tw0, tw1) if (rank == 0 and main) else None

    # ---- e2e: pinned host images in, logits out, every step, double-buffered ----------------
    copy_s = torch.cuda.Stream(dev)
    comp_s = torch.cuda.current_stream()
    d2h_bytes = host_out[0].numel() * host_out[0].element_size()

    def e2e_loop(n, f):
        ev_copy = [None, None]
        ev_done = [None, None]
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(copy_s):
                if ev_done[b] is not None:
                    copy_s.wait_event(ev_done[b])      # buffer b is free once step i-2 finished
                static_in[f][b].copy_(host_imgs[f][b], non_blocking=True)
                ev_copy[b] = torch.cuda.Event()
                ev_copy[b].record(copy_s)
            comp_s.wait_event(ev_copy[b])
            out = step(b, f)
            host_out[b].copy_(out.float(), non_blocking=True)
            ev_done[b] = torch.cuda.Event()
            ev_done[b].record(comp_s)

    e2e = {}
    for f in formats:
        e2e_loop(max(a.warmup, 3), f)
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_loop(a.steps, f)
        torch.cuda.synchronize()          # includes the last device->host read
        t_e2e = time.perf_counter() - t0
        ctx.barrier()
        t_e2e = ctx.max_over_ranks(t_e2e)
        h2d = host_imgs[f][0].numel() * host_imgs[f][0].element_size()
        e2e[f] = {"value": round(n_img_step * a.steps / t_e2e, 1), "unit": "images/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h_bytes, "ms_per_step": round(t_e2e / a.steps * 1e3, 4),
                  "note": {"f32": "pinned fp32 images", "bf16": "pinned bf16 images (bit-identical logits)",
                           "u8": "pinned uint8 images, (x/255-mean)/std folded into the patch embedding"}[f]
                          + " -> H2D -> forward -> fp32 logits D2H, double-buffered copy stream"}

    # ---- roofline + kernel table: CUPTI records of the replayed graph --------------------------
    roof, kern_table, ktot = None, {}, None
    if rank == 0 or sharded:   # sharded: the forward contains collectives, every rank must run it
        m0 = model.layers[0].mixer
        m0 = getattr(m0, "mixer", m0)          # channel-sharded wrapper
        D_loc = m0.d_inner // (world if sharded else 1)
        L = (img // 16) ** 2 * tpp
        Lp = img // 16 * tpp
        peaks, peak_src = load_peaks()
        tags = eager_tags(torch, lambda: fwd(static_in["f32"][0]), _lib)
        # The timed graph launches the chain kernels with programmatic dependent launch: a kernel's CUPTI duration then
        # includes its wait for the predecessor.  The per-kernel table therefore comes from a SECOND capture of the same
        # step with PDL off (fv_set_pdl(0)): fully serialised kernels, true durations; `ms_per_step` / `value` stay those
        # of the PDL graph.
        prof, pdl_was = None, None
        if use_graph:
            try:
                pdl_was = int(_lib.lib().fv_set_pdl(0))
                gser = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gser, pool=graphs["f32"][0].pool()):
                    fwd(static_in["f32"][0])
                for _ in range(3):
                    gser.replay()
                es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                es0.record()
                for _ in range(10):
                    gser.replay()
                es1.record()
                torch.cuda.synchronize()
                ms_serial = es0.elapsed_time(es1) / 10
                prof = profile_graph_kernels(torch, gser.replay, tags)
            except Exception as ex:   # pragma: no cover
                sys.stderr.write(f"[bench] serialised capture failed ({type(ex).__name__}: {ex})\n")
                prof = profile_graph_kernels(torch, lambda: step(0), tags)
                ms_serial = None
            finally:
                if pdl_was is not None:
                    _lib.lib().fv_set_pdl(pdl_was)
        if prof is not None:
            agg, other_us, total_us = prof
            ktot = {"ours_ms": round(sum(v[0] for v in agg.values()) / 1e3, 4), "other_ms": round(other_us / 1e3, 4),
                    "sum_ms": round(total_us / 1e3, 4),
                    "ms_per_step_serialised": None if ms_serial is None else round(ms_serial, 4),
                    "pdl": bool(pdl_was),
                    "source": "CUPTI activity records of 3 replays of the same step captured with programmatic dependent "
                              "launch OFF (serialised kernels: durations exclude waits for the predecessor); the timed "
                              "graph has it ON and overlaps each kernel's prologue with the previous kernel's tail"}
            for name, (tot_us, cnt) in agg.items():
                ab = algorithmic_bytes(name, Bt, L, Lp, D_loc, E, m0.dt_rank, m0.d_state, 2)
                avg_us = tot_us / max(cnt, 1)
                kern_table[name] = {"launches_per_step": cnt, "avg_us": round(avg_us, 2), "ms_per_step": round(tot_us / 1e3, 4),
                                    "alg_bytes": ab, "gbs": round(ab / (avg_us * 1e-6) / 1e9, 1) if avg_us > 0 else None}
        else:   # eager instrumented pass (CUDA events around every C-ABI call)
            recs = []
            for _ in range(2):
                fwd(static_in["f32"][0])
            _lib.set_profile(recs)
            for _ in range(3):
                fwd(static_in["f32"][0])
            _lib.set_profile(None)
            torch.cuda.synchronize()
            agg = {}
            for name, ev0, ev1, _n in recs:
                d = agg.setdefault(name, [0.0, 0])
                d[0] += ev0.elapsed_time(ev1); d[1] += 1
            for name, (tot, cnt) in agg.items():
                ab = algorithmic_bytes(name, Bt, L, Lp, D_loc, E, m0.dt_rank, m0.d_state, 2)
                avg_ms = tot / cnt
                kern_table[name] = {"launches_per_step": cnt // 3, "avg_us": round(avg_ms * 1e3, 2),
                                    "ms_per_step": round(tot / 3, 4), "alg_bytes": ab,
                                    "gbs": round(ab / (avg_ms * 1e-3) / 1e9, 1) if avg_ms > 0 else None}
            ktot = {"source": "eager instrumented pass (CUDA events); times include launch gaps"}
        cand = {k: v for k, v in kern_table.items() if v["alg_bytes"] > 0 and not k.startswith("fv_gemm")}
        if cand:
            top = max(cand, key=lambda k: cand[k]["ms_per_step"])
            k = kern_table[top]
            roof = {"kernel": top, "bound": "hbm", "achieved": k["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(k["gbs"] / peaks["hbm_gbs"], 4), "frac_of_nominal_8tbs": round(k["gbs"] / 8000.0, 4),
                    "traffic": load_traffic(workload, "fv_block_fwd" if top == "fv_block_fwd_signal" else top),
                    "alg_bytes_per_launch": k["alg_bytes"], "avg_us": k["avg_us"], "peak_source": peak_src,
                    "share_of_step": round(k["ms_per_step"] / max(ms_step, (ktot or {}).get("sum_ms") or 0.0), 4),
                    "timing": ("CUPTI, replay of the same step captured with programmatic dependent launch off (serialised: the "
                               "kernel's own duration; in the timed graph it overlaps its neighbours)"
                               if (prof is not None and pdl_was) else
                               "inside the replayed graph step (CUPTI)" if prof is not None else "eager CUDA events")}

    # ---- sharded 2048^2: the bench's own logits against the CPU oracle ----
    parity = None
    if sharded and not a.no_cpu:
        got = step(0).float().cpu()        # the step holds collectives: EVERY rank replays it; rank 0 checks
        torch.cuda.synchronize()
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import fastvim_oracle as O
            nthr = torch.get_num_threads()
            torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))   # torchrun pins OMP_NUM_THREADS=1
            sd = {k.replace(".mixer.mixer.", ".mixer."): v.detach().float().cpu() for k, v in model.state_dict().items()}
            with torch.no_grad():
                want = O.fastvim_oracle(host_imgs["f32"][0], sd, depth=24)
            torch.set_num_threads(nthr)
            err = float((got - want).abs().max() / want.abs().max())
            parity = {"rel_err_vs_oracle": round(err, 6), "tol": 2e-2, "ok": err <= 2e-2}
        ctx.barrier()

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if main and rank == 0 and world == 1 and not a.no_cpu and w.get("model") != "channel":
        cores = os.cpu_count() or 1
        v, ms, sb, kind = cpu_reference_throughput(workload, a.cpu_budget, 2, 1, cores)
        cpu = {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": kind,
               "sample": ("the reference's own VisionMamba on its selective_scan_ref CPU path (staged unmodified, oracle/_ref/pyref)"
                          if kind == "reference" else
                          "oracle port of the reference CPU path (selective_scan_ref/mamba_inner_ref semantics)")
                         + f", fp32, {cores} threads, 2 timed steps of {sb} images each after 1 warm-up"}

    line = None
    if rank == 0:
        m0 = model.layers[0].mixer
        m0 = getattr(m0, "mixer", m0)
        L = (img // 16) ** 2 * tpp
        line = {"metric": "FastVim inference throughput", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
                "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "desc": w["desc"], "batch_per_gpu": Bt, "global_batch": n_img_step,
                           "sharding": (shard_info or f"d_inner channel-sharded x{world}") if sharded else
                           f"batch-sharded x{world}, no data-path collective",
                           "launch": "cuda_graph" if use_graph else "eager",
                           "l2": "per-step working set (24 blocks x ~%d MB of activations) exceeds the 126 MB L2; no flush"
                                 % (3 * Bt * L * m0.d_inner * 2 // 2**20)},
                "e2e": e2e["f32"],
                "gpu_launches": launches_per_step * a.steps, "gpu_launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kern_table, "kernels_total": ktot}
        if "bf16" in e2e:
            line["e2e_bf16_host"], line["e2e_u8"] = e2e["bf16"], e2e["u8"]
        if parity is not None:
            line["parity_check"] = parity
    # release everything this workload holds on the device (extra workloads follow in the same process)
    del graphs, static_out, static_in, model, runner
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return line


# ----------------------------------------------------------------------------- training workloads
def run_train(a, ctx: Ctx, workload: str, main: bool):
    """One supervised training step per `step`: forward + soft-target CE + backward + gradient exchange (N > 1) +
    fused AdamW, bf16 autocast, fp32 master weights."""
    torch, dist = ctx.torch, ctx.dist
    import torch.nn.functional as F

    from fastvim_b200 import _lib, parallel
    from fastvim_b200.vision import VisionMamba

    world, rank, dev, local = ctx.world, ctx.rank, ctx.dev, ctx.local
    w = WORKLOADS[workload]
    Bt = (a.batch if main and a.batch else w["batch"])
    img = w["img"]
    torch.manual_seed(0)
    model = VisionMamba(img_size=img, embed_dim=w["embed_dim"], depth=24, rms_norm=True, residual_in_fp32=True,
                        fused_add_norm=True, final_pool_type="mean", drop_path_rate=0.0).to(dev)
    model.train()
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_imgs = [torch.randn(Bt, 3, img, img, generator=g).pin_memory() for _ in range(2)]
    host_tgt = [torch.softmax(torch.randn(Bt, 1000, generator=g) * 3, -1).pin_memory() for _ in range(2)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_tgt = [t.to(dev) for t in host_tgt]
    host_loss = torch.zeros(1).pin_memory()
    params = [p for p in model.parameters() if p.requires_grad]

    def fwd_bwd(net, x, t):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(x)
        loss = torch.sum(-t * F.log_softmax(logits.float(), dim=-1), dim=-1).mean()   # SoftTargetCrossEntropy
        loss.backward()
        return loss

    # ---- launch mode.  "graph" (default): the step is launch-bound on the host (~3,000 kernel launches), so forward +
    # backward and the optimizer are captured in CUDA graphs over static input buffers.  With N > 1 the gradient exchange
    # is fastvim_b200.parallel.GradExchange: gradients are bucketed in reverse parameter order and every bucket's NCCL
    # all-reduce is issued on a side stream INSIDE the captured backward as soon as its last gradient exists, so the
    # exchange overlaps the rest of the backward (the reference's DDP does the same, imagenet_classification/train.py).
    # "eager": torch DDP.
    mode = "eager" if a.no_graph else "graph"
    step = None
    exch_desc = None
    if mode == "graph":
        try:
            opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05, fused=True, capturable=True)
            static_x, static_t = dev_imgs[0].clone(), dev_tgt[0].clone()
            exch = parallel.GradExchange(params, world, bucket_mb=a.bucket_mb) if world > 1 else None
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    opt.zero_grad(set_to_none=True)
                    fwd_bwd(model, static_x, static_t)
                    if exch is not None:
                        exch.eager_allreduce_()
                    opt.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if exch is not None:
                exch_desc = exch.attach()      # flat buckets + hooks that pack and all-reduce them
            else:
                opt.zero_grad(set_to_none=True)
            g1 = torch.cuda.CUDAGraph()
            _lib.reset_launch_count()
            with torch.cuda.graph(g1):
                if exch is not None:
                    exch.begin()               # drop the old gradients: autograd assigns, the hooks pack the buckets
                static_loss = fwd_bwd(model, static_x, static_t)
                if exch is not None:
                    exch.finish()              # join the side stream (every bucket reduced)
                opt.step()
            launches = _lib.launch_count()

            def step(x, t):
                static_x.copy_(x, non_blocking=True)
                static_t.copy_(t, non_blocking=True)
                g1.replay()
                return static_loss
        except Exception as ex:  # capture not possible: fall back to the eager step
            sys.stderr.write(f"[bench] training-step graph capture failed ({type(ex).__name__}: {ex}); running eagerly\n")
            mode, step = "eager", None
            torch.cuda.synchronize()
            model.zero_grad(set_to_none=True)
    if step is None:
        net = model
        if world > 1:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                            static_graph=True, bucket_cap_mb=100)
        opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05, fused=True)

        def step(x, t):
            loss = fwd_bwd(net, x, t)
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss

    for _ in range(max(a.warmup, 3)):
        step(dev_imgs[0], dev_tgt[0])
    if mode == "eager":
        _lib.reset_launch_count()
        step(dev_imgs[0], dev_tgt[0])
        launches = _lib.launch_count()
    sampler = ClockSampler(local)
    if rank == 0 and main:
        sampler.start()
        time.sleep(0.05)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        loss = step(dev_imgs[i & 1], dev_tgt[i & 1])
    e1.record()
    ctx.barrier()
    tw1 = time.perf_counter()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    value = world * Bt * a.steps / (ms_total * 1e-3)
    clocks = sampler.stop(tw0, tw1) if (rank == 0 and main) else None
    # e2e: images + soft targets from pinned host memory, loss back to the host, every step
    for i in range(3):
        x = host_imgs[i & 1].to(dev, non_blocking=True); t = host_tgt[i & 1].to(dev, non_blocking=True)
        host_loss.copy_(step(x, t).detach().reshape(1), non_blocking=True)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        x = host_imgs[i & 1].to(dev, non_blocking=True); t = host_tgt[i & 1].to(dev, non_blocking=True)
        host_loss.copy_(step(x, t).detach().reshape(1), non_blocking=True)
    torch.cuda.synchronize()
    t_e2e = ctx.max_over_ranks(time.perf_counter() - t0)
    ctx.barrier()
    # kernel-level split of the captured step (rank 0): top kernels by device time
    top = None
    if rank == 0 and mode == "graph" and world == 1:
        prof = profile_graph_kernels(torch, lambda: g1.replay(), None, n=2)
        if prof is not None:
            agg, other_us, total_us = prof
            rows = sorted(agg.items(), key=lambda kv: -kv[1][0])[:10]
            top = {"ours_ms": round(sum(v[0] for v in agg.values()) / 1e3, 3), "other_ms": round(other_us / 1e3, 3),
                   "top": {k: {"ms": round(v[0] / 1e3, 3), "launches": v[1]} for k, v in rows}}
    line = None
    if rank == 0:
        h2d = host_imgs[0].numel() * 4 + host_tgt[0].numel() * 4
        line = {"metric": "FastVim training throughput", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms_total / a.steps, 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "desc": w["desc"], "batch_per_gpu": Bt, "global_batch": Bt * world,
                           "sharding": (f"batch-sharded x{world}; " + (exch_desc or "no exchange")) if mode == "graph" else
                                       f"batch-sharded DDP x{world}, NCCL gradient all-reduce overlapped with backward",
                           "optimizer": "AdamW fused, lr 1e-3, wd 0.05, fp32 master weights",
                           "launch": "cuda_graph (fwd + bwd + exchange + optimizer)" if mode == "graph" else "eager",
                           "l2": "activations of one step exceed the 126 MB L2; no flush"},
                "e2e": {"value": round(world * Bt * a.steps / t_e2e, 1), "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": round(t_e2e / a.steps * 1e3, 4)},
                "gpu_launches": launches * a.steps, "gpu_launches_per_step": launches, "clocks": clocks,
                "loss": float(loss.item()), "roofline": None, "cpu_baseline": None, "kernels_total": top}
    if mode == "graph":
        del g1
    del model, opt
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return line


# ----------------------------------------------------------------------------- GPU competitor
def gpu_competitor(ctx: Ctx):
    """The reference's own CUDA selective scan (mamba-1p1p1/csrc/selective_scan, compiled unmodified for sm_100a into
    oracle/_ref by oracle/build_ref.py) timed beside fv_selective_scan_fwd / _bwd on identical bf16 tensors at the four
    op shapes of SURVEY.md 8d.  CUDA-graph replay of 10 launches over rotating buffers (> L2 in total where the shape
    allows), best of 3, so neither side pays host launch overhead.  Measurement only: never on the product path."""
    torch = ctx.torch
    import importlib.util

    so = os.path.join(ROOT, "oracle", "_ref", "selective_scan_cuda.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/selective_scan_cuda.so not built (needs /root/reference in the build container)"}
    try:
        spec = importlib.util.spec_from_file_location("selective_scan_cuda", so)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    except Exception as ex:
        return {"unavailable": f"cannot load the reference CUDA scan: {type(ex).__name__}: {ex}"}
    from fastvim_b200 import ops

    def timeit(fn, nrot, iters=10):
        for i in range(3):
            fn(i % nrot)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(iters):
                fn(i % nrot)
        gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters * 1e3)
        del gr
        return best

    out = {"what": "reference selective_scan_cuda.fwd/.bwd (oracle/_ref, sm_100a build of the unmodified sources) vs "
                   "fv_selective_scan_fwd/_bwd, bf16, D + delta_bias + softplus, no z; us per launch", "shapes": {}}
    dt = torch.bfloat16
    for (batch, dim, L, N) in [(256, 384, 14, 16), (128, 1536, 14, 16), (32, 768, 112, 16), (1, 384, 128, 16)]:
        per = batch * dim * L * 2 * 4
        nrot = max(2, min(8, int(300e6 // max(per, 1)) + 1))
        gen = torch.Generator(device=ctx.dev).manual_seed(0)
        r = lambda *s: torch.rand(*s, device=ctx.dev, generator=gen)
        n = lambda *s: torch.randn(*s, device=ctx.dev, generator=gen)
        A, D, db = -0.5 * r(dim, N), n(dim), 0.5 * r(dim)
        sets = [dict(u=n(batch, dim, L).to(dt), delta=(0.5 * r(batch, dim, L)).to(dt), B=n(batch, 1, N, L).to(dt),
                     C=n(batch, 1, N, L).to(dt), dout=n(batch, dim, L).to(dt)) for _ in range(nrot)]
        try:
            t_ref_f = timeit(lambda i: ref.fwd(sets[i]["u"], sets[i]["delta"], A, sets[i]["B"], sets[i]["C"], D, None, db, True), nrot)
            t_our_f = timeit(lambda i: ops.selective_scan_fwd(sets[i]["u"], sets[i]["delta"], A, sets[i]["B"], sets[i]["C"], D,
                                                              None, db, True), nrot)
            xs = [ref.fwd(s_["u"], s_["delta"], A, s_["B"], s_["C"], D, None, db, True)[1] for s_ in sets]
            t_ref_b = timeit(lambda i: ref.bwd(sets[i]["u"], sets[i]["delta"], A, sets[i]["B"], sets[i]["C"], D, None, db,
                                               sets[i]["dout"], xs[i], None, None, True, False), nrot)
            t_our_b = timeit(lambda i: ops.selective_scan_bwd(sets[i]["dout"], sets[i]["u"], sets[i]["delta"], A, sets[i]["B"],
                                                              sets[i]["C"], D, None, db, True), nrot)
            out["shapes"]["%dx%dx%dx%d" % (batch, dim, L, N)] = {
                "ref_fwd_us": round(t_ref_f, 2), "ours_fwd_us": round(t_our_f, 2), "fwd_speedup": round(t_ref_f / t_our_f, 2),
                "ref_bwd_us": round(t_ref_b, 2), "ours_bwd_us": round(t_our_b, 2), "bwd_speedup": round(t_ref_b / t_our_b, 2)}
        except Exception as ex:
            out["shapes"]["%dx%dx%dx%d" % (batch, dim, L, N)] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        del sets
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- main
def run_ours(a):
    ctx = Ctx()
    w = WORKLOADS[a.workload]
    runner = run_train if w.get("train") else run_infer
    line = runner(a, ctx, a.workload, True)
    default = a.workload == "fastvim_t_224" and not a.batch and not a.no_extra
    if default:
        extras = {}
        for name in EXTRA_OF_DEFAULT:
            try:
                r = (run_train if WORKLOADS[name].get("train") else run_infer)(a, ctx, name, False)
            except Exception as ex:
                r = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                ctx.torch.cuda.synchronize()
            if ctx.rank == 0:
                keep = ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e", "roofline", "gpu_launches_per_step",
                        "kernels", "kernels_total", "parity_check", "loss", "error")
                extras[name] = {k: r[k] for k in keep if r and k in r}
        if ctx.rank == 0:
            line["extra_workloads"] = extras
            if ctx.world == 1 and not a.no_competitor:
                try:
                    line["gpu_competitor"] = gpu_competitor(ctx)
                except Exception as ex:
                    line["gpu_competitor"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fastvim_t_224", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads / gpu_competitor on the default line")
    ap.add_argument("--no-competitor", action="store_true")
    ap.add_argument("--no-e2e-variants", action="store_true", help="only the fp32-host e2e loop")
    ap.add_argument("--bucket-mb", type=float, default=64.0, help="gradient bucket size of the captured exchange")
    ap.add_argument("--out-mode", default="hybrid", choices=["hybrid", "gather", "reduce"],
                    help="single-image multi-GPU mode: hybrid token/channel sharding over peer memory (default), or the "
                         "round-1 NCCL channel sharding (all-gather y before out_proj | row-sharded out_proj + all-reduce)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the CPU legs")
    a = ap.parse_args()
    if a.impl == "reference":
        a.cpu_budget = max(a.cpu_budget, 90.0) if a.cpu_budget == 20.0 else a.cpu_budget
        return run_reference_arm(a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
