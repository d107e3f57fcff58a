"""CPU, world_size 2, gloo: the host-side logic of the channel-sharded mixer (parameter partition and the
collective sequence: all-reduce of the x_proj partial product, all-reduce of the LayerNorm statistics,
all-gather / all-reduce around out_proj) reproduces the un-sharded oracle.  The per-channel stages are
evaluated with the oracle's functions (no CUDA kernels on CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _sharded_mixer_cpu(h, mixer, rank, world, rows, cols, out_mode):
    """Same dataflow as fastvim_b200.sharded.ChannelShardedMamba.forward, oracle arithmetic per stage."""
    import fastvim_oracle as O
    from fastvim_b200.sharded import shard_mixer_params

    pk = shard_mixer_params(mixer, rank, world, torch.float32)
    D, Dl, N, R = mixer.d_inner, mixer.d_inner // world, mixer.d_state, mixer.dt_rank
    B, L, _ = h.shape
    xz = F.linear(h, pk["in_w"], pk["in_b"])
    x, z = xz[..., :Dl].transpose(1, 2), xz[..., Dl:]
    xc_f = O.causal_conv1d_oracle(x, pk["conv_w"][0], pk["conv_b"][0])
    xc_b = O.causal_conv1d_oracle(x.flip(-1), pk["conv_w"][1], pk["conv_b"][1]).flip(-1)   # original order
    u = torch.stack([O.pool_oracle(xc_f, rows, cols), O.pool_oracle(xc_b, rows, cols)])    # (2, B, Dl, Lp)
    xdbl = torch.einsum("gbdl,gdk->gblk", u, pk["x_w_t"])                                   # partial over local channels
    dist.all_reduce(xdbl)
    s = 0
    for d in range(2):
        dt = torch.einsum("dr,blr->bdl", pk["dt_w"][d], xdbl[d][..., :R])
        Bm, Cm = xdbl[d][..., R:R + N].transpose(1, 2), xdbl[d][..., R + N:].transpose(1, 2)
        uu = u[d]
        if d == 1:
            uu, dt, Bm, Cm = uu.flip(-1), dt.flip(-1), Bm.flip(-1), Cm.flip(-1)
        sd = O.selective_scan_oracle(uu, dt, -torch.exp(pk["A_log"][d]), Bm, Cm, None, None, pk["dt_b"][d], True)
        s = s + (sd.flip(-1) if d == 1 else sd)
    v = (O.broadcast_oracle(s, rows, cols) + pk["D"][0][None, :, None] * xc_f + pk["D"][1][None, :, None] * xc_b) / 2
    v = v.transpose(1, 2)                                                                   # (B, L, Dl)
    stats = torch.stack([v.sum(-1), (v * v).sum(-1)], -1)
    dist.all_reduce(stats)
    mean = stats[..., 0] / D
    rstd = torch.rsqrt(stats[..., 1] / D - mean * mean + mixer.layernorm.eps)
    y = ((v - mean[..., None]) * rstd[..., None] * pk["ln_w"] + pk["ln_b"]) * F.silu(z)
    if out_mode == "reduce":
        out = F.linear(y, pk["out_w_loc"])
        dist.all_reduce(out)
        return out
    ys = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(ys, y.contiguous())
    return F.linear(torch.cat(ys, -1), pk["out_w"])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fastvim_oracle as O
    from fastvim_b200.mixer import Mamba

    torch.manual_seed(0)
    rows, cols, dm = 6, 5, 32
    mixer = Mamba(dm, token_size=[rows, cols], layer_idx=0).eval()
    with torch.no_grad():
        for p in mixer.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    h = torch.randn(2, rows * cols, dm)
    want = O.mixer_oracle(h, {k: v.detach() for k, v in mixer.state_dict().items()}, (rows, cols))
    errs = []
    with torch.no_grad():
        for mode in ("gather", "reduce"):
            got = _sharded_mixer_cpu(h, mixer, rank, world, rows, cols, mode)
            errs.append(float((got - want).abs().max() / want.abs().max()))
    if rank == 0:
        q.put(errs)
    dist.destroy_process_group()


def test_channel_sharded_dataflow_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    errs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max(errs) < 1e-5, errs


def test_shard_ranges_cover_all_channels():
    from fastvim_b200.sharded import shard_range

    for D, G in [(384, 8), (384, 2), (1536, 8), (64, 4)]:
        spans = [shard_range(D, r, G) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == D
        assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
    with pytest.raises(ValueError):
        shard_range(100, 0, 8)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastvim_b200 import parallel

    torch.manual_seed(0)
    shapes = [(7, 3), (5,), (2, 4, 6), (1,)]
    base = [torch.randn(*s) for s in shapes]                      # identical on every rank
    grads = [b * (rank + 1) for b in base]                        # rank-dependent "local" gradients
    flat = parallel.flatten_grads(grads)
    assert flat.shape == (sum(b.numel() for b in base),)
    parallel.allreduce_sum_(flat)
    parallel.scatter_mean_grads_(grads, flat, world)
    want_scale = sum(range(1, world + 1)) / world                 # mean over ranks of (rank + 1)
    err = max(float((g - b * want_scale).abs().max()) for g, b in zip(grads, base))
    if rank == 0:
        q.put(err)
    dist.destroy_process_group()


def test_batch_sharded_gradient_exchange_world2_gloo():
    """The N > 1 training path: gradients flattened, summed with ONE collective, scaled by 1/world while scattered back."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6, err


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastvim_b200 import parallel

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 5), torch.nn.Tanh(),
                              torch.nn.Linear(5, 3))
    unused = torch.nn.Parameter(torch.ones(4))                    # never receives a gradient: reduced in finish()
    params = list(net.parameters()) + [unused]
    exch = parallel.GradExchange(params, world, bucket_mb=60 * 4 / 2**20)   # 60-element buckets -> several buckets
    assert len(exch.plan) >= 3 and sorted(i for b in exch.plan for i in b) == list(range(len(params)))
    assert exch.plan[0][0] == len(params) - 1                    # buckets start from the LAST parameter
    exch.attach()
    errs = []
    for it in range(2):                                          # two steps: begin() must re-arm the hooks
        x = torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * it + rank))
        exch.begin()
        net(x).square().sum().backward()
        exch.finish()
        # reference: every rank's local gradients, averaged
        want = [torch.zeros_like(p) for p in params]
        for r in range(world):
            ref = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 5), torch.nn.Tanh(),
                                      torch.nn.Linear(5, 3))
            ref.load_state_dict(net.state_dict())
            xr = torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * it + r))
            ref(xr).square().sum().backward()
            for w_, p in zip(want, ref.parameters()):
                w_ += p.grad / world
        errs.append(max(float((p.grad - w_).abs().max()) for p, w_ in zip(params, want)))
        assert all(exch.fired)
    exch.detach()
    if rank == 0:
        q.put(max(errs))
    dist.destroy_process_group()


def test_bucketed_overlapped_gradient_exchange_world2_gloo():
    """GradExchange (bucketed all-reduce fired from post-accumulate hooks, the captured-step exchange of bench.py):
    averaged gradients equal the mean of the per-rank gradients, over two consecutive steps, incl. an unused parameter."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29890 + os.getpid() % 40
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6, err


def test_bucket_plan():
    from fastvim_b200.parallel import plan_buckets

    assert plan_buckets([10, 10, 10, 10], 20) == [[3, 2], [1, 0]]
    assert plan_buckets([5, 100, 5], 20) == [[2], [1], [0]]
    assert plan_buckets([1, 1, 1], 100) == [[2, 1, 0]]
