"""Probe (run under torchrun on >= 2 GPUs): which peer-memory mechanism works on this box.
 1. torch.distributed._symmetric_memory (CUDA VMM handles)  2. legacy CUDA IPC via UntypedStorage._share_cuda_()"""
import os, sys, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    torch.cuda.synchronize(); dist.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    v = float(peer[:8].sum().item())
    res["symm_mem"] = f"ok peer_sum8={v} ptrs={len(hdl.buffer_ptrs)} pad={hdl.signal_pad_size}"
except Exception as ex:
    res["symm_mem"] = f"FAIL {type(ex).__name__}: {str(ex)[:200]}"
try:
    t2 = torch.full((1 << 20,), float(rank + 1), device=dev)
    h = t2.untyped_storage()._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, h)
    q = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*objs[q])
    peer2 = torch.tensor([], dtype=torch.float32, device=dev).set_(st)
    torch.cuda.synchronize(); dist.barrier()
    res["cuda_ipc"] = f"ok peer_sum8={float(peer2[:8].sum().item())} ptr={hex(peer2.data_ptr())}"
    big = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    hb = big.untyped_storage()._share_cuda_(); ob = [None] * world; dist.all_gather_object(ob, hb)
    pb = torch.tensor([], dtype=torch.uint8, device=dev).set_(torch.UntypedStorage._new_shared_cuda(*ob[q]))
    dst = torch.empty_like(big)
    for _ in range(3): dst.copy_(pb)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): dst.copy_(pb)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    res["p2p_read_GBps"] = round(big.numel() / dt / 1e9, 1)
except Exception as ex:
    res["cuda_ipc"] = f"FAIL {type(ex).__name__}: {str(ex)[:200]}"
print(f"[rank {rank}] {res}", flush=True)
dist.barrier(); dist.destroy_process_group()
