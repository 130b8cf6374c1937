"""torchrun target: the peer-memory norm exchange against NCCL on the same data (also under CUDA-graph replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ecoflap_b200 import dist as edist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
print('avail', edist.PeerNormExchange.available(), torch.cuda.is_available(), edist.is_dist(), str(dist.get_backend(None)), flush=True)
ex = edist.PeerNormExchange(40000, dev)
torch.manual_seed(100 + rank)
ok = True
for n in (10368, 30720, 7, 40000, 4096):
    for rep in range(3):
        x = torch.rand(n, device=dev) * (rank + 1)
        ref = x.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        got = x.clone()
        ex.sync(got)
        torch.cuda.synchronize()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        if err > 1e-6 or not same:
            ok = False
        if rank == 0:
            print(f"n={n} rep={rep} rel err {err:.2e} identical across ranks {same}", flush=True)
# graph replay + timing against NCCL
x = torch.rand(30720, device=dev)
for name, fn in (("p2p", lambda: ex.sync(x)), ("nccl", lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {e0.elapsed_time(e1) * 1e3 / 200:.1f} us per exchange of 30720 floats (graph replay, world {world})", flush=True)
    dist.barrier()
    del g
if rank == 0:
    print("P2P_CHECK", "OK" if ok else "FAILED", flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0 if ok else 1)
