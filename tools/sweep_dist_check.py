"""torchrun target (2 GPUs): the real pruner sweep with the calibration batches sharded over the ranks -- T5 Wanda and
ViT SparseGPT on the tiny e2e models -- against the fixtures generated from the unmodified reference
(tests/golden/e2e_pruners.npz), plus bit-identical weights on all ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import e2e_cases as cases
from ecoflap_b200.compression import load_pruner

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gold = np.load(os.path.join(ROOT, "tests/golden/e2e_pruners.npz"))
ok = True


def check(model, prefix, need):
    global ok
    worst = 1.0
    for k, got in cases.prunable_state(model).items():
        key = f"{prefix}__{k}"
        if key not in gold.files or (gold[key] == 0).mean() < 0.05:
            continue
        worst = min(worst, float(((got == 0) == (gold[key] == 0)).mean()))
    # every rank must hold the very same pruned model
    flat = torch.cat([p.data.float().flatten() for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(flat, ref))
    if rank == 0:
        print(f"{prefix}: worst mask agreement with the reference fixture {worst:.4f} (need {need}), identical on all ranks: {same}", flush=True)
    ok = ok and worst >= need and same


m = cases.t5_model().cuda()
p = load_pruner("t5_wanda_pruner", m, cases.t5_loader(), cfg=dict(prune_spec="2-0.5-1.0-1.0", num_samples=16, model_prefix="t5_model"))
p.prune()
check(m, "t5_wanda", 0.995)
m = cases.vit_model().cuda()
p = load_pruner("vit_sparsegpt_pruner", m, cases.vit_loader(batch=1, n=48), cfg=dict(prune_spec="3-0.6-1.0-1.0", num_samples=48, model_prefix="visual"))
p.prune()
check(m, "vit_sparsegpt", 0.85)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SWEEP_DIST_CHECK", "OK" if int(t.item()) else "FAILED", flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0 if int(t.item()) else 1)
