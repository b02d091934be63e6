"""torchrun -n 2 check: fvdb.nn.SyncBatchNorm over two ranks == torch BatchNorm1d over the concatenated rows (outputs,
input gradients, parameter gradients after all-reduce, running statistics)."""
import os, sys
sys.path.insert(0, "fvdb-core_b200"); sys.path.insert(0, ".")
import torch, torch.distributed as dist
import fvdb

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
C = 64
rows = [5000, 3100][rank % 2]
gen = torch.Generator().manual_seed(100 + rank)
x = (torch.randn((rows, C), generator=gen) * (1.0 + rank) + 0.3 * rank).to(dev).requires_grad_()
dy = torch.randn((rows, C), generator=gen).to(dev)
bn = fvdb.nn.SyncBatchNorm(C, activation="relu").to(dev)
y = bn(fvdb.JaggedTensor([x]))
gx, gw, gb = torch.autograd.grad(y.jdata, (x, bn.weight, bn.bias), dy)
dist.all_reduce(gw); dist.all_reduce(gb)
xs = [torch.empty((r, C), device=dev) for r in (5000, 3100)]
dys = [torch.empty((r, C), device=dev) for r in (5000, 3100)]
dist.all_gather(xs, x.detach()) if False else None
# gather ragged rows through a padded buffer
def gather_rows(t):
    pad = torch.zeros((5000, C), device=dev); pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: (5000, 3100)[i % 2]] for i, o in enumerate(out)])
X, DY = gather_rows(x.detach()).requires_grad_(), gather_rows(dy)
ref = torch.nn.BatchNorm1d(C).to(dev)
Y = torch.relu(ref(X))
GX, GW, GB = torch.autograd.grad(Y, (X, ref.weight, ref.bias), DY)
lo = sum((5000, 3100)[i % 2] for i in range(rank)); hi = lo + rows
rel = lambda a, b: float((a - b).norm() / b.norm())
errs = dict(y=rel(y.jdata.detach(), Y[lo:hi].detach()), gx=rel(gx, GX[lo:hi]), gw=rel(gw, GW), gb=rel(gb, GB),
            rm=rel(bn.running_mean, ref.running_mean), rv=rel(bn.running_var, ref.running_var))
ok = all(v < 2e-5 for v in errs.values())
print(f"rank {rank}: {'OK' if ok else 'FAIL'} {errs}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
