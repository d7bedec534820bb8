"""2+ rank smoke of the data-parallel step with progress markers (debug tool; run under torchrun with a short timeout)."""
import datetime
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cdcmdr_b200 as cm
import bench as Bn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])


def log(*a):
    print(f"[r{rank} {time.time() % 1000:8.2f}]", *a, file=sys.stderr, flush=True)


torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
log("pg up")
B = int(os.environ.get("SMOKE_B", "4096"))
Bn.Cfg.cdcmdr_precision = os.environ.get("SMOKE_PREC", "bf16")
torch.manual_seed(Bn.SEED)
model = cm.CDC(Bn.field_dims(), Bn.E, Bn.T, Bn.N_DOMAIN, "ple", Bn.EXPERT_DIMS, Bn.TOWER_DIMS, Bn.DOMAIN_IDX, dropout=Bn.DROPOUT,
               config=Bn.Cfg(), **Bn.L2).to(dev).train()
model.set_groups([d % Bn.T for d in range(Bn.N_DOMAIN)])
base = model.base_model_instance
cm.parallel.attach_data_parallel(base, dist.group.WORLD)
opt = cm.Adam(model.parameters(), **Bn.ADAM)
x, y, d = Bn.make_batches(1, B, 7, rank)[0]
xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
log("model built; eager steps")
for i in range(3):
    out = model.train_step(xt, yt, opt, mode="split", domain_i=d)
    torch.cuda.synchronize()
    log("eager step", i, model.step_losses(out))
g = cm.GraphedTrainStep(model, opt, B, Bn.F, mode="split", domain_i=d)
g.x.copy_(xt); g.y.copy_(yt)
log("capturing")
g.capture()
torch.cuda.synchronize()
log("captured; launches", g.launches_per_step)
for i in range(3):
    out = g()
    torch.cuda.synchronize()
    log("graph step", i, model.step_losses(out))
del g
dist.barrier()
torch.cuda.synchronize()
log("done")
sys.stderr.flush()
os._exit(0)
