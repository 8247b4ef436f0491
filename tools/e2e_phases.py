"""Where the end-to-end time of the exchange variant goes (N > 1, host buffers): wall time of stage 1,
the neighbour exchange and stage 2 per step, resident against host buffers.  Developer tool.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/e2e_phases.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import b4d, bench
from b4d.sharding import exchange_halo, exchange_planes, slab_plan

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S = 1024
own_b, own_e, zb, ze = slab_plan(S, world, rank, exchange_halo(11, 11))
dn = b4d.Denoiser(local)
slab_dev = bench.make_slab_device(S, zb, ze, dev)
slab_pin = torch.empty(slab_dev.shape, dtype=torch.uint16, pin_memory=True); slab_pin.copy_(slab_dev)
out_pin = torch.empty((own_e - own_b, S, S), dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize()

def step(slab, out):
    t = [time.perf_counter()]
    dn.slab_stage1(slab, zb, S, 24.0); t.append(time.perf_counter())
    lo_h, hi_h = own_b - zb, ze - own_e
    o0, o1 = own_b - zb, own_e - zb
    sd = dn.slab_basic(o0, lo_h, device=dev) if rank > 0 else None
    su = dn.slab_basic(o1 - hi_h, hi_h, device=dev) if rank < world - 1 else None
    rb = torch.empty((lo_h, S, S), dtype=torch.float32, device=dev) if rank > 0 else None
    ra = torch.empty((hi_h, S, S), dtype=torch.float32, device=dev) if rank < world - 1 else None
    exchange_planes(sd, su, rb, ra, rank, world, None)
    if rb is not None: dn.slab_set_basic(0, rb)
    if ra is not None: dn.slab_set_basic(o1, ra)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    y = dn.slab_stage2(own_b, own_e, out=out, device=dev if out is None else None); t.append(time.perf_counter())
    return [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])], dn.last_timings()

res = {}
for name, slab, out in (("resident", slab_dev, None), ("host", slab_pin.numpy(), out_pin.numpy())):
    for _ in range(2): step(slab, out)
    dist.barrier(); torch.cuda.synchronize()
    ph, tm = step(slab, out)
    res[name] = {"stage1_exchange_stage2_ms": ph, "stage2_device_ms": {k: round(v[0], 1) for k, v in tm.items() if v[0] > 0}}
    dist.barrier()
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    for r, x in enumerate(allr): print(r, json.dumps(x))
dist.destroy_process_group()
