"""Run under torchrun on >= 2 GPUs: the copy-engine push exchange (q3dr_exchange behind quad3dr_b200/dist.py PeerExchange) must deliver
exactly what the NCCL all_gather path delivers, step after step.  Prints 'peer gather OK' on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from quad3dr_b200 import engine as E, synth, dist as qdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
leaves = synth.map_small()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(320, 240)
n_total = 37 * world + 1                      # ragged: rank sizes differ by one
all_vps = synth.make_viewpoints(leaves, n_total, config_id=9)
lo, hi = qdist.shard_range(n_total, rank, world)
n_per = [qdist.shard_range(n_total, r, world)[1] - qdist.shard_range(n_total, r, world)[0] for r in range(world)]
vps = all_vps[lo:hi]
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth, device=local)
os.environ["Q3DR_CHUNK"] = "8"                # several chunks per call
stream = torch.cuda.current_stream()
d_rt = torch.from_numpy(np.ascontiguousarray(vps)).to(dev)
win = (0, 320, 0, 240)
gmap.set_result_fields(first_pixel=(rank % 2 == 0), dist_sq=False)  # the exchange ships (leaf, information) whatever the fields
pg = qdist.PeerExchange(gmap, dev, max(n_per), max(n_per) * 320 * 240)  # worst case: every pixel its own voxel
chunks = []
for step in range(3):
    n_use = len(vps) if step != 1 else len(vps) - 3          # a different result size in the middle step
    pg.begin_step()
    res = gmap.score_viewpoints_device(cam.K, 320, win, d_rt.data_ptr(), n_use, stream=stream.cuda_stream)
    pg.end_step(0, stream.cuda_stream)
    chunks.append(gmap.stats().chunks)
    n_use_per = [n if step != 1 else n - 3 for n in n_per]
    got = pg.views()
    off, tot, leaf, info = qdist.device_result_tensors(res, dev)
    ref = qdist.gather_results(off.clone(), tot.clone(), leaf.clone(), info.clone(), max(n_per))
    r_off, r_tot, r_leaf, r_info = ref
    v0 = 0
    e0 = 0
    for r in range(world):
        o, t, l, i = got[r]
        n = n_use_per[r]
        assert t.numel() == n and o.numel() == n + 1, (step, r, "viewpoint count")
        e = int(o[n].item())
        assert l.numel() == e, (step, r, "entry count")
        assert torch.equal(o[1:] - o[:-1], r_off[v0 + 1:v0 + n + 1] - r_off[v0:v0 + n]), (step, r, "counts")
        assert torch.equal(t.view(torch.int32), r_tot[v0:v0 + n].view(torch.int32)), (step, r, "totals")
        assert torch.equal(l, r_leaf[e0:e0 + e]), (step, r, "leaf")
        assert torch.equal(i.view(torch.int32), r_info[e0:e0 + e].view(torch.int32)), (step, r, "info")
        v0 += n
        e0 += e
    assert e0 == r_leaf.numel() and e0 > 1000
# a step whose result does not fit the receive slot must fail on ALL ranks (flag with the failure bit), not hang
pg.close()
pg = qdist.PeerExchange(gmap, dev, max(n_per), 64 if rank == 0 else max(n_per) * 320 * 240)
pg.begin_step()
gmap.score_viewpoints_device(cam.K, 320, win, d_rt.data_ptr(), len(vps), stream=stream.cuda_stream)
try:
    pg.end_step(0, stream.cuda_stream)
    raise SystemExit("expected the overflowing step to fail on every rank")
except E.Q3drError as e:
    assert "exchange" in str(e), e
# ... and the exchange is usable again afterwards
pg.close()
pg = qdist.PeerExchange(gmap, dev, max(n_per), max(n_per) * 320 * 240)
pg.begin_step()
res = gmap.score_viewpoints_device(cam.K, 320, win, d_rt.data_ptr(), len(vps), stream=stream.cuda_stream)
pg.end_step(0, stream.cuda_stream)
assert sum(int(v[2].numel()) for v in pg.views()) > 1000
pg.close()
dist.barrier()
if rank == 0:
    print("peer gather OK", "world", world, "chunks per step", chunks)
dist.destroy_process_group()
