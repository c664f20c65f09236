"""One process driving all GPUs of the box through q3dr_multi (host buffers in, per-GPU host results out).
python tools/bench_multi.py [viewpoints_per_gpu] -> one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quad3dr_b200 import engine as E, synth

per = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
W, H = 640, 480
leaves = synth.map_1m()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(W, H)
n_dev = E.device_count()
t = time.perf_counter()
multi = E.MultiGpuMap(leaves.boxes, leaves.weight, leaves.normal, depth)
t_create = time.perf_counter() - t
vps = synth.make_viewpoints(leaves, min(per * n_dev, 4000), config_id=2)
vps = np.ascontiguousarray(np.tile(vps, (int(np.ceil(per * n_dev / len(vps))), 1))[: per * n_dev])
lib = E.load_library()
import ctypes as C
opts = E.default_opts()
r = E._MultiResult()
K = np.ascontiguousarray(cam.K, dtype=np.float32)
call = lambda: E._check(lib.q3dr_multi_score_viewpoints(multi._h, E._fp(K), W, 0, W, 0, H, E._fp(vps), vps.shape[0], C.byref(opts), C.byref(r)))
for _ in range(2):
    call()
ts = []
for _ in range(3):
    t = time.perf_counter(); call(); ts.append(time.perf_counter() - t)
rays = vps.shape[0] * W * H
print(json.dumps({"n_gpus": n_dev, "viewpoints": int(vps.shape[0]), "rays": rays, "ms": min(ts) * 1e3, "rays_per_sec_e2e": rays / min(ts),
                  "create_s": t_create, "note": "q3dr_multi_score_viewpoints: host poses in, per-GPU pinned host CSR results out, wall clock"}))
