import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, scenes
dev = torch.device("cuda:0")
def timeit(fn, iters=9, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
xyz = torch.from_numpy(scenes.scene_batch(range(32), 20000, "tabletop")).to(dev)
for (n, m) in ((20000, 2048), (2048, 1024), (1024, 512), (512, 256)):
    x = xyz[:, :n].contiguous()
    print(json.dumps({"n": n, "m": m, "fps_us": round(timeit(lambda: A.furthest_point_sampling(x, m)), 1),
                      "fps_xyz_us": round(timeit(lambda: A.furthest_point_sampling_xyz(x, m)), 1)}))
