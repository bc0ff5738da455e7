"""EDT driver for tuning / ncu: builds the C2 (or C3/C1) map from its synthetic cloud, checks the fused pipeline against
the seven-kernel one bit for bit, and prints per-stage CUDA-event times.  usage: prof_edt.py [config] [reps] [--ncu]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import astar_planners_b200 as b200
from astar_planners_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ncu = "--ncu" in sys.argv
c = synth.CONFIGS[cfg]
cloud = synth.pack_cloud(synth.config_cloud(cfg, n_points=1_000_000 if cfg in ("C2", "C4") else None))
gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(1e3, 1e3, 1e3), ground_height=c["origin"][2], device=0)
gm.cloudCallback(cloud, (0.0, 0.0, 1.0))
gm.set_profiling(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for mode in ((0,) if ncu else (0, 1)):
    gm.set_edt_pipeline(mode)
    ts = []
    for r in range(2 if ncu else reps):
        flush.zero_()
        torch.cuda.synchronize()
        gm.build_edt(1.0)
        ts.append(gm.stage_ms())
    ts = np.array(ts[1:])
    med = np.median(ts, axis=0)
    print(cfg, gm.edt_pipeline_used(), "stage ms (median):", " ".join(f"{v*1e3:.1f}" for v in med[2:11]), "-> EDT total us:", f"{med[2:11].sum()*1e3:.1f}")
    if not ncu:
        res[mode] = (gm.dist2().copy(), gm.costq().copy())
if not ncu:
    print("dist2 equal:", np.array_equal(res[0][0], res[1][0]), " costq equal:", np.array_equal(res[0][1], res[1][1]))
    if not np.array_equal(res[0][0], res[1][0]):
        bad = np.argwhere(res[0][0] != res[1][0])
        print("mismatches:", len(bad), bad[:5], res[0][0][tuple(bad[0])], res[1][0][tuple(bad[0])])
gm.close()
