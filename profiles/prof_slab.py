"""Single-GPU driver for the slab-shaped EDT (one z-slab of BASELINE config 5): 2048 x 2048 x nz with synthetic halo gaps;
prints the stage times of the fused kernels.  usage: prof_slab.py [nz] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import astar_planners_b200 as b200

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nxy, res = 2048, 0.1
dev = torch.device("cuda", 0)
size = (nxy * res, nxy * res, nz * res)
gm = b200.GridMap(size, res, origin=(-size[0] / 2, -size[1] / 2, -0.01), local_update_range=(1e4, 1e4, 1e4), ground_height=-0.01)
gm.set_slab_extent(512)
g = torch.Generator(device=dev); g.manual_seed(7)
n_pil = 11000
cx = (torch.rand(n_pil, generator=g, device=dev) - 0.5) * (size[0] - 2)
cy = (torch.rand(n_pil, generator=g, device=dev) - 0.5) * (size[1] - 2)
w = 0.5 + 0.7 * torch.rand(n_pil, generator=g, device=dev)
h = 0.8 + 2.2 * torch.rand(n_pil, generator=g, device=dev) + 40.0 * (torch.rand(n_pil, generator=g, device=dev) < 0.02)
u = torch.rand(n_pil, 1500, 3, generator=g, device=dev)
pts = torch.stack([cx[:, None] + (u[..., 0] - 0.5) * w[:, None], cy[:, None] + (u[..., 1] - 0.5) * w[:, None], u[..., 2] * h[:, None]], dim=-1).reshape(-1, 3)
gxy = (torch.arange(nxy, device=dev, dtype=torch.float32) + 0.5) * res - size[0] / 2
ground = torch.stack(torch.meshgrid(gxy, gxy, indexing="ij"), dim=-1).reshape(-1, 2)
ground = torch.cat([ground, torch.full((ground.shape[0], 1), 0.04, device=dev)], dim=1)
cloud = torch.zeros(pts.shape[0] + ground.shape[0], 4, device=dev, dtype=torch.float32)
cloud[:, :3] = torch.cat([pts, ground]).to(torch.float32)
gm.cloudCallback(cloud, (0.0, 0.0, 1.0), point_step=16)
above = torch.full((nxy, nxy), 7, dtype=torch.int32, device=dev)   # a neighbour slab 7 voxels above everywhere
gm.set_profiling(True)
for mode in (0, 1):
    gm.set_edt_pipeline(mode)
    ts = []
    for r in range(reps):
        gm.build_edt_slab(1.0, None, above)
        ts.append(gm.stage_ms())
    med = np.median(np.array(ts[1:]), axis=0)
    nv = nxy * nxy * nz
    tot = med[2:11].sum()
    print(f"nz={nz} {gm.edt_pipeline_used()}: stages ms {' '.join(f'{v:.2f}' for v in med[2:11])} total {tot:.2f} ms -> {nv / tot / 1e6:.1f} Gvox/s, {nv * 4.125 / tot / 1e6:.0f} GB/s (survey def)")
gm.close()
