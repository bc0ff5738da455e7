"""How many batches in flight?  PlannerPool throughput on the C2 map for 2 / 3 / 4 lanes, on the N = 1 bench batch and on the
shard rank 0 gets at N = 8 (the one with the 177 ms query).  usage: prof_lanes.py [algo]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import astar_planners_b200 as b200
from astar_planners_b200 import synth, shard
import bench

algo = sys.argv[1] if len(sys.argv) > 1 else "lazythetastar"
c, blob = bench.build_world(synth)
gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(60.0, 60.0, 20.0), ground_height=c["origin"][2])
gm.cloudCallback(torch.from_numpy(blob).cuda(), (0.0, 0.0, 1.0), point_step=16)
occ = lambda p: gm.getInflateOccupancy(p)
nq = 131072
dev = torch.device("cuda:0")
for label, world in (("N=1 batch", 1), ("rank 0 of 8", 8)):
    s_all, g_all = synth.random_queries(6, occ, c["origin"], c["size"], nq * world, min_dist=5.0, z_range=(0.3, 3.0))
    lo, hi = shard.query_shard(nq * world, 0, world)
    ds, dg = torch.from_numpy(np.ascontiguousarray(s_all[lo:hi])).to(dev), torch.from_numpy(np.ascontiguousarray(g_all[lo:hi])).to(dev)
    pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=1000000, semantics="canonical")
    pl.setGridMap(gm)
    dres = torch.empty(nq * 128, dtype=torch.uint8, device=dev)
    pl.findPaths(ds, dg, results=dres)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(); pl.findPaths(ds, dg, results=dres); b.record(); torch.cuda.synchronize()
    print(f"{algo} {label}: single batch {nq / a.elapsed_time(b) * 1e3:.0f} q/s")
    pl.close()
    for lanes, rounds in ((2, 3), (3, 2), (4, 2), (6, 1)):
        ms, q, same = bench.run_pipelined(b200, torch, gm, algo, "canonical", ds, dg, nq, dres, rounds=rounds, lanes=lanes)
        print(f"   lanes {lanes} x rounds {rounds}: {q / ms * 1e3:.0f} q/s identical={same}")
