"""Profiling driver for the batched search kernels: C2 map, one warm-up batch + one measured batch.

usage: python profiles/prof_search.py ALGO SEMANTICS NQ        (e.g. lazythetastar canonical 8192)
Run plain first, then under `ncu --set full --import-source on -k regex:search -s 1 -c 1`.
"""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("3d-astar-path-planners_b200")
synth = importlib.import_module("3d-astar-path-planners_b200.synth")


def main():
    algo, sem, nq = sys.argv[1], sys.argv[2], int(sys.argv[3])
    c = synth.CONFIGS["C2"]
    cloud = synth.pack_cloud(synth.config_cloud("C2", n_points=1_000_000), 16)
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(60.0, 60.0, 20.0),
                      ground_height=c["origin"][2])
    gm.cloudCallback(cloud, (0.0, 0.0, 1.0))
    s, g = synth.random_queries(6, gm.getInflateOccupancy, c["origin"], c["size"], nq, min_dist=5.0, z_range=(0.3, 3.0))
    pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=1000000, cost_weight=0.0, semantics=sem)
    pl.setGridMap(gm)
    dev = torch.device("cuda:0")
    ds, dg = torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
    dres = torch.empty(nq * 128, dtype=torch.uint8, device=dev)
    for it in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pl.findPaths(ds, dg, results=dres)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{algo}/{sem} batch {it}: {nq} queries {dt * 1e3:.2f} ms = {nq / dt:.0f} q/s launches {pl.last_launches()}")
    r = np.frombuffer(dres.cpu().numpy().tobytes(), dtype=b200.RESULT_DTYPE) if hasattr(b200, "RESULT_DTYPE") else None
    if r is not None:
        t = np.sort(r["time_us"])
        d = np.linalg.norm(s - g, axis=1)
        print("solved", int((r["status"] == 0).sum()), "expansions/query", float(r["iter_num"].mean()),
              "nodes>16384:", int((r["explored_nodes"] > 16384).sum()), "max nodes", int(r["explored_nodes"].max()))
        print("time_us: sum %.0f ms  p50 %.0f  p90 %.0f  p99 %.0f  p99.9 %.0f  max %.0f us" % (
            t.sum() / 1e3, t[len(t) // 2], t[int(len(t) * .9)], t[int(len(t) * .99)], t[int(len(t) * .999)], t[-1]))
        print("us per expansion (slot time / expansions): %.2f" % (r["time_us"].sum() / max(1, r["iter_num"].sum())))
        print("corr(time, dist) = %.3f   corr(iters, dist) = %.3f" % (np.corrcoef(r["time_us"], d)[0, 1], np.corrcoef(r["iter_num"], d)[0, 1]))


if __name__ == "__main__":
    main()
