#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path (contract: see the task statement / DESIGN.md §Measurement).

Default workload = BASELINE.json configs[1] ("C2"): one *frame* = rebuild the 512x512x128 map from a synthetic
1M-point PointCloud2: clear + voxelise/inflate scatter + exact squared EDT + quantised safety-cost field.
  value : ms/frame with the cloud already resident in HBM (CUDA events, L2 flushed between frames)
  e2e   : ms/frame through the host-facing C-ABI call (pinned host cloud -> H2D inside the timed region,
          local-bounds read-back D2H)
  extra : batched path queries/s on the same map (lazythetastar / astar / thetastar) and EDT GB/s
N > 1 (torchrun): the map is replicated (every rank rebuilds its replica — "replicas only" for the map
build), the query batch is sharded with no collective; max-over-ranks timing.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on the same frame.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_SAFE = 1.0
# (algorithm, semantics, queries per rank): faithful = bit-identical to the reference's std::set behaviour
# batch sizes are large enough that a reference-pathological query (faithful A* has a 1-in-2000 query with 100x the
# expansions) amortises, as it does in the 1M-query configuration C4
QUERY_ALGOS = (("lazythetastar", "canonical", 131072), ("astar", "faithful", 32768), ("astar", "canonical", 131072),
               ("thetastar", "faithful", 16384), ("thetastar", "canonical", 65536))


def build_world(synth):
    c = synth.CONFIGS["C2"]
    cloud = synth.config_cloud("C2", n_points=1_000_000)
    return c, synth.pack_cloud(cloud, 16)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions: NVML every ~2 ms from a thread
    (`nvidia-smi -lms 100` as the fallback: the timed region is only ~20 ms long)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc = index, uuid, [], None
        self.sm, self.reasons, self.max_mhz, self.nvml, self.stop_flag = [], set(), None, None, False
        # seconds between NVML polls; raised for host-facing regions (see main).  B200_BENCH_POLL_MS overrides (diagnostic).
        self.interval = float(os.environ.get("B200_BENCH_POLL_MS", "2")) * 1e-3

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                    "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def run_reference(args, rank, world, out):
    """CPU arm: the oracle's restatement of cloudCallback + (our) EDT/cost on all host threads."""
    if rank != 0:
        return
    from astar_planners_b200 import synth
    from oracle import pyoracle as po
    c, blob = build_world(synth)
    nthreads = po.max_threads()
    om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=(60.0, 60.0, 20.0),
                      ground_height=c["origin"][2])
    cam = (0.0, 0.0, 1.0)

    def frame():
        om.update_cloud(blob, cam, point_step=16, nthreads=nthreads)
        om.build_edt(d_safe=D_SAFE, nthreads=nthreads)
        om.local_bounds()

    for _ in range(max(1, min(args.warmup, 2))):
        frame()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    line = {"impl": "reference", "metric": "map rebuild ms/frame", "value": ms, "unit": "ms/frame", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "config": {"workload": "C2: 1M-point PointCloud2 -> 512x512x128 inflated occupancy + exact squared EDT + cost field",
                       "grid": [512, 512, 128], "points": 1000000, "d_safe_m": D_SAFE},
            "cpu_baseline": {"value": ms, "unit": "ms/frame", "cores": nthreads, "kind": "port",
                             "sample": f"{args.steps} full C2 frames (cloud update + EDT + cost), oracle/liboracle.so, {nthreads} threads"},
            "e2e": {"value": ms, "unit": "ms/frame", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


PARITY_FIELDS = ("solved", "status", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length")


def oracle_mode(po, sem):
    return po.FAITHFUL if sem == "faithful" else po.CANONICAL


def compare_with_oracle(po, b200, om, algo, sem, s, g, res, res_m, nthreads, cost_weight=0.0):
    """Untimed checker: the first len(s) result records of `res` against the CPU oracle on the same queries."""
    op = po.OraclePlanner(algo, om, res_m, 5.0, 1000000, mode=oracle_mode(po, sem), cost_weight=cost_weight)
    ro, _ = op.find_paths(s, g, nthreads=nthreads)
    bad = [f for f in PARITY_FIELDS if not np.array_equal(res[f][: len(s)], ro[f])]
    return bad


def run_slab_edt(b200, po, torch, dist, dev, rank, world, local_rank, nthreads, nxy=2048, nz_total=512, res=0.1):
    """BASELINE config 5: 2048 x 2048 x 512 exact EDT, z-slab sharded.  Per step: packed int16 column extents ->
    ONE all_gather (4 B per column per slab) -> gap derivation kernel -> two fused EDT kernels on the slab.
    Untimed checks: (1) every voxel of every slab == the unsharded transform built on the same GPU (all seam planes
    included); (2) the same sharded pipeline on 256 x 256 x 512 crops == the CPU lower-envelope oracle."""
    nz = nz_total // world
    size = (nxy * res, nxy * res, nz * res)
    origin = (-size[0] / 2, -size[1] / 2, -0.01 + rank * nz * res)
    gm = b200.GridMap(size, res, origin=origin, local_update_range=(1e4, 1e4, 1e4), ground_height=-0.01, device=local_rank)
    gm.set_slab_extent(nz_total)
    # device-generated forest (same seed on every rank): ground plane + pillars, ~2e7 points
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    n_pil = 11000
    cx = (torch.rand(n_pil, generator=g, device=dev) - 0.5) * (size[0] - 2)
    cy = (torch.rand(n_pil, generator=g, device=dev) - 0.5) * (size[1] - 2)
    w = 0.5 + 0.7 * torch.rand(n_pil, generator=g, device=dev)
    h = 0.8 + 2.2 * torch.rand(n_pil, generator=g, device=dev) + 40.0 * (torch.rand(n_pil, generator=g, device=dev) < 0.02)
    per = 1500
    u = torch.rand(n_pil, per, 3, generator=g, device=dev)
    pts = torch.stack([cx[:, None] + (u[..., 0] - 0.5) * w[:, None], cy[:, None] + (u[..., 1] - 0.5) * w[:, None],
                       u[..., 2] * h[:, None]], dim=-1).reshape(-1, 3)
    gxy = (torch.arange(nxy, device=dev, dtype=torch.float32) + 0.5) * res - size[0] / 2
    ground = torch.stack(torch.meshgrid(gxy, gxy, indexing="ij"), dim=-1).reshape(-1, 2)
    ground = torch.cat([ground, torch.full((ground.shape[0], 1), 0.04, device=dev)], dim=1)
    cloud = torch.zeros(pts.shape[0] + ground.shape[0], 4, device=dev, dtype=torch.float32)
    cloud[:, :3] = torch.cat([pts, ground]).to(torch.float32)
    del u, pts, ground
    gm.cloudCallback(cloud, (0.0, 0.0, 1.0), point_step=16)

    def make_step(m, n_xy):
        ext = torch.empty(n_xy, n_xy, 2, dtype=torch.int16, device=dev)
        allext = torch.empty(world, n_xy, n_xy, 2, dtype=torch.int16, device=dev)
        below = torch.empty(n_xy, n_xy, dtype=torch.int32, device=dev)
        above = torch.empty(n_xy, n_xy, dtype=torch.int32, device=dev)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

        def step():
            evs[0].record()
            m.column_extents16(ext)
            evs[1].record()
            dist.all_gather_into_tensor(allext.view(torch.int32), ext.view(torch.int32))  # NCCL has no int16: (lo, hi) travels as one int32
            evs[2].record()
            m.slab_gaps(allext, [nz] * world, rank, below, above)
            evs[3].record()
            m.build_edt_slab(D_SAFE, below if rank > 0 else None, above if rank < world - 1 else None)
            evs[4].record()
        return step, evs, (ext, allext, below, above)

    step, evs, keep = make_step(gm, nxy)
    gm.set_profiling(True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    reps = 5
    stage = np.zeros((reps, 6))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    for i in range(3):  # stage split of three more (separately synchronised) steps
        step()
        torch.cuda.synchronize()
        sm = gm.stage_ms()
        stage[i, :4] = [evs[k].elapsed_time(evs[k + 1]) for k in range(4)]
        stage[i, 4:6] = sm[9:11] if gm.edt_pipeline_used() == "fused" else [sum(sm[2:6]), sum(sm[6:9])]
    stage_ms = np.median(stage[:3], axis=0)
    nvox_total = nxy * nxy * nz_total
    out = {"grid": [nxy, nxy, nz_total], "slabs": world, "points": int(cloud.shape[0]), "ms": float(ms),
           "pipeline": gm.edt_pipeline_used(),
           "stage_ms_rank0": {"column_extents16": float(stage_ms[0]), "all_gather(nccl)": float(stage_ms[1]), "slab_gaps": float(stage_ms[2]),
                              "build_edt_slab": float(stage_ms[3]), "edt_pass_zy": float(stage_ms[4]), "edt_pass_x": float(stage_ms[5])},
           "gbs_survey_def_total": (nvox_total / 8 + 4 * nvox_total) / (float(ms) * 1e-3) / 1e9,
           "gbs_survey_def_per_gpu": (nvox_total / 8 + 4 * nvox_total) / world / (float(ms) * 1e-3) / 1e9,
           "exchange_bytes_per_rank": int(world * nxy * nxy * 4)}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        out["hbm_frac_per_gpu"] = out["gbs_survey_def_per_gpu"] / float(json.load(open(peaks_path))["hbm_gbs"])

    # ---- check 1 (untimed): every voxel of this rank's slab == the unsharded 2048 x 2048 x 512 transform on this GPU
    errors = []
    full = b200.GridMap((size[0], size[1], nz_total * res), res, origin=(origin[0], origin[1], -0.01),
                        local_update_range=(1e4, 1e4, 1e4), ground_height=-0.01, device=local_rank)
    full.cloudCallback(cloud, (0.0, 0.0, 1.0), point_step=16)
    full.build_edt(D_SAFE)
    fd, fc = full.device_fields()
    sd, sc = gm.device_fields()
    z0 = rank * nz
    same_d = bool(torch.equal(fd[:, :, z0:z0 + nz], sd))
    same_c = bool(torch.equal(fc[:, :, z0:z0 + nz], sc))
    seam_planes = [z for z in (z0, z0 + nz - 1)]
    seams_ok = all(bool(torch.equal(fd[:, :, z], sd[:, :, z - z0])) for z in seam_planes)
    finite = int((sd < (1 << 29)).sum().item())
    far = int(sd.max().item())
    if not (same_d and same_c and seams_ok):
        errors.append(f"rank {rank}: slab differs from the unsharded transform (dist2 {same_d}, costq {same_c}, seams {seams_ok})")
    del fd, fc, sd, sc
    full.close()
    gm.close()
    del keep, cloud
    torch.cuda.empty_cache()

    # ---- check 2 (untimed): the same sharded pipeline on 256 x 256 x 512 crops of the world against the CPU lower envelope
    crops, n_crop = 4, 256
    crop_bad = 0
    rng = np.random.default_rng(5)
    for c in range(crops):
        ox, oy = (rng.integers(0, nxy - n_crop, 2) * res - nxy * res / 2).tolist()
        csize = (n_crop * res, n_crop * res, nz * res)
        cm = b200.GridMap(csize, res, origin=(ox, oy, -0.01 + rank * nz * res), local_update_range=(1e4, 1e4, 1e4), ground_height=-0.01,
                          device=local_rank)
        cm.set_slab_extent(nz_total)
        # a small forest of its own (seeded per crop, identical on every rank), including columns that are empty in whole slabs
        gc = torch.Generator(device=dev)
        gc.manual_seed(100 + c)
        npc = 60
        pcx = ox + torch.rand(npc, generator=gc, device=dev) * csize[0]
        pcy = oy + torch.rand(npc, generator=gc, device=dev) * csize[1]
        ph = 0.5 + torch.rand(npc, generator=gc, device=dev) * (nz_total * res - 1.0) * (torch.rand(npc, generator=gc, device=dev) < 0.5)
        uu = torch.rand(npc, 4000, 3, generator=gc, device=dev)
        pp = torch.stack([pcx[:, None] + (uu[..., 0] - 0.5) * 0.8, pcy[:, None] + (uu[..., 1] - 0.5) * 0.8, uu[..., 2] * ph[:, None]], dim=-1).reshape(-1, 3)
        # floating obstacles (not connected to the ground): what makes the below / above gaps differ per column
        fl = torch.rand(300, 3, generator=gc, device=dev) * torch.tensor([csize[0], csize[1], nz_total * res - 0.2], device=dev) + \
            torch.tensor([ox, oy, 0.0], device=dev)
        cc = torch.zeros(pp.shape[0] + fl.shape[0], 4, device=dev, dtype=torch.float32)
        cc[:, :3] = torch.cat([pp, fl]).to(torch.float32)
        cm.cloudCallback(cc, (0.0, 0.0, 1.0), point_step=16)
        cstep, _, ckeep = make_step(cm, n_crop)
        cstep()
        torch.cuda.synchronize()
        d_slab, _ = cm.device_fields()
        occ_slab = torch.from_numpy(cm.inflate()).to(dev)
        all_d = [torch.empty_like(d_slab) for _ in range(world)] if rank == 0 else None
        all_o = [torch.empty_like(occ_slab) for _ in range(world)] if rank == 0 else None
        dist.gather(d_slab.contiguous(), all_d, dst=0)
        dist.gather(occ_slab, all_o, dst=0)
        if rank == 0:
            grid = torch.cat(all_o, dim=2).cpu().numpy()
            got = torch.cat(all_d, dim=2).cpu().numpy()
            om = po.OracleMap((ox, oy, -0.01), (csize[0], csize[1], nz_total * res), res)
            om.set_inflate(grid)
            want, _ = om.build_edt(d_safe=0.0, nthreads=nthreads)
            if not np.array_equal(got, want):
                crop_bad += 1
                errors.append(f"C5 crop {c}: sharded EDT differs from the CPU lower envelope in {int((got != want).sum())} voxels")
            del om
        cm.close()
        del ckeep
    out["parity"] = {"slab_equals_unsharded_all_voxels": bool(same_d and same_c), "seam_planes_checked": seam_planes,
                     "finite_voxels_in_slab": finite, "max_dist2_in_slab": far,
                     "cpu_oracle_crops": {"crops": crops, "grid": [n_crop, n_crop, nz_total], "column_blocks_32x32": crops * 64,
                                          "slabs": world, "mismatching_crops": crop_bad} if rank == 0 else None}
    return out, errors


def run_c1(b200, po, synth, device, errors):
    """BASELINE config C1: one Theta* query (0,0,0)->(40,40,4) on the 250x250x50 @0.2 m forest; GPU latency of the
    host-facing findPath next to the single-threaded CPU restatement (the reference itself is single-threaded)."""
    c = synth.CONFIGS["C1"]
    cloud = synth.pack_cloud(synth.config_cloud("C1"))
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(1e3, 1e3, 1e3), ground_height=c["origin"][2], device=device)
    gm.cloudCallback(cloud, (0.0, 0.0, 1.0))
    om = None
    if po is not None:
        om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=(1e3, 1e3, 1e3), ground_height=c["origin"][2])
        om.update_cloud(cloud, (0.0, 0.0, 1.0))
    out = {"query": "(0,0,0)->(40,40,4)", "grid": list(gm.dims), "resolution": 0.2}
    s, g = (0.0, 0.0, 0.0), (40.0, 40.0, 4.0)
    for algo in ("thetastar", "astar", "lazythetastar"):
        for sem in (("faithful", "canonical") if algo != "lazythetastar" else ("canonical",)):
            pl = b200.Planner(algo, resolution=0.2, lambda_heuristic=5.0, allocate_num=1000000, semantics=sem)
            pl.setGridMap(gm)
            pl.findPath(s, g)
            ts, r = [], None
            for _ in range(7):
                gm.sync()
                t0 = time.perf_counter()
                r = pl.findPath(s, g)
                ts.append((time.perf_counter() - t0) * 1e3)
            e = {"gpu_ms_per_query_e2e": float(np.median(ts)), "gpu_kernel_ms": r["time_spent"] * 1e-3, "solved": bool(r["solved"]),
                 "expansions": r["iter_num"], "explored_nodes": r["explored_nodes"], "los_checks": r["line_of_sight_checks"],
                 "los_samples": r["los_samples"], "path_length": r["path_length"], "n_path": len(pl.getPath())}
            if om is not None:
                op = po.OraclePlanner(algo, om, 0.2, 5.0, 1000000, mode=oracle_mode(po, sem))
                tc, ro = [], None
                for _ in range(3):
                    t0 = time.perf_counter()
                    ro, pth = op.find_path(s, g)
                    tc.append((time.perf_counter() - t0) * 1e3)
                e["cpu_1_thread_ms_per_query"] = float(np.median(tc))
                same = (int(ro["explored_nodes"]) == r["explored_nodes"] and float(ro["g_final"]) == r["g_final_node"] and
                        float(ro["path_length"]) == r["path_length"] and np.array_equal(pth, pl.getPath()))
                e["identical_to_oracle"] = bool(same)
                if not same:
                    errors.append(f"C1 {algo}/{sem}: result differs from the oracle")
            out[f"{algo}_{sem}"] = e
            pl.close()
    gm.close()
    return out


def run_c3(b200, po, synth, torch, dev, device, host_threads, errors):
    """BASELINE config C3: 10 000 random costlazythetastar queries with safety cost on the 256x256x64 forest (+ thetastar,
    the reference-pinned LoS-heavy algorithm, on the same queries)."""
    c = synth.CONFIGS["C3"]
    cloud = synth.pack_cloud(synth.config_cloud("C3"))
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(1e3, 1e3, 1e3), ground_height=c["origin"][2], device=device)
    gm.cloudCallback(cloud, (0.0, 0.0, 1.0))
    gm.build_edt(1.0)
    om = None
    if po is not None:
        om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=(1e3, 1e3, 1e3), ground_height=c["origin"][2])
        om.update_cloud(cloud, (0.0, 0.0, 1.0))
        om.build_edt(1.0, nthreads=host_threads)
    nq = 10000
    s, g = synth.random_queries(5, gm.getInflateOccupancy, c["origin"], c["size"], nq)
    ds, dg = torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
    out = {"grid": list(gm.dims), "queries": nq, "d_safe_m": 1.0, "cost_weight": 1.0}
    for algo, sem in (("costlazythetastar", "canonical"), ("thetastar", "faithful"), ("thetastar", "canonical")):
        pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=1000000, cost_weight=1.0, semantics=sem)
        pl.setGridMap(gm)
        dres = torch.empty(nq * 128, dtype=torch.uint8, device=dev)
        pl.findPaths(ds, dg, results=dres)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pl.findPaths(ds, dg, results=dres)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        t0 = time.perf_counter()
        res, _ = pl.findPaths(s, g)
        ms_h = (time.perf_counter() - t0) * 1e3
        e = {"queries_per_s": nq / (ms * 1e-3), "queries_per_s_e2e": nq / (ms_h * 1e-3), "solved": int(res["solved"].sum()),
             "expansions_per_query": float(res["iter_num"].mean()), "los_checks_per_query": float(res["los_checks"].mean()),
             "los_samples_per_query": float(res["los_samples"].mean()), "query_time": percentile_block(res["time_us"])}
        if om is not None:
            op = po.OraclePlanner(algo, om, 0.1, 5.0, 1000000, mode=oracle_mode(po, sem), cost_weight=1.0)
            n_nt, n_1t = 2048, 128
            t0 = time.perf_counter()
            ro, _ = op.find_paths(s[:n_nt], g[:n_nt], nthreads=host_threads)
            t_nt = time.perf_counter() - t0
            t0 = time.perf_counter()
            op.find_paths(s[:n_1t], g[:n_1t], nthreads=1)
            t_1t = time.perf_counter() - t0
            bad = [f for f in PARITY_FIELDS if not np.array_equal(res[f][:n_nt], ro[f])]
            e.update({"cpu_queries_per_s_all_threads": n_nt / t_nt, "cpu_threads": host_threads, "cpu_queries_per_s_1_thread": n_1t / t_1t,
                      "cpu_sample": [n_nt, n_1t], "identical_to_oracle_on_sample": not bad})
            if bad:
                errors.append(f"C3 {algo}/{sem}: fields {bad} differ from the oracle")
        out[f"{algo}_{sem}"] = e
        pl.close()
    gm.close()
    return out


def run_pipelined(b200, torch, gm, algo, sem, ds, dg, nq, dres_single, rounds=3, lanes=2):
    """Steady-state serving form of the batched planners (`b200.PlannerPool`): `lanes` planner handles on the same map, each
    with its own CUDA stream and host thread, each running `rounds` batches back to back.  A batch's last straggling queries
    then run beside the next batch's first queries instead of leaving the SMs idle.  Device-timed: one start event, one end
    event per lane, the longest span counts.  Returns (ms, queries, identical) — identical: every batch's result block equals
    the single-batch result bit for bit (status / lengths / costs; the per-query clock field excluded)."""
    pool = b200.PlannerPool(gm, algo, lanes=lanes, resolution=0.1, lambda_heuristic=5.0, allocate_num=1000000, cost_weight=0.0, semantics=sem)
    outs = [torch.empty(nq * 128, dtype=torch.uint8, device=ds.device) for _ in range(lanes * rounds)]
    pool.find_paths_many([(ds, dg, outs[i]) for i in range(lanes)])  # warm-up: workspace allocation of every lane
    torch.cuda.synchronize()
    lane_streams = [torch.cuda.ExternalStream(st, device=ds.device) for st in pool.streams]  # the pool's own cudaStream_t's
    start = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in lane_streams]
    start.record()
    for st in lane_streams:
        st.wait_event(start)
    pool.find_paths_many([(ds, dg, outs[i]) for i in range(lanes * rounds)])
    for e, st in zip(ends, lane_streams):
        e.record(st)
    torch.cuda.synchronize()
    ms = max(start.elapsed_time(e) for e in ends)
    ref = dres_single.cpu().numpy().view(b200.RESULT_DTYPE)
    same = True
    for out in outs:
        got = out.cpu().numpy().view(b200.RESULT_DTYPE)
        for f in ("solved", "status", "n_path", "explored_nodes", "iter_num", "g_final", "path_length"):
            same = same and np.array_equal(ref[f], got[f])
    pool.close()
    return ms, lanes * rounds * nq, bool(same)


def percentile_block(t_us):
    t = np.asarray(t_us, dtype=np.float64)
    return {"p50_us": float(np.percentile(t, 50)), "p99_us": float(np.percentile(t, 99)), "max_us": float(t.max()),
            "max_over_p99": float(t.max() / max(np.percentile(t, 99), 1e-9))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the query-throughput extras")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout must carry exactly ONE JSON line: send everything libraries print (e.g. the NCCL version banner)
    # to stderr and keep the real stdout for the final line
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist
    import astar_planners_b200 as b200
    from astar_planners_b200 import shard, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    po = None
    if not args.no_cpu or world > 1:
        from oracle import pyoracle as po  # the checker / CPU baseline; never on the measured path
    host_threads = po.max_threads() if po else (os.cpu_count() or 1)
    check_threads = max(1, host_threads // world)  # every rank runs its own checker at N > 1
    errors = []  # parity failures found by the untimed checks: reported in the line AND turned into a non-zero exit code

    c, blob = build_world(synth)
    n_pts = blob.shape[0]
    cam = (0.0, 0.0, 1.0)
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(60.0, 60.0, 20.0),
                      ground_height=c["origin"][2], device=local_rank)
    nvox = int(np.prod(gm.dims))
    d_cloud = torch.from_numpy(blob).to(dev)
    h_cloud = torch.from_numpy(blob).pin_memory()
    h_cloud_np = h_cloud.numpy()
    pageable_np = blob.copy()  # what cloudCallback(const sensor_msgs::PointCloud2ConstPtr&) really hands over: pageable memory
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def frame_device():
        gm.cloudCallback(d_cloud, cam, point_step=16)
        gm.build_edt(D_SAFE)

    def frame_host(buf):
        gm.cloudCallback(buf, cam, point_step=16)
        gm.build_edt(D_SAFE)
        return gm.local_bounds()  # D2H read-back of the frame's result (6 x int64 keys)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.start()  # runs through warm-up, the timed region and the e2e region (the timed region alone is ~15 ms)
    gm.set_profiling(True)
    for _ in range(args.warmup):
        flush.fill_(1)
        frame_device()
    torch.cuda.synchronize()
    fused = gm.edt_pipeline_used() == "fused"

    # ---- timed region: EXACTLY `steps` frames, L2 flushed (untimed) before each, CUDA events per frame
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = np.zeros((args.steps, 16))
    t_wall = time.perf_counter()
    launches0 = gm.launch_count()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        ev[i][0].record()
        frame_device()
        ev[i][1].record()
        stage[i] = gm.stage_ms()  # syncs the stream (outside the event bracket)
    gpu_launches = gm.launch_count() - launches0  # counted by the library at every kernel launch it issues
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    per = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(per.sum())
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps

    # ---- e2e: host cloud through the C-ABI, H2D inside, bounds read back (pinned = the headline; pageable beside it)
    def time_e2e(buf):
        # Untimed warm-up until the frame time has settled: on these (virtualised) boxes the first ~50 DMA reads of a freshly
        # pinned buffer run at 15-25 GB/s and only then at the link's 54 GB/s (profiles/diag_h2d.py), so three warm-up frames
        # would time the host's page-mapping warm-up, not the path.  At least max(warmup, 5) frames, at most 200 (~0.2 s).
        if buf is h_cloud_np:
            # plain copies first (they warm the buffer's DMA path without running the frame): until the copy time has been
            # flat for 20 copies and at least 0.5 s have passed — the slow state has been seen to outlast 100 transfers — 4 s at most
            ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cms, t_start = [], time.perf_counter()
            while True:
                ca.record()
                d_cloud.copy_(h_cloud, non_blocking=True)
                cb.record()
                torch.cuda.synchronize()
                cms.append(ca.elapsed_time(cb))
                el = time.perf_counter() - t_start
                if el > 4.0 or (el > 0.5 and len(cms) >= 40 and max(cms[-20:]) <= 1.03 * min(cms)):
                    break
        hist = []
        for i in range(200):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            frame_host(buf)
            torch.cuda.synchronize()
            hist.append(time.perf_counter() - t0)
            if i + 1 >= max(args.warmup, 20) and max(hist[-8:]) <= 1.03 * min(hist):
                break
        barrier()
        ts = []
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            frame_host(buf)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        tot = float(np.sum(ts))
        if world > 1:
            t = torch.tensor([tot], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t.item())
        return tot / args.steps, float(np.median(ts))

    clocks = sampler.stop()

    # ---- roofline (per-launch CUDA-event durations from the timed region).
    # The dominant unit of the frame is the EDT stage.  With the fused pipeline it is TWO kernels of nearly equal length
    # coupled through a 2 B/voxel intermediate that is designed to stay in L2, so neither kernel's own DRAM bytes mean much
    # alone (ncu, which flushes L2 around every kernel, sees 19 MB for the first and misses the write-back of the second):
    # the roofline entry is therefore the STAGE, with SURVEY 8d's algorithmic bytes (N/8 bit grid in, 4N dist2 out) plus
    # the 2N costq the stage also writes; intermediates, stacks and scratch are never counted.  `traffic` = the two
    # kernels' ncu DRAM bytes summed (profiles/traffic.json); per-kernel interface bytes and times are listed beside it.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    st = stage.mean(axis=0)
    if fused:
        kernels = {  # name: (ms, bytes the kernel reads + writes through its interface)
            "clear(memset)": (st[0], nvox / 8),
            "scatter_kernel": (st[1], 16 * n_pts + nvox / 8),
            "edt_fused_y_kernel": (st[9], nvox / 8 + 2 * nvox),               # bit grid in, 16-bit (dy, dz) intermediate out
            "edt_fused_x_kernel": (st[10], 2 * nvox + 4 * nvox + 2 * nvox),   # intermediate in, dist2 (int32) + costq (uint16) out
        }
        edt_names = ["edt_fused_y_kernel", "edt_fused_x_kernel"]
        edt_ms = float(st[9] + st[10])
    else:
        kernels = {
            "clear(memset)": (st[0], nvox / 8), "scatter_kernel": (st[1], 16 * n_pts + nvox / 8),
            "edt_z_kernel": (st[2], nvox / 8 + nvox), "edt_scan_y_kernel": (st[3], nvox + 8 * nvox), "edt_cross_kernel(y)": (st[4], 0),
            "edt_fill_y_kernel": (st[5], 8 * nvox + 4 * nvox), "edt_scan_x_kernel": (st[6], 4 * nvox + 8 * nvox), "edt_cross_kernel(x)": (st[7], 0),
            "edt_fill_x_kernel": (st[8], 8 * nvox + 4 * nvox + 2 * nvox),
        }
        edt_names = [k for k in kernels if k.startswith("edt_")]
        edt_ms = float(st[2:9].sum())
    stage_bytes = nvox / 8 + 4 * nvox + 2 * nvox
    achieved = stage_bytes / (edt_ms * 1e-3) / 1e9 if edt_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "EDT stage = " + " + ".join(edt_names), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": stage_bytes,
                "kernel_ms": edt_ms, "share_of_frame": edt_ms / ms_per_step,
                "algorithmic_bytes_definition": "SURVEY 8d: N/8 (bit grid in) + 4N (int32 dist2 out), + 2N (uint16 costq out); "
                                                "intermediates, stacks and scratch are not counted",
                "per_kernel": {k: {"ms": float(v[0]), "interface_bytes": float(v[1]),
                                   "interface_gbs": float(v[1] / (v[0] * 1e-3) / 1e9) if v[0] > 0 else 0.0} for k, v in kernels.items()}}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path))
        per = {k: tr.get(k) for k in kernels if tr.get(k) is not None}
        roofline["traffic_per_kernel"] = per
        if all(k in per for k in edt_names):
            roofline["traffic"] = float(sum(per[k] for k in edt_names))
            roofline["algorithmic_le_traffic"] = bool(stage_bytes <= roofline["traffic"])
    # the stage figures SURVEY 8d / BASELINE.json's "EDT HBM GB/s" name
    edt_bytes_survey = nvox / 8 + 4 * nvox
    edt_bytes_with_cost = edt_bytes_survey + 2 * nvox
    frame_bytes = (nvox / 8 + 16 * n_pts + nvox / 8) + edt_bytes_with_cost
    extra = {"edt_pipeline": "fused (2 kernels, hull stacks in shared memory)" if fused else "hbm-stack (7 kernels)",
             "edt_ms": float(edt_ms),
             "edt_hbm_gbs_survey_def": edt_bytes_survey / (edt_ms * 1e-3) / 1e9 if edt_ms > 0 else 0.0,
             "edt_hbm_frac_survey_def": edt_bytes_survey / (edt_ms * 1e-3) / 1e9 / peak if edt_ms > 0 else 0.0,
             "edt_hbm_gbs_with_costq": edt_bytes_with_cost / (edt_ms * 1e-3) / 1e9 if edt_ms > 0 else 0.0,
             "frame_hbm_frac": frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
             "wall_ms_timed_region": wall_ms,
             }
    if world > 1:
        extra["multi_gpu_note"] = ("map build: replicas only — every rank rebuilds the same replica, so ms/frame is flat in N by "
                                   "construction; what scales is the sharded query throughput in extra.scaling")

    # ---- extras: batched query throughput on this map (sharded over ranks, no collective)
    scaling = {}
    qsampler = None
    if not args.no_extra:
        # every rank samples ITS GPU's clocks over the query phase: per-rank rates differ when eight GPUs of one box do
        # not hold the same clock under load, and the scaling block should show that next to the rates
        qsampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
        qsampler.interval = 0.02  # the query batches last 0.1-3 s each
        qsampler.start()
        occ = lambda p: gm.getInflateOccupancy(p)
        gm.build_edt(D_SAFE)
        om = None
        if po is not None:
            om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=(60.0, 60.0, 20.0), ground_height=c["origin"][2])
            om.update_cloud(blob, cam, point_step=16, nthreads=check_threads)
        # fastLOS batch: 1M segments between random free points (device-resident in/out)
        nseg = 1 << 20
        sa, sb = synth.random_queries(60 + rank, occ, c["origin"], c["size"], nseg, min_dist=1.0, z_range=(0.3, 3.0))
        da, db = torch.from_numpy(sa).to(dev), torch.from_numpy(sb).to(dev)
        vis = torch.empty(nseg, dtype=torch.uint8, device=dev)
        nsmp = torch.empty(nseg, dtype=torch.int32, device=dev)
        gm.los_batch(da, db, vis_out=vis, samples_out=nsmp)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gm.los_batch(da, db, vis_out=vis, samples_out=nsmp)
        b.record()
        torch.cuda.synchronize()
        los_ms = a.elapsed_time(b)
        t = torch.tensor([los_ms, float(nseg), float(nsmp.sum().item())], device=dev, dtype=torch.float64)
        if world > 1:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            los_ms = float(mx[0])
        extra["fastlos_segments_per_s"] = float(t[1]) / (los_ms * 1e-3)
        extra["fastlos_samples_per_s"] = float(t[2]) / (los_ms * 1e-3)
        extra["fastlos_visible_frac"] = float(vis.float().mean().item())
        if om is not None:  # untimed check: 20 000 of the segments against the CPU restatement of fastLOS
            if not np.array_equal(vis[:20000].cpu().numpy(), om.los_batch(sa[:20000], sb[:20000], nthreads=check_threads)):
                errors.append(f"rank {rank}: fastLOS booleans differ from the oracle")
        # the same segments through the voxel-traversal line of sight (rank-local number)
        vis2 = torch.empty(nseg, dtype=torch.uint8, device=dev)
        ncell = torch.empty(nseg, dtype=torch.int32, device=dev)
        gm.los_batch_dda(da, db, vis_out=vis2, cells_out=ncell)
        torch.cuda.synchronize()
        a.record()
        gm.los_batch_dda(da, db, vis_out=vis2, cells_out=ncell)
        b.record()
        torch.cuda.synchronize()
        dda_ms = a.elapsed_time(b)
        extra["ddalos_segments_per_s_per_gpu"] = nseg / (dda_ms * 1e-3)
        extra["ddalos_cells_per_s_per_gpu"] = float(ncell.sum().item()) / (dda_ms * 1e-3)
        extra["ddalos_agrees_with_fastlos_frac"] = float((vis2 == vis).float().mean().item())
        del da, db, vis, vis2, nsmp, ncell

        cpu_q = {}
        for algo, sem, nq in QUERY_ALGOS:
            # every rank draws the same global batch and takes its contiguous shard (map replicated, no collective)
            s_all, g_all = synth.random_queries(6, occ, c["origin"], c["size"], nq * world, min_dist=5.0, z_range=(0.3, 3.0))
            lo, hi = shard.query_shard(nq * world, rank, world)
            s, g = np.ascontiguousarray(s_all[lo:hi]), np.ascontiguousarray(g_all[lo:hi])
            pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=1000000, cost_weight=0.0, semantics=sem)
            pl.setGridMap(gm)
            ds, dg = torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
            dres = torch.empty(nq * 128, dtype=torch.uint8, device=dev)
            pl.findPaths(ds, dg, results=dres)  # warm-up at full size (workspace allocation, both tiers)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pl.findPaths(ds, dg, results=dres)
            b.record()
            torch.cuda.synchronize()
            ms_own = a.elapsed_time(b)
            t0 = time.perf_counter()
            res, _ = pl.findPaths(s, g)  # host-facing (H2D + D2H inside)
            ms_h = (time.perf_counter() - t0) * 1e3
            tot = torch.tensor([ms_own, ms_h, float(nq), float(res["solved"].sum()), float(res["iter_num"].sum())], device=dev,
                               dtype=torch.float64)
            ms, own_rates = ms_own, [nq / (ms_own * 1e-3)]
            if world > 1:
                mx = tot.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
                ms, ms_h = float(mx[0]), float(mx[1])
                rates = torch.zeros(world, device=dev, dtype=torch.float64)
                rates[rank] = own_rates[0]
                dist.all_reduce(rates, op=dist.ReduceOp.SUM)
                own_rates = [float(x) for x in rates.cpu()]
            key = f"{algo}_{sem}"
            qps = float(tot[2]) / (ms * 1e-3)
            extra[f"{key}_queries_per_s"] = qps
            extra[f"{key}_queries_per_s_e2e"] = float(tot[2]) / (ms_h * 1e-3)
            extra[f"{key}_queries"] = int(tot[2])
            extra[f"{key}_solved"] = int(tot[3])
            extra[f"{key}_expansions_per_query"] = float(tot[4]) / float(tot[2])
            extra[f"{key}_query_time"] = percentile_block(res["time_us"])
            # weak scaling as measured in THIS run: aggregate rate against n_gpus x the mean rate a rank reaches on its own shard
            scaling[key] = {"queries_per_s": qps, "per_rank_queries_per_s": own_rates,
                            "efficiency_vs_mean_rank_rate": qps / (world * float(np.mean(own_rates))),
                            "slowest_over_fastest_rank_time": float(max(own_rates) / min(own_rates))}
            # ---- the same batches in steady-state serving form: several planner handles / streams / host threads on the one map,
            # so one batch's stragglers overlap other batches' work.  Canonical: 4 in flight, 2 batches each (two lanes hide the
            # tail of an ordinary batch, four also that of a shard whose longest query alone outlasts the rest of its batch).
            # Faithful: 4 in flight, 1 each — there a batch lasts as long as its one reference-pathological query while 90 % of
            # the slots idle.
            if True:
                p_lanes, p_rounds = (4, 2) if sem == "canonical" else (4, 1)  # lanes measured in profiles/prof_lanes.py
                p_ms, p_q, p_same = run_pipelined(b200, torch, gm, algo, sem, ds, dg, nq, dres, rounds=p_rounds, lanes=p_lanes)
                own_p = p_q / (p_ms * 1e-3)
                p_rates = [own_p]
                if world > 1:
                    mx = torch.tensor([p_ms], device=dev, dtype=torch.float64)
                    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                    p_ms = float(mx[0])
                    rr_ = torch.zeros(world, device=dev, dtype=torch.float64)
                    rr_[rank] = own_p
                    dist.all_reduce(rr_, op=dist.ReduceOp.SUM)
                    p_rates = [float(x) for x in rr_.cpu()]
                p_qps = p_q * world / (p_ms * 1e-3)
                extra[f"{key}_queries_per_s_pipelined"] = p_qps
                scaling[key]["pipelined"] = {"queries_per_s": p_qps, "per_rank_queries_per_s": p_rates, "batches_in_flight": p_lanes,
                                             "batches_per_lane": p_rounds, "queries": int(p_q * world),
                                             "efficiency_vs_mean_rank_rate": p_qps / (world * float(np.mean(p_rates))),
                                             "over_single_batch_rate": p_qps / qps}
                if not p_same:
                    errors.append(f"rank {rank}: {key}: pipelined batches differ from the single-batch results")
            # ---- untimed multi-GPU correctness: (a) a slice of this rank's shard against the CPU oracle; (b) rank 0 recomputes
            # a slice of EVERY rank's shard on its own GPU and compares it bit for bit with what that rank reported
            nchk = 256
            if om is not None and (world > 1 or not args.no_cpu):
                bad = compare_with_oracle(po, b200, om, algo, sem, s[:nchk], g[:nchk], res, 0.1, check_threads)
                if bad:
                    errors.append(f"rank {rank}: {key}: fields {bad} differ from the oracle on its shard")
            if world > 1:
                mine = np.zeros((nchk, 4), dtype=np.float64)
                for j, f in enumerate(("g_final", "path_length", "explored_nodes", "n_path")):
                    mine[:, j] = res[f][:nchk]
                allr = torch.zeros(world, nchk, 4, device=dev, dtype=torch.float64)
                allr[rank] = torch.from_numpy(mine).to(dev)
                dist.all_reduce(allr, op=dist.ReduceOp.SUM)
                if rank == 0:
                    for r in range(1, world):
                        rlo, _ = shard.query_shard(nq * world, r, world)
                        rr, _ = pl.findPaths(np.ascontiguousarray(s_all[rlo:rlo + nchk]), np.ascontiguousarray(g_all[rlo:rlo + nchk]))
                        want = np.stack([rr[f].astype(np.float64) for f in ("g_final", "path_length", "explored_nodes", "n_path")], 1)
                        if not np.array_equal(want, allr[r].cpu().numpy()):
                            errors.append(f"{key}: rank {r}'s shard results differ from rank 0's recomputation")
            # ---- CPU baseline for this configuration (rank 0 at N = 1): oracle, all threads on 4096 queries, one thread on 256
            if rank == 0 and world == 1 and not args.no_cpu:
                op = po.OraclePlanner(algo, om, 0.1, 5.0, 1000000, mode=oracle_mode(po, sem))
                n_nt, n_1t = min(4096, nq), 256
                t0 = time.perf_counter()
                op.find_paths(s[:n_nt], g[:n_nt], nthreads=host_threads)
                t_nt = time.perf_counter() - t0
                t0 = time.perf_counter()
                op.find_paths(s[:n_1t], g[:n_1t], nthreads=1)
                t_1t = time.perf_counter() - t0
                cpu_q[key] = {"queries_per_s_all_threads": n_nt / t_nt, "threads": host_threads, "sample_all_threads": n_nt,
                              "queries_per_s_1_thread": n_1t / t_1t, "sample_1_thread": n_1t,
                              "gpu_over_cpu_all_threads": qps / (n_nt / t_nt)}
            pl.close()
            del ds, dg, dres
        qclk = qsampler.stop()
        if world > 1:
            allclk = [None] * world
            dist.all_gather_object(allclk, {"sm_mhz": qclk.get("sm_mhz"), "sm_min_mhz": qclk.get("sm_min_mhz"), "reasons": qclk.get("reasons")})
        else:
            allclk = [{"sm_mhz": qclk.get("sm_mhz"), "sm_min_mhz": qclk.get("sm_min_mhz"), "reasons": qclk.get("reasons")}]
        scaling["per_rank_clocks_during_queries"] = allclk
        extra["scaling"] = scaling
        if cpu_q:
            extra["cpu_baseline_queries"] = {"kind": "port (oracle/liboracle.so, same semantics as the GPU column)", "map": "C2", **cpu_q}

        # ---- BASELINE config C1 (single Theta* query (0,0,0)->(40,40,4), 250x250x50 @0.2 m) and C3 (10 000 costlazythetastar
        # queries, 256x256x64, LoS-heavy): rank 0 only, N = 1 only
        if rank == 0 and world == 1:
            try:
                extra["c1_single_query"] = run_c1(b200, po if not args.no_cpu else None, synth, local_rank, errors)
                extra["c3_costlazythetastar"] = run_c3(b200, po if not args.no_cpu else None, synth, torch, dev, local_rank, host_threads, errors)
            except Exception as e:  # never let an extra take the headline down
                extra["c1_c3_error"] = repr(e)[:300]

    # ---- extra (N >= 2): BASELINE config 5 — 2048x2048x512 exact EDT, z-slab sharded
    if world >= 2 and not args.no_extra and 512 % world == 0:
        gm_keep = gm
        try:
            extra["edt_slab_c5"], errs = run_slab_edt(b200, po, torch, dist, dev, rank, world, local_rank, check_threads)
            errors.extend(errs)
        except Exception as e:
            extra["edt_slab_c5"] = {"error": repr(e)[:300]}
            errors.append(f"rank {rank}: C5 slab EDT raised {repr(e)[:200]}")
        gm = gm_keep

    # ---- extra (rank 0, N = 1): the depth-image fusion path — 640x480 frames with the reference's default
    # intrinsics into a second C2-sized map, host image in, fused grids resident; the sequential CPU restatement beside it
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            rng_ = (5.5, 5.5, 4.5)
            gd = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=rng_, ground_height=c["origin"][2],
                              device=local_rank)
            gd.set_depth_params()
            od = None if args.no_cpu else po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=rng_,
                                                       ground_height=c["origin"][2])
            if od is not None:
                od.set_depth_params()
            dpp = b200.DEPTH_DEFAULTS
            boxes = synth.pillar_boxes(c["seed"], c["origin"], c["size"], c["n_pillars"])
            tg, tc, stt = [], [], {}
            for f in range(7):
                pos = np.array([-3.0 + 0.12 * f, 2.0 - 0.05 * f, 1.2])
                R, _ = synth.look_at(0.7 + 0.03 * f, 0.15)
                img = synth.render_depth(boxes, c["origin"][2], pos, R, 480, 640, dpp["fx"], dpp["fy"], dpp["cx"], dpp["cy"],
                                         max_depth=8.0, dropout=0.02, seed=f)
                gd.sync()
                t0 = time.perf_counter()
                gd.depthPoseCallback(img, pos, R)
                gd.sync()
                tg.append((time.perf_counter() - t0) * 1e3)
                if od is not None:
                    t0 = time.perf_counter()
                    od.update_depth(img, pos, R)
                    tc.append((time.perf_counter() - t0) * 1e3)
                stt = gd.depth_stats()
            if od is not None and not (np.array_equal(gd.occupancy(), od.occupancy()) and np.array_equal(gd.inflate(), od.inflate())):
                errors.append("depth fusion: grids differ from the oracle after 7 frames")
            extra["depth_fusion"] = {"frame": "640x480 CV_16UC1, reference default intrinsics and filter (skip_pixel 2)",
                                     "gpu_ms_per_frame_e2e": float(np.median(tg[2:])), "frames": len(tg),
                                     "cpu_sequential_ms_per_frame": float(np.median(tc[2:])) if tc else None,
                                     "last_frame": stt, "h2d_bytes_per_frame": 480 * 640 * 2}
            gd.close()
        except Exception as e:  # never let the extra take the headline down
            extra["depth_fusion"] = {"error": repr(e)[:200]}

    # ---- e2e, measured LAST (after the query extras, a minute after the pinned buffer was allocated: its DMA rate needs both
    # transfers and time to settle on these virtualised boxes, profiles/diag_h2d.py).
    # Every NVML query takes the driver's lock for a moment; polled every 2 ms it delays the ~25 API calls of a host-facing frame
    # by 9 % (measured: 0.760 ms with the 2 ms poller, 0.694 ms without), so this region has its own sampler at 50 ms.
    gm.build_edt(D_SAFE)
    esampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    esampler.interval = 0.05
    esampler.start()
    e2e_ms, e2e_p50 = time_e2e(h_cloud_np)
    # the link this e2e number was measured over: plain pinned 16 MB host-to-device copies of the same buffer (CUDA events).
    # The frame cannot start its EDT before the last point has arrived, so e2e ~ 16 MB / link + EDT (+ ~0.05 ms); the boxes of
    # this pool gave 23-54 GB/s for this copy on different days
    ha, hb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d_ms = []
    for _ in range(10):
        ha.record()
        d_cloud.copy_(h_cloud, non_blocking=True)
        hb.record()
        torch.cuda.synchronize()
        h2d_ms.append(ha.elapsed_time(hb))
    h2d_gbs = blob.nbytes / (float(np.median(h2d_ms)) * 1e-3) / 1e9
    e2e_pageable_ms, e2e_pageable_p50 = time_e2e(pageable_np)
    e2e_clocks = esampler.stop()
    extra["h2d_link"] = {"pinned_16MB_copy_gbs": h2d_gbs, "copy_ms": float(np.median(h2d_ms)),
                         "note": "plain cudaMemcpyAsync of the e2e cloud buffer, measured right after the e2e region"}
    extra["e2e_pageable_host_cloud"] = {"ms_per_step": e2e_pageable_ms, "p50_ms_per_step": e2e_pageable_p50,
                                        "note": "same call with a pageable (non-pinned) host cloud, as a ROS PointCloud2::data is"}

    # ---- cpu_baseline: the oracle on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import pyoracle as po
        nthreads = po.max_threads()
        om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=(60.0, 60.0, 20.0),
                          ground_height=c["origin"][2])
        om.update_cloud(blob, cam, point_step=16, nthreads=nthreads)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 20):
            om.update_cloud(blob, cam, point_step=16, nthreads=nthreads)
            om.build_edt(d_safe=D_SAFE, nthreads=nthreads)
            om.local_bounds()
            reps += 1
        cpu_ms = (time.perf_counter() - t0) * 1e3 / reps
        t1 = time.perf_counter()
        om.update_cloud(blob, cam, point_step=16, nthreads=1)
        om.build_edt(d_safe=D_SAFE, nthreads=1)
        cpu_ms_1t = (time.perf_counter() - t1) * 1e3
        cpu = {"value": cpu_ms, "unit": "ms/frame", "cores": nthreads, "kind": "port",
               "sample": f"{reps} full C2 frames (cloud update + EDT + cost) with {nthreads} threads; single-thread frame: {cpu_ms_1t:.0f} ms",
               "single_thread_ms": cpu_ms_1t}

    # parity failures anywhere -> every rank learns about it, the line carries them, the exit code is non-zero
    n_err = len(errors)
    if world > 1:
        t = torch.tensor([float(n_err)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n_err = int(t.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, errors)
        errors = [e for part in gathered for e in part]
    if rank == 0:
        kernels_per_step = (["scatter_kernel", "edt_fused_y_kernel (z half 0)", "edt_fused_y_kernel (z half 1)", "edt_fused_x_kernel (z half 0)",
                             "edt_fused_x_kernel (z half 1)"] if fused else
                            ["scatter_kernel", "edt_z_kernel", "edt_scan_y_kernel", "edt_cross_kernel", "edt_fill_y_kernel",
                             "edt_scan_x_kernel", "edt_cross_kernel", "edt_fill_x_kernel"])
        line = {
            "metric": "map rebuild ms/frame", "value": ms_per_step, "unit": "ms/frame", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "config": {"workload": "C2: 1M-point PointCloud2 -> 512x512x128 inflated occupancy + exact squared EDT + cost field",
                       "grid": list(gm.dims), "points": int(n_pts), "d_safe_m": D_SAFE, "l2": "flushed (256 MiB write) before every timed frame",
                       "multi_gpu": "replicas only: every rank rebuilds the same map replica (value = that frame time, NOT divided by n_gpus); "
                                    "query batches are sharded with no collective (extra.scaling)"},
            "e2e": {"value": e2e_ms, "unit": "ms/frame", "h2d_bytes_per_step": int(blob.nbytes), "d2h_bytes_per_step": 48,
                    "ms_per_step": e2e_ms, "p50_ms_per_step": e2e_p50, "host_buffer": "pinned", "clocks": e2e_clocks,
                    "d2h_note": "the fields stay resident for the planners; only the 6 local-bound keys are read back"},
            "gpu_launches": int(gpu_launches), "kernels_per_step": kernels_per_step + ["(+1 cudaMemsetAsync clear; cost_lut_kernel only when d_safe changes)"],
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
            "parity_checks": {"failed": n_err, "errors": errors[:20]},
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()
    if n_err:
        sys.exit(3)


if __name__ == "__main__":
    main()
