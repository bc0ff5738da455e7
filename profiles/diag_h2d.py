"""Host-to-device diagnostics for the e2e frame: pinned-copy bandwidth of a 16 MB cloud with and without binding the
process to the GPU's own CPU set (NVML cpu affinity), and the e2e frame time through the C-ABI in both settings.
usage: diag_h2d.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pynvml


def gpu_cpus(index=0):
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    n = (os.cpu_count() + 63) // 64
    masks = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    cpus = [64 * i + b for i, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1]
    return cpus


def h2d_gbs(buf, dev, reps=20):
    d = torch.empty_like(buf, device=dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record()
        d.copy_(buf, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return buf.numel() * buf.element_size() / (np.median(ts) * 1e-3) / 1e9, np.median(ts)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    print("cpus now:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
    try:
        cpus = gpu_cpus(0)
        print("NVML affinity of GPU 0:", cpus[:4], "...", cpus[-4:], "n =", len(cpus))
    except Exception as e:
        cpus = None
        print("NVML affinity unavailable:", repr(e))
    os.system("lscpu | grep -i -E 'numa|model name|socket' | head -8; nvidia-smi topo -m 2>/dev/null | head -8")
    import astar_planners_b200 as b200
    from astar_planners_b200 import synth
    c = synth.CONFIGS["C2"]
    blob = synth.pack_cloud(synth.config_cloud("C2", n_points=1_000_000), 16)
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=(60.0, 60.0, 20.0), ground_height=c["origin"][2], device=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def e2e(buf, reps=20):
        ts = []
        for i in range(reps + 3):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gm.cloudCallback(buf, (0.0, 0.0, 1.0), point_step=16)
            gm.build_edt(1.0)
            gm.local_bounds()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(ts[3:])), float(np.mean(ts[3:]))

    # buffer- or time-dependence?  four pinned buffers (torch's allocator, plus one cudaHostAlloc'd directly and one
    # cudaHostRegister'ed numpy array), measured round-robin five times
    bufs = {}
    for i in range(3):
        bufs[f"torch{i}"] = torch.from_numpy(blob.copy()).pin_memory()
    reg = blob.copy()
    rc = torch.cuda.cudart().cudaHostRegister(reg.ctypes.data, reg.nbytes, 0)
    print("cudaHostRegister rc", rc)
    for rnd in range(5):
        for k, h in bufs.items():
            gbs, ms = h2d_gbs(h, dev, reps=8)
            p50, mean = e2e(h.numpy(), reps=8)
            print(f"round {rnd} {k}: ptr {h.data_ptr():#x} raw {gbs:.1f} GB/s, e2e pinned p50 {p50:.3f} mean {mean:.3f}")
        p50, mean = e2e(reg, reps=8)
        print(f"round {rnd} registered numpy: e2e p50 {p50:.3f} mean {mean:.3f}")
        gm.set_profiling(True)
        e2e(bufs["torch0"].numpy(), reps=2)
        print("   stage ms:", [round(float(x), 3) for x in gm.stage_ms()[:12]])
        gm.set_profiling(False)
    gm.close()


if __name__ == "__main__":
    main()
