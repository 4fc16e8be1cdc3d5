#!/usr/bin/env python3
"""bench.py -- headline benchmark of the d2q9-bgk timestep loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--timesteps T] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): MLUPS = lattice-cell updates per second / 1e6, and the achieved
HBM bandwidth it implies (72 B per update) against the measured B200 peak.

Workload: the synthetic channel of BASELINE.json configs[3]/[4] -- nx = 16384, ny = 16384
rows per GPU (weak scaling), walls on rows 0 and ny-1, Bernoulli(1 %) obstacles, rest
state at t = 0.  One bench "step" is one pass of the step loop: T consecutive timesteps
(`--timesteps`, default 200, the lower end of SURVEY.md section 8d's 200-1000) issued by
one lbm_gpu_run() call.  The lattice (19.3 GB per GPU) is far larger than L2 (126 MB),
so nothing stays cached between timesteps or steps.

  value   lattice resident in HBM, K steps timed with CUDA events on the library's
          stream (max over ranks): all cells x T x K / time.
  e2e     the same K steps through the C-ABI from HOST buffers, everything inside the
          timed region every step: lbm_gpu_create (pinned int32 obstacle array -> HBM,
          like the reference's int obstacles[]), lbm_gpu_run (T timesteps, av_vels to
          the host), lbm_gpu_final_fields (u_x,u_y,|u|,pressure of every cell -> pinned
          host arrays, what write_values needs), lbm_gpu_destroy.
  --impl reference   the reference's own timestep_new2, compiled from /root/reference
          into oracle/_ref (or the oracle port if that is absent), on the host cores
          with all threads, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tools.make_inputs import channel_mask  # noqa: E402

# LBM_BENCH_NX / LBM_BENCH_ROWS shrink the grid for the harness's own tests; the benchmark
# contract is the default 16384 x 16384 per GPU.
NX = int(os.environ.get("LBM_BENCH_NX", "16384"))
ROWS_PER_GPU = int(os.environ.get("LBM_BENCH_ROWS", "16384"))
DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85
BYTES_PER_UPDATE = 72.0            # 9 fp32 loads + 9 fp32 stores
CPU_SAMPLE_ROWS = min(1024, ROWS_PER_GPU)   # CPU legs run a 16384 x 1024 slab of the same generator
HBM_FALLBACK_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


# ------------------------------------------------------------------------- clocks ----
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # samples under load: power above 40 % of the observed maximum
        if sm:
            pmax = max(power)
            load = [s for s, p in zip(sm, power) if p >= 0.4 * pmax] or sm
            return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": pmax,
                    "samples": len(sm), "reasons": sorted(reasons)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except (OSError, KeyError, ValueError):
        return HBM_FALLBACK_GBS, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def ncu_traffic_per_launch():
    """dram bytes per step-kernel launch from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "ncu_step_kernel_summary.json")
    try:
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


# ----------------------------------------------------------------- CPU reference ----
class RefParam(ctypes.Structure):      # t_param of the reference, d2q9-bgk.c:64-73
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("maxIters", ctypes.c_int),
                ("reynolds_dim", ctypes.c_int), ("density", ctypes.c_float), ("accel", ctypes.c_float),
                ("omega", ctypes.c_float)]


def cpu_reference_mlups(budget_s, threads=None, prefer="omp"):
    """Time the reference's own timestep_new2 (d2q9-bgk.c:228) on a 16384 x 1024 slab of the
    bench workload for about budget_s seconds.  Uses oracle/_ref (the reference compiled
    from /root/reference by oracle/build_oracle.py); falls back to the oracle port.
    This is the one place outside tests/ that executes anything under oracle/: it is the
    measured CPU baseline, never part of the GPU path."""
    nx, ny = NX, CPU_SAMPLE_ROWS
    ncores = os.cpu_count() or 1
    threads = threads or ncores
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    kind = "reference"
    lib = None
    name = "libref_f32_omp.so" if (prefer == "omp" and threads > 1) else "libref_f32_fast.so"
    if os.path.exists(os.path.join(ref_dir, name)):
        lib = ctypes.CDLL(os.path.join(ref_dir, name))
        fn = lib.timestep_new2
        fn.argtypes = [RefParam, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        fn.restype = ctypes.c_float
        prm = RefParam(nx, ny, 1, 10, DENSITY, ACCEL, OMEGA)
        step = lambda a, b, o: fn(prm, a, b, o)
        if name == "libref_f32_fast.so":
            threads = 1
    else:
        kind = "port"
        so = os.path.join(ROOT, "oracle", "liblbm_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_oracle.py")])
        lib = ctypes.CDLL(so)
        fn = lib.oracle_timestep_f32
        fl = ctypes.c_float
        fn.argtypes = [ctypes.c_int, ctypes.c_int, fl, fl, fl] + [ctypes.c_void_p] * 5
        fn.restype = fl
        scratch = np.empty(nx * ny, dtype=np.float32)
        step = lambda a, b, o: fn(nx, ny, DENSITY, ACCEL, OMEGA, a, b, o, scratch.ctypes.data, None)
    mask = channel_mask(nx, ny).astype(np.int32)
    w = (np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4) * DENSITY).astype(np.float32)
    a = np.tile(w, (ny * nx, 1))
    b = np.empty_like(a)
    pa, pb, po = a.ctypes.data, b.ctypes.data, mask.ctypes.data
    step(pa, pb, po)                      # warm-up (page faults of b)
    pa, pb = pb, pa
    n, t0 = 0, time.perf_counter()
    while True:
        step(pa, pb, po)
        pa, pb = pb, pa
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 10000:
            break
    return {"value": nx * ny * n / dt / 1e6, "unit": "MLUPS", "cores": int(threads), "kind": kind,
            "sample": "%dx%d slab of the same channel generator, %d timesteps of timestep_new2 in %.1f s "
                      "(%s)" % (nx, ny, n, dt, name if kind == "reference" else "oracle port"),
            "host_cores_visible": ncores}, n, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    per_step_budget = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    if os.environ.get("LBM_BENCH_CPU_BUDGET_S"):          # tests: shorter CPU sample
        per_step_budget = float(os.environ["LBM_BENCH_CPU_BUDGET_S"])
    for _ in range(args.warmup):
        cpu_reference_mlups(per_step_budget / 4)
    vals, tot_updates, tot_time = [], 0.0, 0.0
    info = None
    for _ in range(args.steps):
        info, n, dt = cpu_reference_mlups(per_step_budget)
        tot_updates += NX * CPU_SAMPLE_ROWS * n
        tot_time += dt
    value = tot_updates / tot_time / 1e6
    info["value"] = value
    line = {
        "impl": "reference", "metric": "MLUPS (d2q9-bgk lattice updates per second / 1e6)",
        "value": value, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_time / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    ny = ROWS_PER_GPU * world if args.scaling == "weak" else ROWS_PER_GPU
    return {"workload": "synthetic %dx%d channel (walls rows 0 and ny-1, Bernoulli 1%% obstacles, seed 20240229), "
                        "%d rows per GPU" % (NX, ny, ny // world),
            "nx": NX, "ny": ny, "timesteps_per_step": args.timesteps,
            "density": DENSITY, "accel": ACCEL, "omega": OMEGA,
            "l2": "inputs larger than L2 (%.1f GB lattice per GPU vs 126 MB), no flush needed"
                  % (NX * (ny // world) * 72 / 1e9),
            "parallelism": "row slabs, %d GPU(s), halo rows pushed by the step kernel over NVLink" % world}


# ------------------------------------------------------------------------ GPU arm ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--timesteps", type=int, default=200, help="lattice timesteps per bench step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 16384 rows per GPU (default, the contract); strong: 16384 rows in total "
                         "(BASELINE.json configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                      # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d"
                     % (args.gpus, args.gpus))
        args.gpus = world

    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import lbm_b200 as L
    slabs = __import__("importlib").import_module("advanced-hpc-lbm_b200.slabs")

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ny = ROWS_PER_GPU * world if args.scaling == "weak" else ROWS_PER_GPU
    row0, nrows = L.split_rows(ny, world)[rank]
    T, K, W = args.timesteps, args.steps, args.warmup
    cells_total = float(NX) * float(ny)

    # this rank's rows of the obstacle mask, as the reference holds it: one int per cell,
    # in page-locked host memory (ordinary memory if the box refuses to pin that much)
    host_memory = "pinned"

    class _Pageable:
        def __init__(self, shape, dtype):
            self.array = np.empty(shape, dtype=dtype)

        def free(self):
            self.array = None

    def host_array(shape, dtype):
        nonlocal host_memory
        try:
            return L.PinnedArray(shape, dtype)
        except L.LbmError:
            host_memory = "pageable"
            return _Pageable(shape, dtype)

    pin_obst = host_array((nrows, NX), np.int32)
    pin_obst.array[...] = channel_mask(NX, ny, rows=(row0, row0 + nrows))
    global_free = sum_over_ranks(float((pin_obst.array == 0).sum()))

    def make_lattice():
        # LBM_GPU_POOL: device memory of a destroyed lattice is reused by the next create
        if world == 1:
            return L.Lattice(NX, ny, DENSITY, ACCEL, OMEGA, obstacles=pin_obst.array, flags=L.POOL)
        lat = L.Lattice(NX, ny, DENSITY, ACCEL, OMEGA, obstacles=pin_obst.array, slab=(row0, nrows),
                        device_ids=[local_rank], flags=L.POOL)
        below, above = slabs.exchange_descriptors(lat.ipc_export(), rank, world, dist)
        lat.ipc_connect(below, above)
        barrier()
        lat.ipc_prepare()
        barrier()
        lat.set_global_free_cells(int(global_free))
        return lat

    # ---- value: lattice resident, device-timed ------------------------------------
    lat = make_lattice()
    for _ in range(W):
        lat.run_timed(T)
    launches0 = lat.info().kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(K):
        dev_ms += lat.run_timed(T)          # returns with the device idle (stream sync inside)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    wall_s = max_over_ranks(wall_s)
    launches = lat.info().kernel_launches - launches0
    av_tail = lat.run(2)                    # sanity: the numbers are finite and positive
    assert np.all(np.isfinite(av_tail)) and np.all(av_tail > 0), av_tail
    info = lat.info()
    lat.close()

    value = cells_total * T * K / (dev_ms * 1e-3) / 1e6
    kernel_ms = dev_ms / (T * K)                                  # one launch per timestep per GPU
    cells_per_launch = float(NX) * nrows
    achieved = cells_per_launch * BYTES_PER_UPDATE / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- e2e: host buffers -> C-ABI -> host buffers, every step -----------------------
    e2e = None
    if not args.no_e2e:
        out = [host_array((nrows, NX), np.float32) for _ in range(4)]
        av_host = np.empty(T, dtype=np.float32)

        verbose = bool(os.environ.get("LBM_BENCH_VERBOSE"))

        def one_e2e_step():
            t = [time.perf_counter()]
            lt = make_lattice(); t.append(time.perf_counter())
            lt.run(T, out=av_host); t.append(time.perf_counter())
            lt.final_fields(out=[o.array for o in out]); t.append(time.perf_counter())
            lt.close(); t.append(time.perf_counter())
            if verbose and rank == 0:
                sys.stderr.write("e2e step: create %.1f run %.1f fields %.1f destroy %.1f ms\n"
                                 % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])))

        one_e2e_step()                       # warm-up (first touch of the pinned pages etc.)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            one_e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        assert np.all(np.isfinite(av_host)) and np.isfinite(out[3].array[::997, ::997]).all()
        e2e = {"value": cells_total * T * K / e2e_s / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(pin_obst.array.nbytes) * world,
               "d2h_bytes_per_step": int(4 * out[0].array.nbytes + av_host.nbytes) * world,
               "ms_per_step": 1e3 * e2e_s / K, "host_memory": host_memory,
               "what": "lbm_gpu_create(LBM_GPU_POOL, int32 obstacles from pinned host) + lbm_gpu_run(T) -> av_vels on host + "
                       "lbm_gpu_final_fields(u_x,u_y,|u|,pressure) -> pinned host + lbm_gpu_destroy"}
        for o in out:
            o.free()
    pin_obst.free()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_reference_mlups(12.0)
        serial, _, _ = cpu_reference_mlups(8.0, threads=1, prefer="serial")
        cpu["serial_value"] = serial["value"]
        cpu["serial_sample"] = serial["sample"]

    if dist is not None:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    line = {
        "metric": "MLUPS (d2q9-bgk lattice updates per second / 1e6)",
        "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "achieved_hbm_gbs_per_gpu": achieved,
        "wall_ms_per_step": 1e3 * wall_s / K,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic_per_launch(),
                     "peak_source": peak_src,
                     "frac_of_8tbs_datasheet": achieved / 8000.0,   # frac > 1 means: faster than a plain copy
                     "algorithmic_bytes_per_launch": cells_per_launch * BYTES_PER_UPDATE,
                     "kernel": "lbm_step_vec4<float>", "kernel_ms": kernel_ms},
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "kernel_variant": int(info.kernel),
    }
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
