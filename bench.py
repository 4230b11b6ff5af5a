#!/usr/bin/env python
"""bench.py — particle-updates/s of the Verlet cloth step (BASELINE.json's metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                          the reference's own CPU StepPhysics
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU (N = 2, 4, 8)

A "step" is one StepPhysics over the whole cloth (one substep).  Workloads (BASELINE.json configs):
  N = 1 : 2048 x 2048 cloth, reference parameters, flat-sheet start                       (config 3)
  N > 1 : 8192 x 8192 cloth cut into N LINKED row bands (strong scaling): the step kernel itself stores each band's
          two boundary rows into the neighbour's halo over NVLink peer memory and orders the GPUs with flag words,
          tile by tile; torch.distributed (NCCL) only carries the endpoints, barriers and the timing reductions (config 4)
  --workload batch : 4096 independent 128 x 128 cloths sharded over the ranks             (config 5, no communication)

Prints ONE JSON line (rank 0).  Timing: CUDA events on the stream the kernels run on, W (>= 3) warm-up steps, then
exactly K steps between barrier + synchronize, max over ranks.  Everything that is set-up (process group, IPC
mapping of the neighbours' buffers, the first launches) happens before the warm-up ends; the timed region contains
kernel launches only.  The state (>= 200 MB) is larger than the 126 MB L2, so consecutive steps cannot be served
from cache.  Every N > 1 line also carries `scaling_base` (the same 8192^2 cloth on rank 0's GPU alone, same run),
`parallel_efficiency` against it, and `parity`: the bands' state after all the steps compared bit for bit with the
whole cloth stepped on one GPU (exact mode: by the independent gather kernel; fast mode: by the same kernel, whose
result does not depend on the decomposition).

Arithmetic mode (--mode): `fast` (default) is north_star's tolerance mode — FMA contraction, MUFU.RSQ, checked against
the reference CPU path to <= 1e-5 of the cloth extent after 100 steps and 1e-3 after 1000 (tests/test_parity_gpu.py,
incl. at this bench's 2048^2) — and runs the streaming gather kernel oc_k_stream; `exact` is bit-identical to the
reference CPU path (oc_k_march2).  Every N = 1 line reports the other mode as `other_mode`.
Every N = 1 line also carries `scaling_base` (8192^2 on this GPU), `batch_config5` (512 x 128^2 cloths) and `mid_size`
(256^2 = BASELINE config 2 and 512^2: 1000 steps in one launch of the band-resident kernel oc_k_bandres).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_updates_per_s"
UNIT = "particle-updates/s"
ALG_BYTES_PER_UPDATE = 48          # read X, X_last + write X, X_last, 3 x fp32 each


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled while the timed region runs
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU, sampled in a thread while the timed region runs: NVML
    directly (nvidia-ml-py, every 2 ms — the timed region of a 20-step run lasts under 2 ms), or the recipe's
    `nvidia-smi --query-gpu=... -lms` line when NVML cannot be loaded.  Samples inside [t0, t1] count; if the region was
    shorter than the sampling period, the nearest sample either side of it."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []            # (time, sm MHz, set of reason names)
        self.smax = None
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may renumber them for this process
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    idx = int(ids[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))

            def poll():
                while not self.stop_flag:
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.rows.append((time.time(), mhz, {n for n, b in names if bits & b}))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.nvml = pynvml
            self.source = "nvml, 2 ms"
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi -lms 100"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_wait = time.time()
            while not self.rows and time.time() - t_wait < 5.0:       # nvidia-smi needs a moment before its first line
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                mhz = float(f[1]); self.smax = float(f[2])
            except ValueError:
                continue
            reasons = {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8])
                       if v.lower().startswith("active")}
            self.rows.append((time.time(), mhz, reasons))

    def stop(self, t0, t1):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        time.sleep(0.15 if self.proc is not None else 0.01)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = list(self.rows)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        if not inside:                                                  # region shorter than the sampling period
            before = [r for r in rows if r[0] < t0][-1:]
            after = [r for r in rows if r[0] > t1][:1]
            inside = before + after
        sm = sorted(r[1] for r in inside)
        reasons = set()
        for r in inside:
            reasons |= r[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons), "samples": len(sm),
                "source": self.source}


# ---------------------------------------------------------------------------------------------
# workload description shared by both arms (the driver compares the two `config` objects)
# ---------------------------------------------------------------------------------------------
def workload_name(args):
    if args.workload == "batch":
        return "4096 independent 128x128 cloths sharded over the ranks (BASELINE config 5)"
    if args.gpus == 1:
        return f"{args.n}x{args.n} cloth, single B200 (BASELINE config 3)"
    return f"{args.n}x{args.n} cloth, {args.gpus} row bands (BASELINE config 4)"


def config_of(args):
    return {"workload": workload_name(args), "cloth": [args.n, args.n] if args.workload == "cloth" else [128, 128],
            "cloths": 1 if args.workload == "cloth" else 4096, "bands": args.gpus if args.workload == "cloth" else 1,
            "mode": args.mode, "parameters": "reference defaults (V:59-62, V:97-104, V:123-130), flat-sheet start (V:254-260)"}


# ---------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU implementation (oracle/_ref, verbatim StepPhysics)
# ---------------------------------------------------------------------------------------------
def cpu_time(side, steps, warmup, threads=1, verbatim=True):
    """Times `steps` StepPhysics of a side x side cloth on the host.  verbatim: the reference's own text
    (oracle/_ref, single thread — that is how the reference runs it); else the C restatement (OpenMP)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    helpers.ensure_built()
    use_ref = verbatim and helpers.have_ref()
    if use_ref:
        sim = helpers.Ref(side, side)
    else:
        os.environ["OMP_NUM_THREADS"] = str(threads)
        sim = helpers.Oracle(side, side)
    if warmup:
        sim.step(warmup)
    t0 = time.perf_counter()
    sim.step(steps)
    dt = time.perf_counter() - t0
    if hasattr(sim, "close"):
        sim.close()
    return side * side * steps / dt, dt, ("reference" if use_ref else "port")


def cpu_reference(n_side, steps, warmup, budget_s=20.0):
    """The verbatim reference StepPhysics on a bounded sample: an n x n cloth sized so that (steps + warmup) steps
    fit the budget (the full side if it fits)."""
    rate_guess = 5.0e6                                   # updates/s at large grids
    per_step = budget_s / max(1, steps + warmup)
    side = int(min(n_side, max(21, (per_step * rate_guess) ** 0.5)))
    val, dt, kind = cpu_time(side, steps, warmup)
    return {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "cpu": cpu_model(), "host_logical_cores": os.cpu_count(),
            "sample": f"{steps} steps of a {side}x{side} cloth (reference parameters), single thread as the reference runs it: "
                      f"1 core used of {os.cpu_count()}"}, dt / steps * 1e3


def cpu_baseline_configs():
    """the reference's CPU path at 21^2 x 1000, 256^2 x 1000 and 2048^2 x 10 steps, one core."""
    out = []
    for side, steps in ((21, 1000), (256, 1000), (2048, 10)):
        val, dt, kind = cpu_time(side, steps, 0)
        out.append({"cloth": f"{side}x{side}", "steps": steps, "value": val, "unit": UNIT, "seconds": dt, "kind": kind, "cores": 1})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    base, ms = cpu_reference(args.n, steps, max(0, args.warmup), budget_s=90.0)
    # for context only: the C restatement of the same algorithm (oracle/oc_oracle.c, OpenMP over particles) on every
    # host thread, same sample size — what a multi-threaded CPU port of the reference would reach on this host
    port = None
    try:
        import re
        side = int(re.search(r"of a (\d+)x", base["sample"]).group(1))
        val, dt, _ = cpu_time(side, steps, 1, threads=os.cpu_count() or 1, verbatim=False)
        port = {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": f"{steps} steps of a {side}x{side} cloth, OpenMP restatement on all host threads"}
    except Exception as e:                                   # never let the extra figure break the arm
        port = {"error": str(e)[:200]}
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if (args.gpus > 1 and args.workload == "cloth") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (closed-form flat sheet of InitGL, reference parameters)",
            "config": config_of(args),
            "detail": {"note": "bounded sample of the workload on the host CPU, see cpu_baseline.sample"},
            "cpu_baseline": base, "cpu_port_all_threads": port,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def sha_rows(a):
    return hashlib.sha256(memoryview(a)).hexdigest()


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import opencloth_b200 as oc
    from opencloth_b200 import bands as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    exact = 1 if args.mode == "exact" else 0
    K, W = args.steps, max(args.warmup, 3)
    nx = ny = args.n
    drv = None
    band = None
    linked = False
    if args.workload == "batch":
        nx = ny = 128
        total_batch = 4096
        b0, b1 = (total_batch * rank) // world, (total_batch * (rank + 1)) // world
        cloth = oc.Cloth(nx, ny, batch=b1 - b0, device=local, exact=exact, substeps_per_launch=args.k)
        particles_total = total_batch * nx * ny
    elif world == 1:
        cloth = oc.Cloth(nx, ny, device=local, exact=exact, substeps_per_launch=args.k)
        particles_total = nx * ny
    else:
        linked = args.exchange == "linked"
        # fast mode: every handle of the run (bands, the whole cloth of the parity check) uses oc_k_stream, whose result does
        # not depend on the decomposition; AUTO would pick it anyway at 8192^2 / N <= 8
        band_kernel = oc.OC_KERNEL_AUTO if exact else oc.OC_KERNEL_STREAM
        band = B.CudaBand(nx, ny, world, rank, 2 if linked else args.halo_rows, local, exact=exact,
                          substeps_per_launch=1 if linked else args.k, kernel=band_kernel)
        cloth = band.cloth
        particles_total = nx * ny
    # a non-default torch stream: the library launches on it and torch events time it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    if band is not None and not linked:
        drv = B.BandDriver(band, rank, world, overlap=args.overlap)     # binds its own stream (NCCL orders against it)
        stream = drv.stream
        torch.cuda.set_stream(stream)
    else:
        cloth.set_stream(stream.cuda_stream)
        if band is not None:
            drv = B.LinkedBandDriver(band, rank, world)
            drv.link()                                                   # IPC mapping of the neighbours' buffers, halos, barriers

    def advance(n, finish=False):
        if drv is not None:
            drv.step(n)
            if finish:
                drv.finish()
        else:
            cloth.step(n)

    steps_taken = 0
    if drv is not None and not linked:
        W = max(W, args.halo_rows + 1)          # NCCL exchange: two complete exchange periods before the clock starts
    advance(W, finish=True)
    steps_taken += W
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = cloth.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
        # The ranks leave the host barrier up to a millisecond apart (measured), and a band that starts early only
        # waits for its neighbours: each rank's own clock would then include the others' lateness.  A device-side
        # barrier (a one-word all-reduce the timed stream waits for, no host synchronisation after it) starts all
        # GPUs together, with their launches already queued behind it; ev0 comes after it.
        gate = torch.zeros(1, device=dev)
        dist.all_reduce(gate)
    t_host0 = time.time()
    ev0.record(stream)
    advance(K, finish=True)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    t_host1 = time.time()
    if world > 1:
        dist.barrier()
    steps_taken += K
    ms = ev0.elapsed_time(ev1)
    launches = cloth.launch_count - launches0
    ms_ranks = [ms]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        ms_ranks = [float(x.item()) for x in g]
        ms = max(ms_ranks)
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    value = particles_total * K / (ms * 1e-3)

    # ---- N > 1: bitwise parity of the bands with the whole cloth on one GPU, and the 1-GPU rate of the same cloth ----
    parity = None
    scaling_base = None
    if band is not None:
        if not args.no_parity:
            hx = np.empty((cloth.n_local, 3), np.float32)
            hl = np.empty((cloth.n_local, 3), np.float32)
            cloth.download(out=(hx, hl))
            mine = (band.begin, band.end, sha_rows(hx), sha_rows(hl))
            every = [None] * world
            dist.all_gather_object(every, mine)
            if rank == 0:
                # the independent cross-check kernel (one thread per particle, 12-neighbour gather, scalar exact math)
                whole = oc.Cloth(nx, ny, device=local, exact=exact, kernel=oc.OC_KERNEL_GATHER if exact else oc.OC_KERNEL_STREAM)
                whole.step(steps_taken)
                wx, wl = whole.download()
                whole.close()
                bad = []
                for (b, e, sx, sl_) in every:
                    if sha_rows(wx[b * nx:e * nx]) != sx or sha_rows(wl[b * nx:e * nx]) != sl_:
                        bad.append([b, e])
                parity = {"bitwise": len(bad) == 0, "rows": ny, "bands": [[b, e] for (b, e, _, _) in every], "steps": steps_taken,
                          "against": ("whole cloth on one GPU, oc_k_gather (independent kernel), SHA-256 of X and X_last per band" if exact
                                      else "whole cloth on one GPU, same kernel and mode (fast mode is deterministic), SHA-256 of X and X_last per band"),
                          "mismatching_bands": bad}
                del wx, wl
            dist.barrier()
        if rank == 0:
            big = oc.Cloth(nx, ny, device=local, exact=exact, substeps_per_launch=args.k)
            big.step(W)
            n_b = max(3, min(max(K, 20), 200))
            b_ms = big.step_timed(n_b)
            big.close()
            scaling_base = {"workload": f"{nx}x{ny} cloth, single B200 (this workload on rank 0's GPU alone, same run)",
                            "value": nx * ny * n_b / (b_ms * 1e-3), "unit": UNIT, "steps": n_b, "mode": args.mode}
        dist.barrier()

    # ---- e2e: the same metric through the C-ABI with HOST buffers (pinned): upload, step, download ----
    n_local = cloth.n_local
    hx = torch.empty((n_local, 3), dtype=torch.float32).pin_memory()
    hl = torch.empty((n_local, 3), dtype=torch.float32).pin_memory()
    cloth.download_into(hx.data_ptr(), hl.data_ptr(), 3)
    e_steps = max(3, min(K, args.e2e_steps))

    def e2e_step():
        cloth.upload_from(hx.data_ptr(), hl.data_ptr(), 3)      # H2D of X, X_last (pinned, 12 B/particle each)
        if linked:
            drv.resync()                                       # linked row bands: the halo rows follow the uploaded state
        advance(1, True)                                       # one StepPhysics
        cloth.download_into(hx.data_ptr(), hl.data_ptr(), 3)    # D2H of X, X_last; synchronises

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": particles_total * e_steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": particles_total * 24, "d2h_bytes_per_step": particles_total * 24,
           "steps": e_steps, "note": "per step: oc_upload(X, X_last from pinned host) + oc_step(1) + oc_download(X, X_last)"
                                     + (" (+ halo refresh between the bands)" if band is not None else "")}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = value / max(1, world) * ALG_BYTES_PER_UPDATE / 1e9          # per GPU
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.mode}_k{args.k}_{nx}")
        except Exception:
            pass
        if band is None:
            exch = "none"
        elif linked:
            exch = ("in-kernel: every step the band-edge tiles store 2 rows per side into the neighbour's halo over NVLink peer memory "
                    "(CUDA IPC) and release per-strip flag words the neighbour's edge tiles wait for; no host exchange")
        else:
            exch = "NCCL send/recv, " + ("overlapped" if args.overlap else "blocking") + f", one per {args.halo_rows // 2} steps"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak" if (world == 1 or args.workload == "batch") else "strong",
                "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (closed-form flat sheet of InitGL V:254-260, reference parameters; state larger than L2)",
                "config": config_of(args),
                "detail": {"mode": args.mode + (" (bit-identical to the reference CPU path)" if exact else " (FMA/rsqrt, within 1e-5 / 1e-3 of extent)"),
                           "substeps_per_launch": args.k,
                           "kernel": ("oc_k_march (staged, k substeps per launch)" if args.k > 1 else
                                      ("oc_k_march2 (fused marching stencil, every spring once, 2 columns/thread, packed FP32x2)" if exact else
                                       "oc_k_stream (fused streaming gather over twin tiles, 12 springs per particle, packed FP32x2)")),
                           "l2": "state 48 B x particles per step > 126 MB L2 (no flush needed)" if particles_total * 48 / max(1, world) > 126e6 else "state fits L2",
                           "halo_rows": (2 if linked else args.halo_rows) if band is not None else 0, "exchange": exch,
                           "halo_bytes_per_step_per_gpu": (2 * 2 * nx * 16 if (band is not None and linked) else None),
                           "ms_per_rank": [m_ / K for m_ in ms_ranks]},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_update": ALG_BYTES_PER_UPDATE,
                             "note": "per GPU; achieved = updates/s/GPU x 48 B; the kernel is FP32-issue bound, not DRAM bound (DESIGN.md)"},
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e}
        if band is not None:
            line["exchanges"] = drv.exchanges if not linked else 0
            line["exchange_ms_per_step"] = 0.0 if linked else None
            line["scaling_base"] = scaling_base
            line["parallel_efficiency"] = value / (world * scaling_base["value"]) if scaling_base else None
            line["parity"] = parity
        if world == 1 and args.workload == "cloth":
            # the library's other arithmetic mode on the same workload (not the headline; CUDA events in oc_step_timed)
            other = oc.Cloth(nx, ny, device=local, exact=0 if exact else 1, substeps_per_launch=args.k)
            other.step(W)
            n_o = max(3, min(max(K, 20), 500))
            o_ms = other.step_timed(n_o)
            o_val = nx * ny * n_o / (o_ms * 1e-3)
            other.close()
            line["other_mode"] = {"mode": "fast" if exact else "exact", "value": o_val, "unit": UNIT, "steps": n_o,
                                  "roofline_frac": o_val * ALG_BYTES_PER_UPDATE / 1e9 / peak}
        if world == 1 and K < 500 and args.workload == "cloth":
            # a short timed region carries the fill of the launch chain's first step and the drain of its last one (about one
            # tile lifetime, ~60-90 us, 7 % of 20 steps at 2048^2): the same workload, same mode, over 2000 steps for comparison
            ss = oc.Cloth(nx, ny, device=local, exact=exact, substeps_per_launch=args.k)
            ss.step(W)
            n_s = 2000
            s_ms = ss.step_timed(n_s)
            ss.close()
            s_val = particles_total * n_s / (s_ms * 1e-3)
            line["steady_state"] = {"value": s_val, "unit": UNIT, "steps": n_s, "ms_per_step": s_ms / n_s, "mode": args.mode,
                                    "roofline_frac": s_val * ALG_BYTES_PER_UPDATE / 1e9 / peak,
                                    "note": "same workload and mode over 2000 steps (CUDA events inside oc_step_timed); `value` above is the K-step figure"}
        if world == 1 and args.workload == "cloth" and nx != 8192:
            # N > 1 runs split the 8192^2 cloth of BASELINE config 4 (strong scaling): its 1-GPU rate, measured here, is
            # the base a parallel efficiency has to be computed against (not this line's 2048^2 value)
            big = oc.Cloth(8192, 8192, device=local, exact=exact, substeps_per_launch=args.k)
            big.step(W)
            n_b = max(3, min(max(K, 20), 200))
            b_ms = big.step_timed(n_b)
            big.close()
            line["scaling_base"] = {"workload": "8192x8192 cloth, single B200 (the N>1 workload on one GPU)", "value": 8192 * 8192 * n_b / (b_ms * 1e-3),
                                    "unit": UNIT, "steps": n_b, "mode": args.mode}
        if world == 1 and args.workload == "cloth" and not args.no_batch:
            # BASELINE config 5 on this GPU (its share of the 8-GPU job: 512 of the 4096 cloths), same mode
            bc = oc.Cloth(128, 128, batch=512, device=local, exact=exact)
            bc.step(W)
            n_c = max(3, min(max(K, 20), 200))
            c_ms = bc.step_timed(n_c)
            bc.close()
            line["batch_config5"] = {"workload": "512 independent 128x128 cloths (one GPU's share of BASELINE config 5; no communication)",
                                     "value": 512 * 128 * 128 * n_c / (c_ms * 1e-3), "unit": UNIT, "steps": n_c, "mode": args.mode}
        if world == 1 and args.workload == "cloth" and not args.no_batch:
            # BASELINE config 2 (256 x 256, the parity configuration) and 512 x 512: mid-size cloths are served by the
            # band-resident kernel (oc_k_bandres: state in shared memory, ALL the substeps of the call in one launch)
            mid = {}
            for side in (256, 512):
                mc = oc.Cloth(side, side, device=local, exact=exact)
                mc.step(W)
                n_m = 1000
                l0 = mc.launch_count
                m_ms = mc.step_timed(n_m)
                mid[f"{side}x{side}"] = {"value": side * side * n_m / (m_ms * 1e-3), "unit": UNIT, "steps": n_m, "ms_per_1000_steps": m_ms * 1000.0 / n_m,
                                         "launches": mc.launch_count - l0}
                mc.close()
            mid["mode"] = args.mode
            mid["kernel"] = "oc_k_bandres (AUTO for one whole cloth of about 10^3 .. 3*10^5 particles)"
            line["mid_size"] = mid
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_reference(nx, 3, 1, budget_s=15.0)
            line["cpu_baseline"] = base
            try:
                line["cpu_baseline_configs"] = cpu_baseline_configs()
            except Exception as e:
                line["cpu_baseline_configs"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cloth", choices=["cloth", "batch"])
    ap.add_argument("--n", type=int, default=0, help="cloth side; default 2048 at N=1, 8192 at N>1")
    ap.add_argument("--mode", default="fast", choices=["exact", "fast"],
                    help="fast (default): the tolerance mode of north_star (<= 1e-5 of the extent after 100 steps, 1e-3 after 1000); "
                         "exact: bit-identical to the reference CPU path.  The line always carries the other one as other_mode")
    ap.add_argument("--k", type=int, default=1, help="substeps per launch (temporal blocking)")
    ap.add_argument("--exchange", default="linked", choices=["linked", "nccl"],
                    help="row bands: linked = in-kernel peer stores + flag words (default); nccl = host-driven send/recv every halo_rows/2 steps")
    ap.add_argument("--halo-rows", type=int, default=24, help="--exchange nccl: halo rows either side (one exchange per halo_rows/2 steps)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--overlap", action="store_true", help="--exchange nccl: start the exchange on a side stream during the last substep of a group")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bitwise comparison with the whole cloth on one GPU")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.gpus = max(args.gpus, world) if world > 1 else args.gpus
    if args.n == 0:
        args.n = 2048 if args.gpus == 1 else 8192
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
