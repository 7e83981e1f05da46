#!/usr/bin/env python
"""bench.py -- action+gradient evals/s of the annealing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): Lorenz96 D=100, N=5001 model points
(Simpson-Hermite needs an odd N), L=40 observed components, a batch of 64
independent paths per GPU, synthetic twin-experiment data.  One *step* = one fused
action+gradient evaluation of the whole batch = 64 evals.  Under torchrun every rank owns its
own batch of 64 paths (weak scaling, no collective on the data path); time = max over ranks.

JSON keys: see the contract in the task statement.  ``value`` = device-resident evals/s,
``e2e`` = the same through ``va_ode.Annealer.A_gradA`` with pinned host buffers (H2D of XP and
D2H of A and grad inside the timed region), ``roofline`` = algorithmic bytes / kernel time
against MEASURED_PEAKS.json, ``cpu_baseline`` = the NumPy oracle port on one host core.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, N_MODEL, DT, K_FORCING = 100, 5001, 0.025, 8.17
PATHS_PER_GPU = 64
RM, RF0, ALPHA, BETA_EVAL, N_BETA = 4.0, 4e-6, 2.5, 10, 20
LIDX = [i for i in range(D) if i % 5 in (0, 2)]
METRIC = "action+gradient evals/sec (Lorenz96 D=100 N=5001 SimpsonHermite, 64 paths/GPU)"
UNIT = "evals/s"


def twin_data(seed=100):
    """SURVEY.md 8(d) C2 recipe: RK4-integrate L96, drop a transient, add N(0, 0.5^2) noise."""
    from varanneal_b200 import datagen
    _, truth, Y = datagen.lorenz96_twin(D=D, N=N_MODEL, dt=DT, k=K_FORCING, sigma=0.5, Lidx=LIDX,
                                        seed=seed, transient=1000)
    return truth, Y


def initial_paths(B, first_seed):
    X0 = np.empty((B, N_MODEL, D))
    P0 = np.empty((B, 1))
    for b in range(B):
        rng = np.random.RandomState(first_seed + b)
        X0[b] = 20.0 * rng.rand(N_MODEL, D) - 10.0
        P0[b, 0] = 4.0 * rng.rand() + 6.0
    return X0, P0


def algorithmic_bytes(B):
    """Bytes one launch must move: read X once and write the gradient once per path
    (16 N D) plus the observations, which the B paths of a batch share, once
    (8 N_data L).  DESIGN.md section 4."""
    return B * 16 * N_MODEL * D + 8 * N_MODEL * len(LIDX)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arms
def _oracle_problem(Y):
    from oracle.ode_port import OdeProblem      # the checker, used here as the CPU baseline only
    return OdeProblem("lorenz96", D, Y, LIDX, DT, "SimpsonHermite", np.array([K_FORCING]), [0], RM)


def _cpu_worker(args):
    Y, seed, reps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    prob = _oracle_problem(Y)
    X0, P0 = initial_paths(1, seed)
    XP = np.append(X0[0].ravel(), P0[0])
    rf = RF0 * ALPHA ** BETA_EVAL
    for _ in range(reps):
        prob.action_grad(XP, rf)
    return reps


def cpu_baseline_single_core(Y, budget_s=12.0):
    prob = _oracle_problem(Y)
    X0, P0 = initial_paths(1, 1000)
    XP = np.append(X0[0].ravel(), P0[0])
    rf = RF0 * ALPHA ** BETA_EVAL
    prob.action_grad(XP, rf)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < budget_s and reps < 400:
        prob.action_grad(XP, rf)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": reps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d evals of one C2 path (NumPy oracle port of va_ode.A_gaussian + analytic "
                      "adjoint, stand-in for pyadolc which is not installable), %.1f s" % (reps, dt)}


def run_reference(args):
    """--impl reference: the CPU path (oracle port; pyadolc/python2 are absent so the reference
    itself cannot run) on all host cores, one path per process like the reference's SGE array."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    _, Y = twin_data()
    cores = os.cpu_count() or 1
    reps = 2
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        jobs = [(Y, 1000 + c, reps) for c in range(cores)]
        for _ in range(args.warmup):
            pool.map(_cpu_worker, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    evals = cores * reps * args.steps
    val = evals / dt
    sample = ("each step = %d processes x %d evals of one C2 path each (NumPy oracle port, "
              "stand-in for pyadolc)" % (cores, reps))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "C2: Lorenz96 D=100, N_model=5001, L=40 observed, SimpsonHermite, "
                        "RF=RF0*alpha**beta at beta=%d, twin-experiment data" % BETA_EVAL,
            "paths_per_gpu": PATHS_PER_GPU, "global_paths": PATHS_PER_GPU * n_gpus,
            "unknowns_per_path": N_MODEL * D + 1, "parallelism": "paths sharded over GPUs, no collective",
            "l2_policy": "inputs larger than L2 (XP + grad = 512 MB per step per GPU vs 126 MB L2)"}


def bind_to_gpu_numa_node(local):
    """Pins this process to the CPUs NVML reports as local to its GPU, before any pinned host
    buffer is allocated: on a multi-socket box the staging buffers of the end-to-end leg are then
    first-touched on the GPU's own NUMA node instead of crossing the socket interconnect.
    VAB_BENCH_BIND=0 disables it.  Returns a short description for the JSON line."""
    if os.environ.get("VAB_BENCH_BIND", "1") == "0":
        return "off"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode() if hasattr(bus, "encode") else bus)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return "cpus %d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
    except Exception as exc:
        return "unavailable (%s)" % type(exc).__name__


# ------------------------------------------------------------------------------------ other configs
def other_configs(device):
    """Evaluation rates of the other BASELINE.json configs at their per-path sizes (device
    resident, a handful of paths): C3 = Lorenz96 D=1000, N=100000 (rk4 and SimpsonHermite), C4 =
    nnet_twin 5x100 M=1000, C5 = bar images [25,30,4] M=10000.  Reported next to the headline line;
    they are not the metric."""
    import ctypes as ct
    import torch
    from varanneal_b200 import _lib, va_nnet
    out = {}
    dev = torch.device("cuda", device)

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- C3 shape straight through the C ABI (no host copies of 0.8 GB paths)
    D3, B3 = 1000, 4
    L3 = [i for i in range(D3) if i % 5 in (0, 2)]
    for disc, N3 in (("rk4", 100000), ("SimpsonHermite", 100001)):
        ctx = _lib.Context(device, torch.cuda.current_stream(dev).cuda_stream)
        n = N3 * D3 + 1
        ld = (n + 15) // 16 * 16
        gen = torch.Generator(device=dev).manual_seed(2000)
        Y = torch.randn(N3, len(L3), dtype=torch.float64, device=dev, generator=gen)
        XP = torch.randn(B3, ld, dtype=torch.float64, device=dev, generator=gen) * 3.0
        XP[:, N3 * D3:] = K_FORCING
        G = torch.empty_like(XP)
        A = torch.zeros(B3, dtype=torch.float64, device=dev)
        pfix = torch.full((1,), K_FORCING, dtype=torch.float64, device=dev)
        desc = _lib.OdeDesc(0, _lib.DISC_IDS[disc], D3, N3, N3, 1, len(L3), 1, 1, 0, DT)
        p = lambda t: ct.c_void_p(t.data_ptr())  # noqa: E731
        _lib.check(ctx.lib.vab_ode_problem_set(ctx.h, desc, _lib.int_array(L3), _lib.int_array([0]), p(Y), None), ctx.h)
        _lib.check(ctx.lib.vab_ode_set_weights(ctx.h, RM, None, RF0, None), ctx.h)
        _lib.check(ctx.lib.vab_ode_set_fixed_params(ctx.h, p(pfix), 0), ctx.h)
        scale = ALPHA ** BETA_EVAL
        ms = timed(lambda: _lib.check(ctx.lib.vab_ode_action_grad(ctx.h, B3, p(XP), ld, scale, p(A), None, None, p(G), ld), ctx.h), 5)
        byt = B3 * 16.0 * N3 * D3 + 8.0 * N3 * len(L3)
        out["C3_%s" % disc] = {"shape": "Lorenz96 D=1000 N=%d L=400, %d paths resident" % (N3, B3),
                               "evals_per_s": B3 / ms * 1e3, "ms_per_launch": ms,
                               "algorithmic_GBps": byt / ms / 1e6, "finite": bool(torch.isfinite(A).all().item())}
        ctx.close()
        del XP, G, Y
    # ---- NN shapes through va_nnet.Annealer
    for name, st, M in (("C4", [100] * 5, 1000), ("C5", [25, 30, 4], 10000)):
        st = np.array(st)
        B4 = 64
        NDnet = int(st.sum())
        NP = int(sum(st[k] * st[k + 1] + st[k + 1] for k in range(len(st) - 1)))
        rng = np.random.RandomState(0)
        an = va_nnet.Annealer(device=device)
        an.set_structure(st)
        an.set_activation("sigmoid")
        an.set_input_data(rng.rand(M, st[0]))
        an.set_output_data(rng.rand(M, st[-1]))
        X0 = rng.rand(B4, M * NDnet)
        P0 = 0.3 * rng.randn(B4, NP)
        an.anneal_init(X0, P0, 1.1, [20.0], 1.0, 1e-2, np.arange(NP), init_to_data=False)
        an._XP[:, :an._n].copy_(torch.from_numpy(np.concatenate([X0, P0], axis=1)))
        ms = timed(lambda: an._action_grad_native(6.7), 10)
        flops = 6.0 * M * sum(st[k] * st[k + 1] for k in range(len(st) - 1)) * B4
        byt = (16.0 * M * NDnet + 16.0 * NP) * B4
        out[name] = {"shape": "va_nnet %s M=%d, %d paths" % (list(map(int, st)), M, B4),
                     "evals_per_s": B4 / ms * 1e3, "ms_per_launch": ms, "fp64_TFLOPs": flops / ms / 1e9,
                     "algorithmic_GBps": byt / ms / 1e6}
        del an
    # ---- C1: the shipped example's shape (D=20, N=161, 8 observed, 101 betas), full ladder
    from varanneal_b200 import datagen, va_ode
    L1 = [0, 2, 4, 6, 8, 10, 14, 16]
    t1, _, Y1 = datagen.lorenz96_twin(D=20, N=161, dt=0.025, k=8.17, sigma=0.5, Lidx=L1, seed=100)
    for B1 in (1, 64):
        rng = np.random.RandomState(12345)
        X0 = 20.0 * rng.rand(B1, 161, 20) - 10.0
        P0 = 4.0 * rng.rand(B1, 1) + 6.0
        an = va_ode.Annealer(device=device)
        an.set_model("lorenz96", 20)
        an.set_data(Y1, t=t1)
        t0 = time.perf_counter()
        an.anneal(X0, P0, 1.5, np.linspace(0, 100, 101), 4.0, 4e-6, L1, [0], dt_model=0.025, init_to_data=True,
                  disc="trapezoid", opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000})
        dt1 = time.perf_counter() - t0
        out["C1_B%d" % B1] = {"shape": "Lorenz96 D=20 N=161 L=8 trapezoid, 101 betas, %d path(s), Annealer.anneal()" % B1,
                              "ladder_wall_s": dt1, "nfev_total": int(an.nfev_array.sum()),
                              "converged_fraction": float(np.mean(an.exitflags == 0)),
                              "A_last_rung_min": float(np.min(an.A_array[..., -1]))}
        del an
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from varanneal_b200 import va_ode

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    bound = bind_to_gpu_numa_node(local)
    if world > 1:
        # NCCL's own log lines (e.g. "NCCL version ...") go to stderr: stdout carries the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = PATHS_PER_GPU
    _, Y = twin_data()
    X0, P0 = initial_paths(B, 1000 + rank * B)
    an = va_ode.Annealer(device=local)
    an.set_model("lorenz96", D)
    an.set_data(Y, t=DT * np.arange(N_MODEL))
    an.anneal_init(X0, P0, ALPHA, [BETA_EVAL], RM, RF0, LIDX, [0], disc="SimpsonHermite",
                   init_to_data=True, opt_args={"gtol": 1e-8, "ftol": 1e-8})
    n = an._n
    XP_host = torch.empty(B, n, dtype=torch.float64, pin_memory=True)
    XP_host[:, :N_MODEL * D] = torch.from_numpy(X0.reshape(B, -1))
    XP_host[:, N_MODEL * D:] = torch.from_numpy(P0)
    an._XP[:, :n].copy_(XP_host)
    scale = an._rf_scale()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident evals/s
    for _ in range(args.warmup):
        an._action_grad_native(scale)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = an.gpu_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        an._action_grad_native(scale)
        ev[i + 1].record(stream)
    barrier()
    launches = an.gpu_launches - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public eval seam, pinned host buffers
    for _ in range(2):
        an.A_gradA(XP_host)
    barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        A_h, G_h = an.A_gradA(XP_host)
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0) * 0.0)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * B * e2e_steps / (float(t.item()) * 1e-3)

    # ---- full ladder: 20 rungs, every rank its own 64 initial paths, one gather at the end
    ladder = None
    if not args.no_ladder:
        from varanneal_b200 import parallel
        X0l, P0l = initial_paths(B, 1000 + rank * B)
        anl = va_ode.Annealer(device=local)
        anl.set_model("lorenz96", D)
        anl.set_data(Y, t=DT * np.arange(N_MODEL))
        barrier()
        native = {}
        from varanneal_b200 import _lib as _vlib
        lib = _vlib.load()
        orig_anneal = lib.vab_anneal

        def timed_anneal(*a):
            t = time.perf_counter()
            r = orig_anneal(*a)
            native["s"] = time.perf_counter() - t
            return r
        lib.vab_anneal = timed_anneal
        t0 = time.perf_counter()
        try:
            anl.anneal(X0l, P0l, ALPHA, np.arange(N_BETA), RM, RF0, LIDX, [0], disc="SimpsonHermite",
                       init_to_data=True, opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000})
        finally:
            lib.vab_anneal = orig_anneal
        tables = np.stack([anl.action_errors_table(init=i) for i in range(B)])
        tables = parallel.gather_blocks(tables, B * world)          # the design's only collective
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall, float(anl.nfev_array.sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            tw = tt.clone()
            dist.all_reduce(tw[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(tt[1:], op=dist.ReduceOp.SUM)
            tt[0] = tw[0]
        ladder = {"wall_s": float(tt[0].item()), "device_ladder_s": native.get("s"),
                  "note": "wall_s = Annealer.anneal() from host X0/P0 to host minpaths (B, Nbeta, N*D+NP) "
                          "+ gathered tables; device_ladder_s = the vab_anneal call inside it (rank 0)",
                  "paths": B * world, "betas": N_BETA,
                  "alpha": ALPHA, "opt_args": "gtol=ftol=1e-8 (examples/Lorenz96_D20)",
                  "nfev_total": int(tt[1].item()), "evals_per_s_incl_optimizer": float(tt[1].item() / tt[0].item()),
                  "converged_fraction": float(np.mean(anl.exitflags == 0)),
                  "A_last_rung_mean": float(tables[:, -1, 1].mean()),
                  "gathered_table_shape": list(tables.shape)}
        del anl

    # the clock sampler has been running since the start of the timed region: device-resident
    # loop, end-to-end loop and the full ladder (tens of seconds under load)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peaks = {}
        pk_src = "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
            pk_src = "measured"
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        kern_ms = float(np.mean(per_step))
        achieved = algorithmic_bytes(B) / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get("ode_walk_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": pk_src,
                         "kernel": "stream_simpson_kernel<ModelL96<4>> (TMA ring, + ~3 us finalize)",
                         "algorithmic_bytes_per_launch": algorithmic_bytes(B),
                         "kernel_ms": kern_ms},
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": int(XP_host.numel() * 8),
                    "d2h_bytes_per_step": int(G_h.numel() * 8 + A_h.numel() * 8),
                    "api": "va_ode.Annealer.A_gradA(pinned XP) -> (A, grad) pinned"},
            "gpu_launches": int(launches), "clocks": clocks, "ladder": ladder, "cpu_binding": bound,
        }
        if world == 1 and not args.no_extra:
            try:
                line["other_configs"] = other_configs(local)
            except Exception as exc:                          # never lose the headline line
                line["other_configs"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single_core(Y)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ladder", action="store_true", help="skip the full-ladder leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4 / C5 evaluation-rate probes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
