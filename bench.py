#!/usr/bin/env python
"""bench.py -- action+gradient evals/s and full-ladder anneal wall time of the annealing hot path
(BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C1|C3|C4|C5]

Default workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): Lorenz96 D=100, N=5001 model
points (Simpson-Hermite needs an odd N), L=40 observed components, a batch of 64
independent paths per GPU, synthetic twin-experiment data.  One *step* = LAUNCHES_PER_STEP
back-to-back fused action+gradient evaluations of the whole batch (so that the timed region is
>= 100 ms and the clocks settle) = 64 x LAUNCHES_PER_STEP evals.  Under torchrun every rank owns
its own batch of 64 paths (weak scaling, no collective on the data path); time = max over ranks.

JSON keys: see the contract in the task statement.  ``value`` = device-resident evals/s,
``e2e`` = the same through ``va_ode.Annealer.A_gradA`` with pinned host buffers (H2D of XP and
D2H of A and grad inside the timed region), ``roofline`` = algorithmic bytes / kernel time
against MEASURED_PEAKS.json, ``cpu_baseline`` = the NumPy oracle port on one host core,
``parity_checked`` = path 0 of the *timed* launch compared with the oracle (A, ||grad||) before
anything is printed, ``ladder`` = the 20-rung anneal of the 64 paths with its per-beta parity
against the reference + SciPy golden of path 0 (``max_rel_dA_vs_cpu``).

The other BASELINE configs are separate legs (``--config``), each printing the same contract line:
  C1  the shipped example (D=20, N=161, 101 betas): ladder wall time, 1 path and 64 paths
  C3  Lorenz96 D=1000, N=100000, rk4, 128 initial paths per GPU (1024 / 8) annealed in waves
  C4  va_nnet twin network 5 x 100, M=1000: evals/s and a short ladder
  C5  va_nnet bar images [25, 30, 4], M=10000, 32 initial networks per GPU (256 / 8): ladder
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
LAUNCHES_PER_STEP = 50
METRIC = "action+gradient evals/sec (Lorenz96 D=100 N=5001 SimpsonHermite, 64 paths/GPU)"
UNIT = "evals/s"
PARITY_TOL = 1e-10


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


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def load_traffic(key):
    """DRAM bytes per launch of the kernel from this round's `ncu --set full` capture
    (profiles/traffic.json names the capture it was read from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        return t.get(key), t.get("source")
    except Exception:
        return None, None


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


_W = {}


def _cpu_init(Y, seed0):
    """Pool initialiser: every worker builds its problem and its path once (not inside the timed
    loop); the worker's index picks its path."""
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
    ident = mp.current_process()._identity
    k = (ident[0] - 1) if ident else 0
    X0, P0 = initial_paths(1, seed0 + k)
    X0[0][:, LIDX] = Y                                   # init_to_data, as the GPU arm
    _W["prob"] = _oracle_problem(Y)
    _W["XP"] = np.append(X0[0].ravel(), P0[0])
    _W["rf"] = RF0 * ALPHA ** BETA_EVAL


def _cpu_worker(reps):
    for _ in range(reps):
        _W["prob"].action_grad(_W["XP"], _W["rf"])
    return reps


def cpu_baseline_single_core(Y, budget_s=12.0):
    prob = _oracle_problem(Y)
    X0, P0 = initial_paths(1, 1000)
    X0[0][:, LIDX] = Y
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


def cpu_ladder_c1(budget_s=240.0):
    """The reference's own ladder on the host: BASELINE.json configs[0] as shipped (D=20, N=161,
    8 observed, 101 betas, trapezoid, gtol = ftol = 1e-8, seed 12345) through the oracle action +
    SciPy L-BFGS-B on one core, warm-starting rung to rung as anneal_step does (va_ode.py:707-789).
    Stops early (and says so) if the budget runs out."""
    import scipy.optimize as opt
    from oracle.ode_port import OdeProblem
    z = np.load(os.path.join(ROOT, "tests", "golden", "l96_ladder_golden.npz"))
    data = z["data"]
    L1 = [0, 2, 4, 6, 8, 10, 14, 16]
    np.random.seed(12345)
    X0 = (20.0 * np.random.rand(161 * 20) - 10.0).reshape((161, 20))
    P0 = np.array([4.0 * np.random.rand() + 6.0])
    X0[:, L1] = data[:, 1:][:, L1]
    prob = OdeProblem("lorenz96", 20, data[:, 1:][:, L1], L1, 0.025, "trapezoid", P0, [0], 4.0)
    xp = np.append(X0.ravel(), P0)
    t0 = time.perf_counter()
    nfev, done, A = 0, 0, None
    for beta in range(101):
        rf = 4e-6 * 1.5 ** beta
        r = opt.minimize(lambda v: prob.action_grad(v, rf), xp, method="L-BFGS-B", jac=True,
                         options={"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000})
        xp, A = r.x, float(r.fun)
        nfev += int(r.nfev)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    wall = time.perf_counter() - t0
    return {"config": "C1 as shipped: Lorenz96 D=20 N=161 L=8 trapezoid, 101 betas, 1 initial path, "
                      "oracle action + SciPy L-BFGS-B, 1 core",
            "wall_s": wall, "betas_done": done, "betas": 101, "nfev": nfev,
            "evals_per_s_incl_optimizer": nfev / wall, "A_last": A}


def c2_slice_recorded():
    """CPU wall time of one C2 initialisation's 20-rung ladder (reference anneal() + SciPy on the
    oracle action), recorded when the golden fixture was generated in the build container
    (tests/golden/make_ladder_golden.py) -- ~20 minutes of one core, too long to repeat here."""
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "c2_slice_ladder_golden.npz"))
        return {"wall_s_one_path_one_core": float(z["meta"][5]), "nfev": int(z["counts"][:, 1].sum()),
                "note": "recorded at fixture generation (8-core build container), not measured in this run"}
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the CPU path (oracle port; pyadolc/python2 are absent so the reference
    itself cannot run) on all host cores, one path per process like the reference's SGE array.
    The problem and the path of every worker are built once, outside the timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    _, Y = twin_data()
    cores = os.cpu_count() or 1
    reps = 2
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(Y, 1000)) as pool:
        jobs = [reps] * cores
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    evals = cores * reps * args.steps
    val = evals / dt
    sample = ("each step = %d processes x %d evals of one C2 path each (NumPy oracle port, "
              "stand-in for pyadolc); problem set-up outside the timed region" % (cores, reps))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_ladder:
        line["ladder"] = cpu_ladder_c1()
        line["ladder"]["c2_slice"] = c2_slice_recorded()
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "C2: Lorenz96 D=100, N_model=5001, L=40 observed, SimpsonHermite, "
                        "RF=RF0*alpha**beta at beta=%d, twin-experiment data" % BETA_EVAL,
            "paths_per_gpu": PATHS_PER_GPU, "global_paths": PATHS_PER_GPU * n_gpus,
            "launches_per_step": LAUNCHES_PER_STEP,
            "unknowns_per_path": N_MODEL * D + 1, "parallelism": "paths sharded over GPUs, no collective",
            "l2_policy": "inputs larger than L2 (XP + grad = 512 MB per launch per GPU vs 126 MB L2)"}


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


# ------------------------------------------------------------------------------------ helpers
class Dist(object):
    """Rank bookkeeping + the barrier / max-over-ranks the contract asks for."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        torch.cuda.set_device(self.local)
        self.bound = bind_to_gpu_numa_node(self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the JSON line only
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, value, op="max"):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def oracle_check(A_dev, G_dev, Ar, gr):
    """Relative deviation of the device's (A, grad) of one path from the oracle's."""
    ea = abs(float(A_dev) - Ar) / abs(Ar)
    eg = float(np.max(np.abs(np.asarray(G_dev) - gr)) / np.max(np.abs(gr)))
    return {"rel_err_A": ea, "rel_err_grad": eg, "grad_norm": float(np.linalg.norm(gr)), "tol": PARITY_TOL,
            "ok": bool(ea <= PARITY_TOL and eg <= PARITY_TOL)}


def base_line(metric, unit, value, dd, args, ms_per_step, config, dtype="f64", scaling="weak"):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": dd.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config}


# ------------------------------------------------------------------------------------ C2 (default)
def leg_c2(args, dd):
    import torch
    from varanneal_b200 import va_ode
    rank, world, local = dd.rank, dd.world, dd.local
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
    XP_host[:, :N_MODEL * D] = torch.from_numpy(X0.reshape(B, -1))      # X0 carries the data (init_to_data)
    XP_host[:, N_MODEL * D:] = torch.from_numpy(P0)
    an._XP[:, :n].copy_(XP_host)
    scale = an._rf_scale()
    stream = torch.cuda.current_stream()
    K = LAUNCHES_PER_STEP

    # ---- device-resident evals/s
    for _ in range(args.warmup):
        for _ in range(K):
            an._action_grad_native(scale)
    dd.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = an.gpu_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    dd.barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        for _ in range(K):
            an._action_grad_native(scale)
        ev[i + 1].record(stream)
    dd.barrier()
    launches = an.gpu_launches - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms_max = dd.reduce(total_ms, "max")
    value = world * B * K * args.steps / (total_ms_max * 1e-3)

    # ---- the timed launch against the oracle (rank 0, path 0): its outputs are still on the device
    parity = None
    if rank == 0:
        prob = _oracle_problem(Y)
        xp0 = XP_host[0].numpy()
        Ar, gr = prob.action_grad(xp0, RF0 * ALPHA ** BETA_EVAL)
        parity = oracle_check(an._A[0].item(), an._G[0, :n].cpu().numpy(), Ar, gr)
        parity["what"] = "path 0 of the last timed launch vs oracle.ode_port (A and max-norm gradient)"
        if not parity["ok"]:
            raise SystemExit("bench: the timed launch disagrees with the oracle: %r" % (parity,))

    # ---- end to end through the public eval seam, pinned host buffers
    for _ in range(2):
        an.A_gradA(XP_host)
    dd.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        A_h, G_h = an.A_gradA(XP_host)
    e1.record(stream)
    dd.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)      # A_gradA returns after its own synchronize: both clocks cover the region
    e2e_val = world * B * e2e_steps / (dd.reduce(e2e_ms, "max") * 1e-3)

    # ---- full ladder: 20 rungs, every rank its own 64 initial paths, one gather at the end
    ladder = None
    if not args.no_ladder:
        from varanneal_b200 import parallel
        from varanneal_b200 import _lib as _vlib
        X0l, P0l = initial_paths(B, 1000 + rank * B)
        anl = va_ode.Annealer(device=local)
        anl.set_model("lorenz96", D)
        anl.set_data(Y, t=DT * np.arange(N_MODEL))
        dd.barrier()
        native = {}
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
        local_tab0 = tables[0].copy()
        tables = parallel.gather_blocks(tables, B * world)          # the design's only collective
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        wall_max = dd.reduce(wall, "max")
        nfev_tot = dd.reduce(float(anl.nfev_array.sum()), "sum")
        ladder = {"wall_s": wall_max, "device_ladder_s": native.get("s"),
                  "note": "wall_s = Annealer.anneal() from host X0/P0 to host minpaths (B, Nbeta, N*D+NP) "
                          "+ gathered tables; device_ladder_s = the vab_anneal call inside it (rank 0)",
                  "paths": B * world, "betas": N_BETA,
                  "alpha": ALPHA, "opt_args": "gtol=ftol=1e-8 (examples/Lorenz96_D20)",
                  "nfev_total": int(nfev_tot), "evals_per_s_incl_optimizer": nfev_tot / wall_max,
                  "converged_fraction": float(np.mean(anl.exitflags == 0)),
                  "A_last_rung_mean": float(tables[:, -1, 1].mean()),
                  "gathered_table_shape": list(tables.shape),
                  "graph_replayed_cycles": anl._ctx.graph_launches}
        if rank == 0:
            # per-beta parity of path 0 (seed 1000) against the reference anneal() + SciPy golden
            try:
                z = np.load(os.path.join(ROOT, "tests", "golden", "c2_slice_ladder_golden.npz"))
                ref = z["table"][:, 1]
                rel = np.abs(local_tab0[:, 1] - ref) / np.abs(ref)
                ladder["max_rel_dA_vs_cpu"] = float(rel.max())
                ladder["rel_dA_vs_cpu_per_beta"] = [float(v) for v in rel]
                signed = (local_tab0[:, 1] - ref) / np.abs(ref)
                ladder["device_lower_than_cpu_on_rungs"] = [int(i) for i in np.where(signed < -1e-6)[0]]
                ladder["max_rel_dA_where_device_is_higher"] = float(np.max(np.maximum(signed, 0.0)))
                ladder["rel_dA_note"] = ("(A_dev - A_cpu) / A_cpu per rung; on the rungs listed in "
                                         "device_lower_than_cpu_on_rungs the device's minimum is the *lower* one "
                                         "(multi-modal stretch of the ladder, DESIGN.md 2.1)")
                if "table_ulp1" in z.files:
                    band = np.abs(z["table_ulp1"][:, 1] - ref) / np.abs(ref)
                    ladder["reference_self_spread_max"] = float(band.max())
                    ladder["reference_self_spread_note"] = (
                        "the same reference + SciPy ladder started from X0 (1 + 2^-52): how far the CPU "
                        "path drifts from itself (shipped tolerances gtol = ftol = 1e-8 stop in flat valleys)")
                ladder["cpu_c2_slice"] = c2_slice_recorded()
            except Exception as exc:
                ladder["max_rel_dA_vs_cpu"] = None
                ladder["parity_error"] = repr(exc)
        del anl

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    peaks, pk_src = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern_ms = float(np.mean(per_step)) / K
    achieved = algorithmic_bytes(B) / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = load_traffic("ode_walk_bytes_per_launch")
    line = base_line(METRIC, UNIT, value, dd, args, total_ms_max / args.steps, workload_config(world))
    line.update({
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": pk_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "kernel": "stream_simpson_kernel<ModelL96<4>> (TMA ring, + ~3 us finalize)",
                     "algorithmic_bytes_per_launch": algorithmic_bytes(B),
                     "kernel_ms": kern_ms},
        "e2e": {"value": e2e_val, "unit": UNIT,
                "h2d_bytes_per_step": int(XP_host.numel() * 8),
                "d2h_bytes_per_step": int(G_h.numel() * 8 + A_h.numel() * 8),
                "api": "va_ode.Annealer.A_gradA(pinned XP) -> (A, grad) pinned, one call = one batch of 64 evals"},
        "gpu_launches": int(launches), "clocks": clocks, "ladder": ladder, "cpu_binding": dd.bound,
        "parity_checked": bool(parity and parity["ok"]), "parity": parity,
        "fp64_tflops_builder_measured": an._ctx.fp64_peak_tflops(),
    })
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_single_core(Y)
    return line


# ------------------------------------------------------------------------------------ C1
def leg_c1(args, dd):
    """The shipped example (BASELINE.json configs[0]): shipped data file, 101 betas, trapezoid,
    gtol = ftol = 1e-8, seed 12345.  value = evals/s including the optimiser for the 64-path batch;
    the 1-path ladder is what the reference's user sees."""
    import torch
    from varanneal_b200 import va_ode
    z = np.load(os.path.join(ROOT, "tests", "golden", "l96_ladder_golden.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_shipped_ladder_golden.npz"))
    data = z["data"]
    L1 = [0, 2, 4, 6, 8, 10, 14, 16]
    out = {}
    sampler = ClockSampler(dd.local)
    if dd.rank == 0:
        sampler.start()
    for B1 in (1, 32, 64):
        rng = np.random.RandomState(12345 + dd.rank)
        X0 = 20.0 * rng.rand(B1, 161, 20) - 10.0
        P0 = 4.0 * rng.rand(B1, 1) + 6.0
        if B1 == 1 and dd.rank == 0:
            X0[0], P0[0] = g["trapezoid/X0"], g["trapezoid/P0"]
        an = va_ode.Annealer(device=dd.local)
        an.set_model("lorenz96", 20)
        an.set_data(data[:, 1:][:, L1], t=data[:, 0])
        dd.barrier()
        t0 = time.perf_counter()
        an.anneal(X0, P0, 1.5, np.linspace(0, 100, 101), 4.0, 4e-6, L1, [0], dt_model=0.025, init_to_data=True,
                  disc="trapezoid", opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000})
        torch.cuda.synchronize()
        wall = dd.reduce(time.perf_counter() - t0, "max")
        nfev = dd.reduce(float(an.nfev_array.sum()), "sum")
        out[B1] = {"ladder_wall_s": wall, "nfev_total": int(nfev), "paths": B1 * dd.world,
                   "evals_per_s_incl_optimizer": nfev / wall,
                   "converged_fraction": float(np.mean(an.exitflags == 0)),
                   "launches": an.gpu_launches, "graph_replayed_cycles": an._ctx.graph_launches}
        if B1 == 1 and dd.rank == 0:
            ref = g["trapezoid/table"][:, 1]
            rel = np.abs(an.A_array[0] - ref) / np.abs(ref)
            band = np.abs(g["trapezoid/table_ulp1"][:, 1] - ref) / np.abs(ref) if "trapezoid/table_ulp1" in g.files else None
            out[B1]["max_rel_dA_vs_cpu"] = float(rel.max())
            out[B1]["median_rel_dA_vs_cpu"] = float(np.median(rel))
            if band is not None:
                out[B1]["reference_self_spread_max"] = float(band.max())
            out[B1]["cpu_reference_wall_s"] = float(g["trapezoid/meta"][5])
    clocks = sampler.stop() if dd.rank == 0 else None
    if dd.rank != 0:
        return None
    v = out[64]
    line = base_line("anneal evals/sec incl. optimiser (C1: shipped Lorenz96 D=20 example, 101 betas, 64 paths/GPU)",
                     UNIT, v["evals_per_s_incl_optimizer"], dd, args, 1e3 * v["ladder_wall_s"],
                     {"workload": "C1: examples/Lorenz96_D20 as shipped (D=20, N=161, L=8, trapezoid, 101 betas, "
                                  "gtol=ftol=1e-8), full ladder = one step", "paths_per_gpu": 64,
                      "l2_policy": "problem is 26 KB per path: L2/launch-latency bound by nature"})
    line.update({"steps": 1, "warmup": 0, "roofline": None, "e2e": {"value": v["evals_per_s_incl_optimizer"], "unit": UNIT,
                 "h2d_bytes_per_step": 64 * 3221 * 8, "d2h_bytes_per_step": 64 * 101 * 3221 * 8,
                 "api": "va_ode.Annealer.anneal(): host X0/P0 in, host minpaths out"},
                 "gpu_launches": v["launches"], "clocks": clocks, "one_path": out[1], "batch32": out[32], "batch": out[64]})
    return line


# ------------------------------------------------------------------------------------ C3
def leg_c3(args, dd):
    """BASELINE.json configs[2]: Lorenz96 D=1000, N=100000, rk4, 1024 initial paths over 8 GPUs =
    128 per GPU, annealed in memory-sized waves (a path's minimiser state is 26 vectors x 0.8 GB).
    A 2-rung ladder with maxiter iterations per rung is the bounded sample that is timed; the
    initial paths are drawn on the device (seeds 2000 + b) -- 128 x 0.8 GB do not belong in host
    memory; keep_paths='none' (the per-rung parameter estimates and the action table are kept)."""
    import torch
    from varanneal_b200 import datagen, va_ode
    D3, N3 = 1000, int(os.environ.get("VAB_C3_N", "100000"))
    Bg = int(os.environ.get("VAB_C3_PATHS", "128"))
    maxiter = int(os.environ.get("VAB_C3_MAXITER", "6"))
    L3 = [i for i in range(D3) if i % 5 in (0, 2)]
    dev = torch.device("cuda", dd.local)
    t_gen = time.perf_counter()
    _, _, Y = datagen.lorenz96_twin(D=D3, N=N3, dt=DT, k=K_FORCING, sigma=0.5, Lidx=L3, seed=100, transient=1000)
    t_gen = time.perf_counter() - t_gen

    def x0_block(b0, b1):
        out = torch.empty(b1 - b0, N3, D3, dtype=torch.float64, device=dev)
        for b in range(b0, b1):
            gen = torch.Generator(device=dev).manual_seed(2000 + dd.rank * Bg + b)
            out[b - b0] = 20.0 * torch.rand(N3, D3, dtype=torch.float64, device=dev, generator=gen) - 10.0
        return out

    P0 = np.array([[4.0 * np.random.RandomState(2000 + dd.rank * Bg + b).rand() + 6.0] for b in range(Bg)])
    an = va_ode.Annealer(device=dd.local)
    an.keep_paths = 'none'
    an.set_model("lorenz96", D3)
    an.set_data(Y, t=DT * np.arange(N3))
    sampler = ClockSampler(dd.local)
    if dd.rank == 0:
        sampler.start()
    dd.barrier()
    t0 = time.perf_counter()
    an.anneal(x0_block, P0, ALPHA, [0, 1], RM, RF0, L3, [0], disc="rk4", init_to_data=True,
              opt_args={"gtol": 0.0, "ftol": 0.0, "maxiter": maxiter})      # the iteration cap alone ends a rung
    torch.cuda.synchronize()
    wall = dd.reduce(time.perf_counter() - t0, "max")
    nfev = dd.reduce(float(an.nfev_array.sum()), "sum")
    # ---- evaluation rate of the resident wave + oracle check of one path of that launch
    Bw = an._B
    an._load_wave(0, Bw)
    scale = ALPHA ** 1.0
    for _ in range(2):
        an._action_grad_native(scale)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        an._action_grad_native(scale)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    parity = None
    if dd.rank == 0 and not args.no_cpu:
        from oracle.ode_port import OdeProblem
        prob = OdeProblem("lorenz96", D3, Y, L3, DT, "rk4", [K_FORCING], [0], RM)
        xp0 = an._XP[0, :an._n].cpu().numpy()
        tcpu = time.perf_counter()
        Ar, gr = prob.action_grad(xp0, RF0 * scale)
        tcpu = time.perf_counter() - tcpu
        parity = oracle_check(an._A[0].item(), an._G[0, :an._n].cpu().numpy(), Ar, gr)
        parity["what"] = "path 0 of the timed rk4 launch (n = %d) vs oracle.ode_port" % an._n
        parity["oracle_eval_s_one_core"] = tcpu
    clocks = sampler.stop() if dd.rank == 0 else None
    if dd.rank != 0:
        return None
    peaks, pk_src = load_peaks()
    byt = Bw * 16.0 * N3 * D3 + 8.0 * N3 * len(L3)
    fp64_peak = an._ctx.fp64_peak_tflops()
    flops = 100.0 * N3 * D3 * Bw                      # ~100 flop per element-row
    ach = byt / ms / 1e6
    line = base_line("anneal evals/sec incl. optimiser (C3: Lorenz96 D=1000 N=%d rk4, %d paths/GPU in waves)" % (N3, Bg),
                     UNIT, nfev / wall, dd, args, 1e3 * wall,
                     {"workload": "C3: Lorenz96 D=1000, N_model=%d, L=400, rk4, 2-rung ladder (alpha 2.5, beta 0..1), "
                                  "exactly maxiter=%d L-BFGS iterations per rung (gtol = ftol = 0: with n = 1e8 unknowns the shipped "
                                  "gtol = 1e-8 is met by the very first gradient; bounded sample), one step = the whole job" % (N3, maxiter),
                      "paths_per_gpu": Bg, "global_paths": Bg * dd.world, "resident_paths_per_wave": Bw,
                      "waves_per_gpu": an.n_waves, "keep_paths": "none",
                      "l2_policy": "0.8 GB per vector: nothing fits L2"})
    line.update({"steps": 1, "warmup": 0,
                 "roofline": {"bound": "hbm", "achieved": ach, "peak": float(peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
                              "frac": ach / float(peaks.get("hbm_gbs", 6650.0)), "traffic": None, "peak_source": pk_src,
                              "kernel": "stream_rk4_kernel (window mode), %d resident paths" % Bw, "kernel_ms": ms,
                              "algorithmic_bytes_per_launch": byt,
                              "fp64": {"achieved_tflops": flops / ms / 1e9, "peak_tflops_builder_measured": fp64_peak,
                                       "frac": flops / ms / 1e9 / fp64_peak if fp64_peak else None,
                                       "flops_per_element": 100}},
                 "e2e": {"value": nfev / wall, "unit": UNIT, "h2d_bytes_per_step": int(Y.nbytes + P0.nbytes),
                         "d2h_bytes_per_step": int(an.A_array.nbytes * 3 + an.params_array.nbytes),
                         "api": "va_ode.Annealer.anneal(X0=callable drawing the paths on the device, keep_paths='none')"},
                 "gpu_launches": an.gpu_launches, "clocks": clocks, "ladder_wall_s": wall, "nfev_total": int(nfev),
                 "eval_rate_resident_wave": Bw / ms * 1e3, "parity_checked": bool(parity and parity["ok"]),
                 "parity": parity, "twin_data_generation_s": t_gen,
                 "A_last_rung_mean": float(an.A_array[:, -1].mean()),
                 "finite": bool(np.all(np.isfinite(an.A_array)))})
    return line


# ------------------------------------------------------------------------------------ C4 / C5
def _nn_problem(name, rank):
    from varanneal_b200 import datagen
    if name == "C4":
        st = np.array([100] * 5)
        M = 1000
        (W, b), = datagen.nnet_twin_params(st, seed=17439860)
        din, dout, _ = datagen.nnet_twin_io(W, b, M, sigma=0.005, seed=43650832)
        RMn = 1.0 / 0.005 ** 2
    else:
        st = np.array([25, 30, 4])
        M = 10000
        data, labels = datagen.bar_images(dim=5, Nsets=2500, imagetype="centered", seed=85964309)
        din, dout = data[:M], labels[:M].astype(np.float64)
        RMn = 1.0
    NDnet = int(st.sum())
    RF0n = 1e-8 * RMn * float(NDnet - st[0]) / float(st[0] + st[-1])     # nnet_twin_anneal.py:46-50
    return st, M, din, dout, RMn, RF0n


def leg_nn(args, dd, name):
    """BASELINE.json configs[3] / [4]: va_nnet evaluation rate (value) with the timed launch checked
    against the oracle, and a ladder over the examples' beta range (alpha 1.1; every 24th of the
    436 betas for C4, all 436 for C5) with every weight estimated and biases fixed at 0
    (nnet_twin_anneal.py:101-119)."""
    import torch
    from varanneal_b200 import va_nnet
    st, M, din, dout, RMn, RF0n = _nn_problem(name, dd.rank)
    B = 64 if name == "C4" else 32                     # C5: 256 initial networks / 8 GPUs
    NDnet = int(st.sum())
    NP = int(sum(st[k] * st[k + 1] + st[k + 1] for k in range(len(st) - 1)))
    Pidx, off = [], 0
    for k in range(len(st) - 1):
        Pidx.extend(range(off, off + st[k] * st[k + 1]))
        off += st[k] * st[k + 1] + st[k + 1]
    Pidx = np.array(Pidx)
    rng = np.random.RandomState(89072545 + dd.rank)
    X0 = rng.rand(B, M * NDnet)
    P0 = np.zeros((B, NP))
    P0[:, Pidx] = (2.0 * rng.rand(B, len(Pidx)) - 1.0) / 10.0
    # C4: every 24th beta of the example's ladder and at most 500 iterations per rung keep the leg
    # within a minute (with 40 000 free weights the early rungs need thousands of iterations each);
    # C5: the example's 436 betas
    betas = np.arange(0.0, 436.0, 24.0) if name == "C4" else np.arange(0.0, 436.0, 1.0)
    maxit = 500 if name == "C4" else 2000
    an = va_nnet.Annealer(device=dd.local)
    an.set_structure(st)
    an.set_activation("sigmoid")
    an.set_input_data(din)
    an.set_output_data(dout)
    an.anneal_init(X0, P0, 1.1, [100.0], RMn, RF0n, Pidx, init_to_data=True)
    an._load_wave(0, B)
    scale = 1.1 ** 100.0
    sampler = ClockSampler(dd.local)
    if dd.rank == 0:
        sampler.start()
    K = 20
    for _ in range(max(args.warmup, 3)):
        an._action_grad_native(scale)
    dd.barrier()
    l0 = an.gpu_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps * K):
        an._action_grad_native(scale)
    e1.record()
    dd.barrier()
    launches = an.gpu_launches - l0
    ms_tot = dd.reduce(e0.elapsed_time(e1), "max")
    ms = ms_tot / (args.steps * K)
    value = dd.world * B / ms * 1e3
    parity = None
    if dd.rank == 0:
        from oracle.nnet_port import NnetProblem
        prob = NnetProblem(st, din, dout, None, P0[0], Pidx, RMn)
        xp0 = an._XP[0, :an._n].cpu().numpy()
        Ar, gr = prob.action_grad(xp0, RF0n * scale)
        parity = oracle_check(an._A[0].item(), an._G[0, :an._n].cpu().numpy(), Ar, gr)
        parity["what"] = "path 0 of the timed launch vs oracle.nnet_port"
        if not parity["ok"]:
            raise SystemExit("bench: the timed %s launch disagrees with the oracle: %r" % (name, parity))
    # the same launches with every contraction on tcgen05 / TMEM / TMA (VAB_NN_TCGEN05=1: Ozaki-split
    # int8 GEMMs, csrc/ozaki_gemm.cu) -- the opt-in path, timed and oracle-checked beside the default one
    tc = None
    if dd.rank == 0 and int(st.max()) <= 128:
        os.environ["VAB_NN_TCGEN05"] = "1"
        try:
            for _ in range(3):
                an._action_grad_native(scale)
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(K):
                an._action_grad_native(scale)
            t1e.record()
            torch.cuda.synchronize()
            ms_tc = t0e.elapsed_time(t1e) / K
            ptc = oracle_check(an._A[0].item(), an._G[0, :an._n].cpu().numpy(), Ar, gr)
            tc = {"kernel_ms": ms_tc, "evals_per_s": B / ms_tc * 1e3, "kernel_family": an._ctx.nn_kernel_family,
                  "parity": ptc, "kernels": "ozaki_slice_kernel + ozaki_mma_kernel (tcgen05.mma.kind::i8, TMEM accumulators, "
                                            "cp.async.bulk.tensor operands) + nn_tc_epilogue_kernel + nn_fix_kernel",
                  "note": "opt-in (VAB_NN_TCGEN05=1); the default path stays on the fp64 tensor pipe, which is faster at these widths"}
            if not ptc["ok"] or tc["kernel_family"] != 5:
                raise SystemExit("bench: the tcgen05 %s launch disagrees with the oracle: %r" % (name, tc))
        finally:
            os.environ.pop("VAB_NN_TCGEN05", None)
    ladder = None
    if not args.no_ladder:
        anl = va_nnet.Annealer(device=dd.local)
        anl.set_structure(st)
        anl.set_activation("sigmoid")
        anl.set_input_data(din)
        anl.set_output_data(dout)
        anl.keep_paths = 'last'          # all rungs of all paths would be 10 GB (C4) / 66 GB (C5) of host memory
        dd.barrier()
        t0 = time.perf_counter()
        anl.anneal(X0.copy(), P0.copy(), 1.1, betas, RMn, RF0n, Pidx, init_to_data=True,
                   opt_args={"gtol": 1e-12, "ftol": 1e-12, "maxfun": 1000000, "maxiter": maxit})
        torch.cuda.synchronize()
        wall = dd.reduce(time.perf_counter() - t0, "max")
        nfev = dd.reduce(float(anl.nfev_array.sum()), "sum")
        ladder = {"wall_s": wall, "paths": B * dd.world, "betas": len(betas), "nfev_total": int(nfev),
                  "evals_per_s_incl_optimizer": nfev / wall, "converged_fraction": float(np.mean(anl.exitflags == 0)),
                  "A_first_last_mean": [float(anl.A_array[:, 0].mean()), float(anl.A_array[:, -1].mean())],
                  "keep_paths": anl.keep_paths, "opt_args": "gtol=ftol=1e-12 (examples/nnet_twin), maxiter %d per rung" % maxit}
    clocks = sampler.stop() if dd.rank == 0 else None
    if dd.rank != 0:
        return None
    peaks, pk_src = load_peaks()
    flops = 6.0 * M * sum(st[k] * st[k + 1] for k in range(len(st) - 1)) * B
    byt = (16.0 * M * NDnet + 16.0 * NP) * B
    fp64_peak = an._ctx.fp64_peak_tflops()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf = flops / ms / 1e9
    gb = byt / ms / 1e6
    if name == "C4":
        dmma_peak = an._ctx.fp64_dmma_peak_tflops()
        roof = {"bound": "tensor", "achieved": tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": tf / dmma_peak if dmma_peak else None,
                "traffic": None, "peak_source": "builder-measured rate of mma.sync.m8n8k4.f64 (vab_measure_fp64_dmma_peak), the instruction the "
                                                "contractions issue; fp64 FMA rate on the CUDA cores: %.1f TFLOP/s" % fp64_peak,
                "kernel": "nn_fb_kernel + nn_gw_kernel + nn_fix_kernel (fp64 tensor pipe)", "kernel_ms": ms,
                "algorithmic_flops_per_launch": flops, "hbm_GBps_algorithmic": gb}
    else:
        roof = {"bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm, "traffic": None,
                "peak_source": pk_src, "kernel": "nn_fba_kernel + nn_gw_kernel", "kernel_ms": ms,
                "algorithmic_bytes_per_launch": byt, "fp64_TFLOPs": tf}
    metric = "action+gradient evals/sec (%s: va_nnet %s M=%d, %d paths/GPU)" % (name, list(map(int, st)), M, B)
    line = base_line(metric, UNIT, value, dd, args, ms * K,
                     {"workload": "%s: va_nnet structure %s, M=%d examples, sigmoid, all weights estimated, RF at beta=100; "
                                  "one step = %d launches" % (name, list(map(int, st)), M, K),
                      "paths_per_gpu": B, "global_paths": B * dd.world,
                      "l2_policy": "XP + grad = %.0f MB per launch vs 126 MB L2" % (2 * B * an._n * 8 / 1e6)})
    if tc is not None:
        tc["fp64_equiv_TFLOPs"] = flops / tc["kernel_ms"] / 1e9
    line.update({"roofline": roof, "tcgen05": tc, "gpu_launches": int(launches), "clocks": clocks, "ladder": ladder,
                 "parity_checked": bool(parity and parity["ok"]), "parity": parity,
                 "e2e": {"value": ladder["evals_per_s_incl_optimizer"] if ladder else None, "unit": UNIT,
                         "h2d_bytes_per_step": int(X0.nbytes + P0.nbytes), "d2h_bytes_per_step": int(X0.nbytes),
                         "api": "va_nnet.Annealer.anneal(): host X0/P0 in, host results out (rate incl. optimiser)"}})
    return line


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    dd = Dist(args)
    leg = {"C2": leg_c2, "C1": leg_c1, "C3": leg_c3,
           "C4": lambda a, d: leg_nn(a, d, "C4"), "C5": lambda a, d: leg_nn(a, d, "C5")}[args.config]
    line = leg(args, dd)
    if dd.rank == 0:
        print(json.dumps(line))
    dd.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / CPU oracle legs")
    ap.add_argument("--no-ladder", action="store_true", help="skip the full-ladder leg")
    ap.add_argument("--no-extra", action="store_true", help="(kept for compatibility; the other configs are --config legs now)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
