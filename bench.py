#!/usr/bin/env python
"""bench.py — headline benchmark of the EOL-Cloth hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sheet1024|sheet256|ensemble64]

A "step" is one Forces::fill (f, M, MDK into the fixed CSR pattern) of one synthetic cloth sheet per GPU.
Default workload = BASELINE.json configs[3]: regular2 1024x1024 sheet (2,093,058 triangles + 3,137,541 interior edges
= 5,230,599 element assemblies per fill).  N > 1 runs one independent replica (own state seed) per GPU — a single mesh
does not shard: weak scaling, no collective on the data path; NCCL is only used for the barrier and the
max-over-ranks of the device time.

  value        element assemblies / s with x, X resident in HBM (device-pointer C-ABI entry, CUDA events on the ctx stream)
  e2e          the same through the host-buffer C-ABI entry (eolc_forces_fill_ex) with page-locked buffers from eolc_host_alloc:
               x/X H2D + f/MDK D2H every step; M is copied only when it changed (it depends on X and the density only) —
               the steady step of a simulation without remeshing; the full fill (M recomputed and copied) is reported beside it
  roofline     algorithmic bytes of one fill (24N+16N+12F+16Ei+24N+8nnz(M)+8nnz(MDK)) / device time of one fill
  cpu_baseline the reference's OWN compiled Forces::fill (oracle/_ref/libforces_ref.so, kind "reference"), 1 thread, 256^2 sample;
               --impl reference runs that code as one instance per host core on strips of the same 1024^2 sheet
  ensemble     BASELINE configs[4] at every N: 4096 independent 64x64 scenes sharded contiguously over the ranks (strong
               scaling, no collective), batched Forces::fill + batched CD2 narrow phase per step
  cd           secondary metric contacts/s: CD2 on the 512x512 box scene (BASELINE configs[2]), device + D2H of the list
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAT = (0.05, 50.0, 0.01, 1.0e-5, 0.0, 1.0)   # simulationSettings.json:27-33
GRAV = (0.0, 0.0, -9.8)
H = 0.5e-2
METRIC = "element f+K assemblies/sec"
UNIT = "elements/s"


# ---------------------------------------------------------------------------------------------- helpers (CPU-testable)
def shard_range(n_units, rank, world):
    """Contiguous shard [lo, hi) of n_units independent scenes for `rank` of `world`."""
    per, rem = divmod(n_units, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(v, device):
    import torch
    d = _dist()
    if d is None:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(v, device):
    import torch
    d = _dist()
    if d is None:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def algorithmic_bytes(N, F, Ei, nnzM, nnzK):
    """x + X + face idx + edge stencil idx + f + M values + MDK values."""
    return 24 * N + 16 * N + 12 * F + 16 * Ei + 24 * N + 8 * nnzM + 8 * nnzK


def measured_traffic(workload):
    """DRAM bytes per launch from the committed ncu capture of the same command (profiles/traffic.json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return d[workload]["bytes"], d[workload]["source"]
    except Exception:
        return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.p is not None:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        try:
            self.f.flush()
            rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
            sm = [float(r[1]) for r in rows]
            out["sm_mhz"] = statistics.median(sm) if sm else None
            out["sm_max_mhz"] = float(rows[0][2]) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = set()
            for r in rows:
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            out["reasons"] = sorted(reasons)
            out["samples"] = len(rows)
        except Exception as e:  # clocks are evidence, never fatal
            out["error"] = str(e)
        finally:
            try:
                os.unlink(self.f.name)
            except Exception:
                pass
        return out


def sheet_counts(n):
    """regular2 n x n sheet: N, F, E, Ei, nnz(M), nnz(MDK) in closed form (pattern(M) = node pairs sharing a face,
    pattern(MDK) adds the opposite-vertex pairs of interior edges; 3x3 blocks)."""
    N, F = n * n, 2 * (n - 1) ** 2
    Ed = 3 * (n - 1) ** 2 + 2 * (n - 1)
    Ei = Ed - 4 * (n - 1)
    return N, F, Ed, Ei, 9 * (N + 2 * Ed), 9 * (N + 2 * Ed + 2 * Ei)


def workload_config(workload, world):
    """The `config` object of the JSON line — the same for both arms (the driver compares them)."""
    n = {"sheet1024": 1024, "sheet256": 256, "ensemble64": 64}[workload]
    N, F, _, Ei, nnzM, nnzK = sheet_counts(n)
    if workload == "ensemble64":
        per = "4096 scenes sharded over %d rank(s)" % world
        par = "ensemble x%d" % world
        ws = algorithmic_bytes(N, F, Ei, nnzM, nnzK) * (4096 // world)
    else:
        per, par, ws = 1, "replicas x%d" % world, algorithmic_bytes(N, F, Ei, nnzM, nnzK)
    return {"workload": workload_name(workload), "mesh": f"regular2 n={n}", "scenes_per_gpu": per, "nodes": N, "faces": F,
            "interior_edges": Ei, "nnz_M": nnzM, "nnz_MDK": nnzK, "parallelism": par,
            "l2_policy": "working set per step (%.0f MB) exceeds the 126 MB L2" % (ws / 1e6)}


def make_sheet(n, seed):
    import eol_cloth_b200 as E
    X, fn = E.meshgen.regular2(n)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=seed)
    return X, fn, es, x


# ---------------------------------------------------------------------------------------------- CPU legs
def cpu_forces_sample(n=256, repeats=3):
    """Oracle Forces::fill on a bounded sample (regular2 n x n), single thread, best of `repeats`."""
    from oracle import oracle as O
    X, fn, es, x = make_sheet(n, 0)
    elements = fn.shape[0] + int((es[:, 3] >= 0).sum())
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        O.forces_fill(fn, es, x, X, MAT, GRAV, H)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return elements / best, elements, best


def cpu_baseline_leg():
    """The reference CPU path beside the GPU number, on rank 0 at N = 1: the reference's OWN compiled Forces::fill
    (oracle/_ref/libforces_ref.so, kind "reference"), one thread like the reference program, regular2 256 x 256 sample; the oracle
    port (kind "port") on the same sample next to it, and alone when the reference library is missing."""
    from oracle import oracle as O
    vp, el, dtp = cpu_forces_sample(256, 2)
    port = {"value": vp, "unit": UNIT, "cores": 1, "kind": "port", "seconds_per_fill": dtp,
            "note": "reference Compute*.cpp object code + restated Forces.cpp glue incl. triplets + setFromTriplets"}
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libforces_ref.so")):
        return dict(port, sample=f"regular2 n=256 sheet ({el} elements), best of 2, {dtp:.2f} s per fill")
    X, fn = O.sheet_regular2(256)
    x = O.drape_state(X, seed=0)
    dt = O.ref_forces_seconds(fn, x, X, MAT, GRAV, H, instances=1, repeats=2)
    return {"value": el / dt, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"regular2 n=256 sheet ({el} elements), Forces::fill alone (mesh built before), best of 2, {dt:.2f} s per fill",
            "note": "the reference's own Forces.cpp + UtilEOL.cpp + Compute*.cpp + ArcSim mesh code compiled unmodified against oracle/mini_eigen "
                    "(oracle/_ref/libforces_ref.so); single thread, as the reference program runs",
            "port": port}


def mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return None


def run_reference(args):
    """--impl reference: THE REFERENCE'S OWN CODE on the same metric and config, on the box's host cores.  Only oracle/ is loaded — no
    product code.  oracle/_ref/libforces_ref.so is /root/reference/src/Forces.cpp + UtilEOL.cpp + conversions.cpp + Compute*.cpp +
    ArcSim's mesh code compiled unmodified (oracle/Makefile; the prebuilt file travels to the GPU box); its Forces::fill is what a
    step times.  The reference program is single-threaded (`single_thread_value`); to use every host core the arm runs one
    independent instance per core side by side — each its own Mesh / Forces objects on its own strip of the workload's sheet — and
    `value` is their combined rate: the most this host gets out of the reference's code, which makes it the conservative denominator.
    Step = one Forces::fill per instance on a BOUNDED SAMPLE: `--ref-rows` grid rows of the same 1024 x 1024 sheet (same coordinates,
    numbering, state) per instance, so that the K + W steps the driver asks for end within minutes.  Beside it: the threaded timing
    variant of the oracle port on 128 rows (`port_threaded`, round 1-2's denominator) and ONE fill of the full configuration with it
    (`full_config`; the reference's own code would need ~60 s and ~40 GB for that)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    n = {"sheet1024": 1024, "sheet256": 256, "ensemble64": 64}[args.workload]
    rows = n if n <= 256 else min(n, max(2, args.ref_rows))
    X, fn = O.sheet_regular2(n, rows=rows)
    es = O.arcsim_edge_stencils(X.shape[0], fn)
    x = O.drape_state(X, seed=0, n_total=n * n)
    elements = fn.shape[0] + int((es[:, 3] >= 0).sum())
    cores = max(1, min(len(os.sched_getaffinity(0)), 64))
    N_full, F_full, _, Ei_full, _, _ = sheet_counts(n)
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libforces_ref.so"))
    port = None
    if have_ref:
        dt1 = O.ref_forces_seconds(fn, x, X, MAT, GRAV, H, instances=1, repeats=2)          # the reference as it is: one thread
        times = O.ref_forces_seconds(fn, x, X, MAT, GRAV, H, instances=cores, repeats=args.warmup + args.steps, all_times=True)
        dt = sum(times[args.warmup:]) / args.steps
        value = cores * elements / dt
        kind = "reference"
        what = ("oracle/_ref/libforces_ref.so: the reference's own Forces.cpp + UtilEOL.cpp + conversions.cpp + Compute*.cpp + ArcSim mesh code, compiled "
                "unmodified against oracle/mini_eigen (Eigen is not in this image: its operators are restated there, bounds-checked); no product code is loaded by this arm")
        value_is = ("%d independent instances of the reference's single-threaded Forces::fill side by side, one per core, each on its own %d-row strip; "
                    "the reference program itself runs ONE (single_thread_value)" % (cores, rows))
        sample = (f"per instance: Forces::fill of {rows} grid rows of the regular2 n={n} sheet ({elements} of its {F_full + Ei_full} elements) per step; {cores} instances per step"
                  if rows < n else f"per instance: Forces::fill of the whole regular2 n={n} sheet ({elements} elements) per step; {cores} instances per step")
    prow = n if n <= 256 else min(n, 128)
    Xp, fnp = O.sheet_regular2(n, rows=prow)
    esp = O.arcsim_edge_stencils(Xp.shape[0], fnp)
    xp = O.drape_state(Xp, seed=0, n_total=n * n)
    elp = fnp.shape[0] + int((esp[:, 3] >= 0).sum())
    t = time.perf_counter()
    O.forces_fill(fnp, esp, xp, Xp, MAT, GRAV, H)
    dtp1 = time.perf_counter() - t
    psteps = args.steps if not have_ref else min(args.steps, 3)
    for _ in range(args.warmup if not have_ref else 1):
        O.forces_fill(fnp, esp, xp, Xp, MAT, GRAV, H, threads=cores)
    t = time.perf_counter()
    for _ in range(psteps):
        O.forces_fill(fnp, esp, xp, Xp, MAT, GRAV, H, threads=cores)
    dtp = (time.perf_counter() - t) / psteps
    port = {"value": elp / dtp, "single_thread_value": elp / dtp1, "unit": UNIT, "cores": cores, "steps": psteps,
            "sample": f"first {prow} grid rows of the sheet ({elp} elements) per step",
            "what": "oracle port (reference Compute*.cpp object code + Eigen-free restatement of the Forces.cpp glue, pinned to libforces_ref.so in "
                    "tests/test_forces_ref_pin.py), element loops on all cores, the two setFromTriplets side by side"}
    if not have_ref:
        dt1, dt, value, kind, elements = dtp1, dtp, elp / dtp, "port", elp
        what, value_is, sample = port["what"], "threaded timing variant of the port (oracle/_ref/libforces_ref.so is missing)", port["sample"]
    full = None
    avail = mem_available_gb()
    if rows < n and not args.no_full:
        if avail is not None and avail < 40.0:
            full = {"skipped": "MemAvailable %.1f GB < 40 GB (the reference path materialises ~23 GB of triplets at 1024^2)" % avail}
        else:
            Xf, fnf = O.sheet_regular2(n)
            esf = O.arcsim_edge_stencils(Xf.shape[0], fnf)
            xf = O.drape_state(Xf, seed=0)
            t = time.perf_counter()
            r = O.forces_fill(fnf, esf, xf, Xf, MAT, GRAV, H, threads=cores)
            dtf = time.perf_counter() - t
            full = {"value": (F_full + Ei_full) / dtf, "unit": UNIT, "seconds": dtf, "elements": F_full + Ei_full, "steps": 1, "kind": "port",
                    "nnz_MDK": int(r["MDK"][2].size), "seconds_elements_assembly": list(r["seconds"]),
                    "note": "one fill of the whole configuration by the threaded port, same cores"}
            del r
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.workload == "ensemble64" else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "value_is": value_is,
                             "single_thread_value": (elements if have_ref else elp) / dt1, "port_threaded": port, "full_config": full,
                             "mem_available_gb": avail, "note": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def workload_name(w):
    return {"sheet1024": "regular2 1024x1024 sheet, Forces::fill (f+M+MDK) into fixed CSR pattern (BASELINE configs[3])",
            "sheet256": "regular2 256x256 sheet, Forces::fill (BASELINE configs[1])",
            "ensemble64": "ensemble of 4096 regular2 64x64 scenes, Forces::fill batched (BASELINE configs[4])"}[w]


def numa_nodes():
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except Exception:
        return None


def bind_to_gpu_numa(local):
    """Keep this rank's host thread and its first-touch allocations (the pinned e2e buffers) on the CPUs next to its GPU: with one rank
    per GPU and unbound processes the 8 x 1.5 GB device->host copies of the e2e leg otherwise meet on one socket's memory.
    Returns a description for the JSON line (None when the topology cannot be read)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, devid)
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "cpus %s" % txt
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import eol_cloth_b200 as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    # stdout carries the ONE JSON line and nothing else: library chatter (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = E.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    if args.workload == "ensemble64":
        n, total_scenes = 64, 4096
        lo, hi = shard_range(total_scenes, rank, world)
        S = hi - lo
        scaling = "strong"
    else:
        n = 1024 if args.workload == "sheet1024" else 256
        S, lo = 1, rank
        scaling = "weak"
    X, fn, es, _ = make_sheet(n, 0)
    N, F = X.shape[0], fn.shape[0]
    plan = E.ForcesPlan(ctx, N, fn, es, X_hint=X)
    Ei = plan.n_interior_edges
    elements = F + Ei
    nnzM, nnzK = plan.nnz
    abytes = algorithmic_bytes(N, F, Ei, nnzM, nnzK)
    assert (N, F, Ei, nnzM, nnzK) == tuple(sheet_counts(n)[i] for i in (0, 1, 3, 4, 5)), "closed-form counts disagree with the plan"

    xs = np.stack([E.meshgen.drape_state(X, seed=lo + s) for s in range(S)]) if S <= 64 else None
    if xs is None:   # big ensembles: perturb per scene on the device-side copy (seeded, cheap)
        base = E.meshgen.drape_state(X, seed=0)
        rng = np.random.default_rng(lo)
        xs = base[None] + rng.uniform(-1e-3, 1e-3, size=(S,) + base.shape)
    x_d = torch.from_numpy(xs).to(dev)
    X_d = torch.from_numpy(np.broadcast_to(X, (S,) + X.shape).copy()).to(dev)
    f_d = torch.empty((S, 3 * N), dtype=torch.float64, device=dev)
    M_d = torch.empty((S, nnzM), dtype=torch.float64, device=dev)
    K_d = torch.empty((S, nnzK), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step_dev():
        plan.fill_dev(x_d.data_ptr(), X_d.data_ptr(), MAT, GRAV, H, f_d.data_ptr(), M_d.data_ptr(), K_d.data_ptr(), n_scenes=S)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(args.warmup):
        step_dev()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local)
    clk.__enter__()                      # samples clocks over BOTH timed regions (device-resident and end-to-end)
    time.sleep(0.05)
    ev0.record(stream)
    for _ in range(args.steps):
        step_dev()
    ev1.record(stream)
    barrier()
    ms_local = ev0.elapsed_time(ev1) / args.steps
    ms = max_over_ranks(ms_local, dev)
    total_elements = sum_over_ranks(float(elements * S), dev)
    value = total_elements / (ms * 1e-3)
    checksum = float(K_d[0].sum().item()) + float(f_d[0].sum().item())
    # the steady step of a simulation without remeshing (EOLC_FILL_M_UNCHANGED): M neither recomputed nor written (SURVEY §8d:
    # "membrane+bending-only variant if M is cached as constant": 8 nnz(M) fewer algorithmic bytes)
    ev0.record(stream)
    for _ in range(args.steps):
        plan.fill_dev(x_d.data_ptr(), X_d.data_ptr(), MAT, GRAV, H, f_d.data_ptr(), M_d.data_ptr(), K_d.data_ptr(), n_scenes=S, m_unchanged=True)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_mu = ev0.elapsed_time(ev1) / args.steps
    # the timed region above is short (K fills of < 1 ms): five more blocks of K fills each, for the spread (not the headline)
    repeat_ms = []
    for _ in range(5):
        ev0.record(stream)
        for _ in range(args.steps):
            step_dev()
        ev1.record(stream)
        torch.cuda.synchronize()
        repeat_ms.append(ev0.elapsed_time(ev1) / args.steps)

    # ---- e2e: host-buffer C-ABI entry (what eolc::host::Forces::fill calls), one scene per step per rank.  Buffers are page-locked
    # memory from eolc_host_alloc (the host layer's PinnedArray).  Steady step: EOLC_FILL_M_UNCHANGED (X and density as in the
    # previous fill => M is the same matrix and is neither recomputed nor copied); full step: everything.
    from eol_cloth_b200 import capi
    hb = [capi.HostBuffer(shape) for shape in ((N, 3), (N, 2), (3 * N,), (nnzM,), (nnzK,))]
    xa, Xa, fa, Ma, Ka = (b.array for b in hb)
    xa[:] = xs[0]; Xa[:] = X
    e2e_steps = max(3, min(args.steps, 10))

    def timed_host(fn_step, reps):
        for _ in range(2):
            fn_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn_step()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / reps * 1e3, dev)

    e2e_full_ms = timed_host(lambda: plan.fill_into(xa, Xa, MAT, GRAV, H, fa, Ma, Ka), e2e_steps)
    e2e_ms = timed_host(lambda: plan.fill_into(xa, Xa, MAT, GRAV, H, fa, Ma, Ka, m_unchanged=True), e2e_steps)
    e2e_checksum = float(Ka[::4099].sum() + fa[::257].sum() + Ma[::4099].sum())
    # the same call with pageable caller buffers (plain numpy arrays): staged through the library's pinned buffer + one more copy
    pg = [np.empty_like(a) for a in (fa, Ma, Ka)]
    e2e_pageable_ms = None
    if rank == 0:      # no collective in here: the other ranks have moved on
        xp_, Xp_ = np.ascontiguousarray(xs[0]), np.ascontiguousarray(X)
        for it in range(4):
            if it == 1:
                t0 = time.perf_counter()
            plan.fill_into(xp_, Xp_, MAT, GRAV, H, pg[0], pg[1], pg[2], m_unchanged=True)
        e2e_pageable_ms = (time.perf_counter() - t0) / 3 * 1e3
    del pg
    clk.__exit__(None, None, None)
    clocks = clk.summary()
    total_el_1 = sum_over_ranks(float(elements), dev)
    e2e_value = total_el_1 / (e2e_ms * 1e-3)
    h2d = 8 * 3 * N                  # steady step: X is the previous fill's (the flag's contract), only x goes up
    h2d_full = 8 * (3 * N + 2 * N)
    d2h = 8 * (3 * N + nnzK)
    d2h_full = 8 * (3 * N + nnzM + nnzK)
    for b_ in hb:
        b_.free()

    peak, peak_src = measured_peak()
    traffic, traffic_src = measured_traffic(args.workload)
    pipeline = os.environ.get("EOLC_FORCES_PIPELINE", "tiles")
    achieved = abytes * S / (ms_local * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "api": "eolc_forces_fill_ex(EOLC_FILL_M_UNCHANGED), page-locked host buffers from eolc_host_alloc: x in; f, MDK out; "
                       "X is the previous step's (the flag's contract) and stays on the device, M is the previous step's matrix "
                       "(depends on X and the density only) and is not copied again",
                "full_fill": {"value": total_el_1 / (e2e_full_ms * 1e-3), "ms_per_step": e2e_full_ms, "h2d_bytes_per_step": h2d_full,
                              "d2h_bytes_per_step": d2h_full,
                              "what": "every step recomputes and copies M too (a step after remeshing)"},
                "pageable_ms_per_step": e2e_pageable_ms, "checksum": e2e_checksum, "numa_nodes": numa_nodes(),
                "limiter": "PCIe: %.0f MB device-to-host per rank-step (%.1f GB/s here); at N > 1 the ranks' copies meet in the host's "
                           "memory system (the box exposes %s NUMA node(s)), so per-rank time grows with N" % (d2h / 1e6, d2h / e2e_ms / 1e6, numa_nodes())},
        "gpu_launches": args.steps * plan.launches_per_fill * (1 if S == 1 else 1),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic if pipeline == "tiles" else None, "traffic_source": traffic_src if pipeline == "tiles" else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_fill": abytes, "launches_per_fill": plan.launches_per_fill,
                     "kernel": "assemble_%s_kernel: one launch = one fill of all scenes of the rank (%d elements)" % (pipeline, elements * S),
                     "ms": ms_local, "repeat_ms": sorted(repeat_ms),
                     "m_unchanged": {"ms": ms_mu, "algorithmic_bytes_per_fill": abytes - 8 * nnzM,
                                     "achieved": (abytes - 8 * nnzM) * S / (ms_mu * 1e-3) / 1e9,
                                     "frac": (abytes - 8 * nnzM) * S / (ms_mu * 1e-3) / 1e9 / measured_peak()[0],
                                     "what": "EOLC_FILL_M_UNCHANGED: the same launch without the M rows (steady step between remeshes)"},
                     "note": "not HBM-bound: FP64 issue + shared-memory traffic bound, see DESIGN.md 3.4"},
        "checksum": checksum,
    }

    # the other roof of this kernel (the honest bound is max(HBM time, FP64 time)): FP64 warp instructions executed per
    # launch (a property of the plan, from the committed ncu capture) against the MEASURED issue rate of the FP64 pipe
    # (profiles/fp64_peak.json: scripts/micro/fp64_peak.cu, dependent-DFMA chains, 1.986 warp instructions per clock and SM)
    try:
        fp = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]["fp64_warp_instructions"]
        pk = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["peak"]
        n_instr = fp["DFMA"] + fp["DMUL"] + fp["DADD"]
        if pipeline == "tiles" and S == 1:
            line["roofline"]["fp64"] = {"warp_instructions_per_fill": n_instr,
                                        "tflops_executed": (2 * fp["DFMA"] + fp["DMUL"] + fp["DADD"]) * 32 / (ms_local * 1e-3) / 1e12,
                                        "peak_tflops_measured": pk["tflops"], "peak_warp_instr_per_s_measured": pk["warp_instr_per_s"],
                                        "frac_of_fp64_issue_peak": n_instr / (pk["warp_instr_per_s"] * ms_local * 1e-3),
                                        "floor_ms_at_measured_peak": n_instr / pk["warp_instr_per_s"] * 1e3,
                                        "source": fp["source"], "peak_source": "profiles/fp64_peak.json (DFMA microbenchmark on B200, this round)"}
    except Exception:
        pass
    if not args.no_ensemble and args.workload != "ensemble64":
        ens = bench_ensemble(ctx, dev, stream, rank, world, barrier)     # every rank takes part
        if rank == 0:
            line["ensemble"] = ens
    solo = world == 1      # the CPU baseline and the secondary objects are measured on the N = 1 line only
    if rank == 0 and solo and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg()
    if rank == 0 and solo and not args.no_cd:
        line["cd"] = bench_cd(ctx, dev, stream)
    if rank == 0 and solo and not args.no_cd and S == 1:
        line["consumer"] = bench_consumer(plan, dev, stream, f_d, M_d, K_d, N, nnzM, nnzK)
        line["consumer"]["normals"] = bench_normals(plan, dev, stream, x_d, N, F)
        line["eol"] = bench_eol(ctx, dev, stream, n, X, fn, es, x_d, X_d, ms_local)
    if rank == 0 and solo and S == 1 and not args.no_cpu:
        line["plan_build_ms"] = bench_plan_build(ctx)
        hl = bench_host_layer(n)
        if hl is not None:
            line["e2e"]["host_layer"] = hl
    if numa:
        line["e2e"]["host_affinity"] = numa
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def gather_ranks(v, device, world):
    import torch
    if world == 1:
        return [float(v)]
    t = torch.tensor([float(v)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    torch.distributed.all_gather(out, t)
    return [float(o.item()) for o in out]


def bench_ensemble(ctx, dev, stream, rank, world, barrier, total_scenes=4096, n=64, steps=10, warmup=3):
    """BASELINE configs[4] / SURVEY §8e: 4096 independent regular2 64x64 scenes (state seed = scene id) over the box of the box
    scene, sharded contiguously over the ranks (strong scaling, no collective on the data path).  A step = one batched Forces::fill
    of the rank's scenes (1 launch) + one batched CD2 narrow phase (records stay on the device, the per-scene offsets come
    back).  Timed with CUDA events on the library's stream, max over ranks."""
    import torch
    import eol_cloth_b200 as E
    from eol_cloth_b200.collisions import make_obstacles
    lo, hi = shard_range(total_scenes, rank, world)
    S = hi - lo
    X, fn = E.meshgen.regular2(n)
    N, F = X.shape[0], fn.shape[0]
    es = E.meshgen.edge_stencils(N, fn)
    centre = np.array([0.9175, -0.25, -0.549])           # SURVEY §8d variant 3b: all three contact types fire
    obs = make_obstacles(E.meshgen.BOX_THRESHOLD, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(centre)[None])
    plan = E.ForcesPlan(ctx, N, fn, es, X_hint=X)
    cdp = E.CollisionPlan(ctx, N, fn, E.meshgen.BOX_THRESHOLD)
    nnzM, nnzK = plan.nnz
    elements = F + plan.n_interior_edges
    xs = np.stack([E.meshgen.box_scene_state(X, seed=lo + s, centre=centre) for s in range(S)])
    x_d = torch.from_numpy(xs).to(dev)
    X_d = torch.from_numpy(np.broadcast_to(X, (S,) + X.shape).copy()).to(dev)
    f_d = torch.empty((S, 3 * N), dtype=torch.float64, device=dev)
    M_d = torch.empty((S, nnzM), dtype=torch.float64, device=dev)
    K_d = torch.empty((S, nnzK), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fill_ms = cd_ms = 0.0
    contacts = 0
    for it in range(warmup + steps):
        if it == warmup:
            barrier()
        ev[0].record(stream)
        plan.fill_dev(x_d.data_ptr(), X_d.data_ptr(), MAT, GRAV, H, f_d.data_ptr(), M_d.data_ptr(), K_d.data_ptr(), n_scenes=S)
        ev[1].record(stream)
        off = cdp.run_resident(x_d.data_ptr(), obs, 0, 0, n_scenes=S)
        ev[2].record(stream)
        torch.cuda.synchronize()
        if it >= warmup:
            fill_ms += ev[0].elapsed_time(ev[1]); cd_ms += ev[1].elapsed_time(ev[2])
        contacts = int(off[-1])
    barrier()
    fill_ms /= steps; cd_ms /= steps
    pair_tests, cd_launches = cdp.stats()
    per_rank = gather_ranks(fill_ms + cd_ms, dev, world)
    step_ms = max(per_rank)
    fill_max, cd_max = max_over_ranks(fill_ms, dev), max_over_ranks(cd_ms, dev)
    tot_contacts = sum_over_ranks(float(contacts), dev)
    tot_tests = sum_over_ranks(float(pair_tests), dev)
    N_, F_, _, Ei_, nM_, nK_ = sheet_counts(n)
    abytes = algorithmic_bytes(N_, F_, Ei_, nM_, nK_) * S
    peak, _ = measured_peak()
    cd_bytes = 24.0 * N * S + 264.0 * contacts       # this rank
    checksum = float(K_d[0].sum().item()) + float(off[1])
    out = {"workload": "ensemble of %d independent regular2 %dx%d scenes (BASELINE configs[4]): batched Forces::fill + batched CD2 over the box" % (total_scenes, n, n),
           "scaling": "strong", "scenes_total": total_scenes, "scenes_this_rank": S, "shard": "contiguous, rank r owns [r S/G, (r+1) S/G)",
           "steps": steps, "warmup": warmup,
           "elements_per_s": total_scenes * elements / (fill_max * 1e-3), "contacts_per_s": tot_contacts / (cd_max * 1e-3),
           "pair_tests_per_s": tot_tests / (cd_max * 1e-3), "scenes_per_s": total_scenes / (step_ms * 1e-3),
           "ms_per_step": step_ms, "fill_ms": fill_max, "cd_ms": cd_max, "per_rank_ms": per_rank,
           "load_balance": (sum(per_rank) / len(per_rank)) / step_ms,
           "contacts_total": int(tot_contacts), "launches_per_step": plan.launches_per_fill + cd_launches,
           "fill_roofline": {"achieved_GBps": abytes / (fill_ms * 1e-3) / 1e9, "frac_of_hbm_peak": abytes / (fill_ms * 1e-3) / 1e9 / peak},
           "cd_roofline": {"algorithmic_bytes": cd_bytes, "achieved_GBps": cd_bytes / (cd_ms * 1e-3) / 1e9,
                           "frac_of_hbm_peak": cd_bytes / (cd_ms * 1e-3) / 1e9 / peak,
                           "what": "positions in (24 B per node and scene) + records out (264 B per contact); not HBM-bound: its kernels are "
                                   "chains of dependent gathers and FP64 tests (DESIGN.md 4)"},
           "timing": "CUDA events on the library stream, max over ranks; contacts stay on the device (eolc_cd_run_batched_resident_dev)",
           "parity": "tests/test_cd_gpu.py::test_ensemble_4096_scenes_sampled_against_reference, tests/test_forces_gpu.py (batched fill)",
           "checksum": checksum}
    plan.close(); cdp.close()
    del x_d, X_d, f_d, M_d, K_d
    torch.cuda.empty_cache()
    return out


def bench_plan_build(ctx):
    """Host cost of a topology change: eolc_forces_plan_create (pattern, tiles, templates, upload) for the 512^2
    sheet in its structured numbering, with a random node / face numbering (every tile its own template), and for the 1024^2 sheet."""
    import eol_cloth_b200 as E
    out = {}
    for label, n, shuffle in (("sheet512", 512, False), ("sheet512_shuffled", 512, True), ("sheet1024", 1024, False)):
        X, fn = E.meshgen.regular2(n)
        if shuffle:
            rng = np.random.default_rng(5)
            perm = rng.permutation(X.shape[0])
            Xn = np.empty_like(X); Xn[perm] = X
            X, fn = Xn, perm[fn].astype(np.int32)[rng.permutation(len(fn))]
        es = E.meshgen.edge_stencils(X.shape[0], fn)
        best = None
        for _ in range(2):
            t = time.perf_counter()
            plan = E.ForcesPlan(ctx, X.shape[0], fn, es, X_hint=X)
            dt = time.perf_counter() - t
            plan.close()
            best = dt if best is None else min(best, dt)
        out[label] = best * 1e3
    out["what"] = ("ms of host time for eolc_forces_plan_create (best of 2: the second build finds the bank-optimised templates in the "
                   "process-wide cache, as a re-plan in a running simulation does), threads = min(cores, 16)")
    return out


def bench_host_layer(n):
    """Adapter-level step time: the C++ host layer (include/eolc_host.hpp) driven from a C++ program with an ArcSim-shaped pointer
    mesh — flatten() + Forces::fill, as adapter/Forces_fill_b200.cpp does each step (minus the copy into Eigen objects)."""
    exe = os.path.join(ROOT, "tests", "cpp", "host_driver")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, "bench", str(n), "3"], capture_output=True, text=True, timeout=300)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        d["what"] = "flatten of the pointer mesh + eolc::host::Forces::fill per step (tests/cpp/host_driver bench)"
        d["elements_per_s_steady"] = d["elements"] / (d["step_M_unchanged_ms"] * 1e-3)
        return d
    except Exception as e:
        return {"error": str(e)[:200]}


def bench_consumer(plan, dev, stream, f_d, M_d, K_d, N, nnzM, nnzK):
    """Secondary: the consumer of the fill on the device: b = -(M v + h f) and one CG iteration on MDK.
    HBM-bound kernels; bytes = matrix values + 4 B of column-block index per 3x3 block + the vectors each kernel touches."""
    import torch
    v = torch.zeros(3 * N, dtype=torch.float64, device=dev)
    b = torch.empty_like(v)
    sol = torch.empty_like(v)
    torch.cuda.synchronize()
    for _ in range(3):
        plan.rhs_dev(M_d.data_ptr(), f_d.data_ptr(), v.data_ptr(), H, b.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        plan.rhs_dev(M_d.data_ptr(), f_d.data_ptr(), v.data_ptr(), H, b.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    rhs_ms = e0.elapsed_time(e1) / 10
    rhs_bytes = 8 * nnzM + 4 * nnzM // 9 + 8 * 3 * N * 3
    plan.solve_cg_dev(K_d.data_ptr(), b.data_ptr(), sol.data_ptr(), tol=1e-300, max_iter=8)
    torch.cuda.synchronize()
    e0.record(stream)
    it, _ = plan.solve_cg_dev(K_d.data_ptr(), b.data_ptr(), sol.data_ptr(), tol=1e-300, max_iter=32)
    e1.record(stream)
    torch.cuda.synchronize()
    cg_ms = e0.elapsed_time(e1) / max(it, 1)
    cg_bytes = 8 * nnzK + 4 * nnzK // 9 + 8 * 3 * N * 12
    peak, _ = measured_peak()
    return {"rhs": {"what": "b = -(M v + h f), Cloth.cpp:345", "ms": rhs_ms, "GB_per_s": rhs_bytes / rhs_ms / 1e6, "frac_of_hbm_peak": rhs_bytes / rhs_ms / 1e6 / peak},
            "cg_iteration": {"what": "Jacobi-preconditioned CG on MDK (GeneralizedSolver.cpp:120-126), 5 launches, incl. the host's convergence check every 8",
                             "ms": cg_ms, "GB_per_s": cg_bytes / cg_ms / 1e6, "frac_of_hbm_peak": cg_bytes / cg_ms / 1e6 / peak}}


def bench_normals(plan, dev, stream, x_d, N, F):
    """Secondary: face + node normals of compute_ws_data on the device, both kernels per call."""
    import torch
    fn_d = torch.empty(3 * F, dtype=torch.float64, device=dev)
    nn_d = torch.empty(3 * N, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        plan.normals_dev(x_d.data_ptr(), fn_d.data_ptr(), nn_d.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        plan.normals_dev(x_d.data_ptr(), fn_d.data_ptr(), nn_d.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = 2 * (24 * N + 12 * F) + 4 * (N + 1) + 4 * 3 * F + 24 * F + 24 * N      # x + face idx (twice), node->face CSR, both outputs
    peak, _ = measured_peak()
    return {"what": "face->n and node->n (ArcSim mesh.cpp:135-143, geometry.cpp:302-316), 2 launches", "ms": ms,
            "GB_per_s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}


def bench_eol(ctx, dev, stream, n, X, fn, es, x_d, X_d, lagrangian_ms):
    """Secondary: the same sheet with the n - 2 interior nodes of its middle grid line flagged EoL (the cloth crossing a box edge):
    Forces::fill through the EOL branch, 3 launches per fill, against the Lagrangian fill of the same state."""
    import torch
    import eol_cloth_b200 as E
    N = X.shape[0]
    eol = np.full(N, -1, np.int32)
    line = np.arange(1, n - 1) * n + n // 2
    eol[line] = np.arange(line.size)
    plan = E.ForcesPlan(ctx, N, fn, es, eol_index=eol, X_hint=X)
    f_d = torch.empty(plan.dof, dtype=torch.float64, device=dev)
    M_d = torch.empty(plan.nnz[0], dtype=torch.float64, device=dev)
    K_d = torch.empty(plan.nnz[1], dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        plan.fill_dev(x_d.data_ptr(), X_d.data_ptr(), MAT, GRAV, H, f_d.data_ptr(), M_d.data_ptr(), K_d.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        plan.fill_dev(x_d.data_ptr(), X_d.data_ptr(), MAT, GRAV, H, f_d.data_ptr(), M_d.data_ptr(), K_d.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out = {"workload": f"regular2 n={n} sheet, {line.size} EoL nodes on the grid line j = n/2", "dof": plan.dof, "nnz_M": plan.nnz[0],
           "nnz_MDK": plan.nnz[1], "launches_per_fill": plan.launches_per_fill, "ms_per_fill": ms, "lagrangian_ms_per_fill": lagrangian_ms,
           "checksum_f_eulerian": float(f_d[3 * N:].sum().item())}
    plan.close()
    return out


def bench_cd(ctx, dev, stream):
    """Secondary metric: CD2 narrow phase on the 512x512 box scene (BASELINE configs[2])."""
    import torch
    import eol_cloth_b200 as E
    from eol_cloth_b200.collisions import make_obstacles
    X, fn = E.meshgen.regular2(512)
    x = E.meshgen.box_scene_state(X, seed=0)
    obs = make_obstacles(E.meshgen.BOX_THRESHOLD, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame()[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, E.meshgen.BOX_THRESHOLD)
    x_d = torch.from_numpy(x).to(dev)
    torch.cuda.synchronize()
    cap = X.shape[0] + 4096
    # caller-owned, page-locked record buffer reused across calls (what a C++ host would do; a fresh pageable numpy array per call
    # costs more in page faults than the whole narrow phase)
    rec_bytes = E.CONTACT_DTYPE.itemsize
    pinned = torch.empty(cap * rec_bytes, dtype=torch.uint8).pin_memory()
    out = pinned.numpy().view(E.CONTACT_DTYPE)
    for _ in range(3):
        c, _ = plan.run(x_d.data_ptr(), obs, 0, 0, x_is_device_ptr=True, out=out)
    reps = 20
    t = time.perf_counter()
    for _ in range(reps):
        c, _ = plan.run(x_d.data_ptr(), obs, 0, 0, x_is_device_ptr=True, out=out)
    dt = (time.perf_counter() - t) / reps
    pair_tests, launches = plan.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        plan.run_resident(x_d.data_ptr(), obs, 0, 0)
    t = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        plan.run_resident(x_d.data_ptr(), obs, 0, 0)
    e1.record(stream)
    torch.cuda.synchronize()
    dt_res = (time.perf_counter() - t) / reps
    dev_ms = e0.elapsed_time(e1) / reps
    for _ in range(2):
        plan.run_resident(x_d.data_ptr(), obs, 0, 0); rows = plan.contact_rows(); csr = plan.contact_rows_csr()
    t = time.perf_counter()
    for _ in range(reps):
        plan.run_resident(x_d.data_ptr(), obs, 0, 0)
        rows = plan.contact_rows()
    dt_rows = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps):
        plan.run_resident(x_d.data_ptr(), obs, 0, 0)
        csr = plan.contact_rows_csr()
    dt_csr = (time.perf_counter() - t) / reps
    out = {"workload": "CD2, regular2 512x512 over the simulationSettingsBox.json box (BASELINE configs[2])",
           "resident": {"ms_per_call": dt_res * 1e3, "stream_ms_per_call": dev_ms, "contacts_per_s": len(c) / dt_res,
                        "what": "eolc_cd_run_batched_resident_dev: records stay on the device, offsets come back (one stream sync)"},
           "resident_plus_rows": {"ms_per_call": dt_rows * 1e3, "rows": int(len(rows[0])), "d2h_bytes": int(112 * len(rows[0])),
                                  "what": "resident run + eolc_cd_contact_rows: the fixed-width 112 B/contact inequality rows of Constraints::fill to page-locked host arrays instead of the 264 B records"},
           "resident_plus_rows_csr": {"ms_per_call": dt_csr * 1e3, "rows": int(len(csr[0]) - 1), "nnz": int(len(csr[1])),
                                      "d2h_bytes": int(4 * len(csr[0]) + 12 * len(csr[1])), "contacts_per_s": len(c) / dt_csr,
                                      "what": "resident run + eolc_cd_contact_rows_csr: the same rows compacted on the device (row_ptr / cols / vals = Aineq's triplet sequence)"},
           "contacts": int(len(c)), "ms_per_call": dt * 1e3, "contacts_per_s": len(c) / dt,
           "pair_tests_per_s": pair_tests / dt, "launches_per_call": launches,
           "timing": "wall clock around eolc_cd_run_dev incl. D2H of the contact list (%d B records, pinned caller buffer) and the host post-pass" % rec_bytes}
    try:
        from oracle import oracle as O
        t = time.perf_counter()
        ref = O.cd(fn, x, E.meshgen.BOX_THRESHOLD, None, None, obs.box_whd, obs.box_E, 0, 0)
        dtc = time.perf_counter() - t
        out["cpu_baseline"] = {"contacts_per_s": len(ref) / dtc, "ms_per_call": dtc * 1e3, "cores": 1, "kind": "port"}
    except Exception as e:
        out["cpu_baseline"] = {"error": str(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sheet1024", choices=["sheet1024", "sheet256", "ensemble64"])
    ap.add_argument("--ref-rows", type=int, default=16, help="--impl reference: grid rows of the sheet per instance and step (one instance per core)")
    ap.add_argument("--no-full", action="store_true", help="--impl reference: skip the single fill of the full configuration")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the configs[4] ensemble object")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cd", action="store_true", help="skip the secondary CD measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
