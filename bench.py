#!/usr/bin/env python
"""Benchmark of the b200 hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size M]

A *step* is one pass of the hot path = one Newton step of the 3-D Q1-hex Poisson problem
(BASELINE.json configs[3]: M^3 hex8, default M=256 -> 16 974 593 dofs): Dirichlet imposition,
fused tangent+residual assembly, Jacobi-PCG to rtol 1e-8, update, residual assembly + norm.
  value      elements / device time of the step (apdx_newton, inputs resident in HBM, CUDA events)
  e2e        the same through the public API autopdex_b200.solver.solver(dofs, settings,
             static_settings) with HOST NumPy buffers (H2D of coordinates / dofs / Dirichlet
             values and D2H of the solution inside the timed region)
  roofline   SpMV kernel: algorithmic bytes nnz*12 + n*16 + (n+1)*4 / CUDA-event time
  cpu_baseline / --impl reference: the reference's CPU path (oracle assembly restating the JAX half +
             the reference's own SciPy calls) on a bounded sample mesh, host cores of this box.
N > 1 (torchrun): strong scaling, the M^3 mesh is split into slabs along i, one rank per GPU,
NCCL halo exchange + all-reduce; timing is the max over ranks.
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

UNIT_CUBE = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]
# dram__bytes_read.sum + dram__bytes_write.sum of one k_spmv_sell<1> launch from `ncu --set full` (profiles/), by mesh size
NCU_TRAFFIC = {256: 2.2265e9}     # 2.0938 GB read + 0.1327 GB written
NCU_TRAFFIC_SOURCE = "profiles/r02a_k_spmv_sell_p256_full.txt"
# SASS count of k_elem_scalar_reg<3,8,8,true>: 2592 DFMA + 368 DMUL + 316 DADD per element
HEX8_POISSON_FLOP = 2 * 2592 + 368 + 316
METRIC = "elements/s through one Newton step (sparse assembly + Jacobi-PCG to 1e-8)"
UNIT = "elements/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- workload ------------------------------------------------------------------------------------------
NX_OVERRIDE = [0]   # diagnostic: elements along i (slowest index) if different from m


def local_poisson_mesh(m, rank, nranks):
    """Slab [rank] of the m^3 brick mesh of the unit cube, nodes in global order (local id = global - node_lo)."""
    from autopdex_b200 import mesher
    mx = NX_OVERRIDE[0] or m
    part = mesher.slab_partition((mx, m, m), rank, nranks)
    g0, g1 = part["plane_lo"], part["plane_hi"]
    lin = np.linspace(-1.0, 1.0, m + 1)
    lin_x = np.linspace(-1.0, 1.0, mx + 1)
    v = np.asarray(UNIT_CUBE)
    S, T, U = np.meshgrid(lin_x[g0:g1], lin, lin, indexing="ij")
    s, t, u = S.reshape(-1, 1), T.reshape(-1, 1), U.reshape(-1, 1)
    coords = ((1 - s) * (1 - t) * (1 - u) * v[0] + (1 + s) * (1 - t) * (1 - u) * v[1] + (1 + s) * (1 + t) * (1 - u) * v[2]
              + (1 - s) * (1 + t) * (1 - u) * v[3] + (1 - s) * (1 - t) * (1 + u) * v[4] + (1 + s) * (1 - t) * (1 + u) * v[5]
              + (1 + s) * (1 + t) * (1 + u) * v[6] + (1 - s) * (1 + t) * (1 + u) * v[7]) / 8
    del S, T, U, s, t, u
    e0, e1 = part["elem_lo"], part["elem_hi"]
    I, J, K = [a.ravel() for a in np.meshgrid(np.arange(e0, e1), np.arange(m), np.arange(m), indexing="ij")]
    sy, sx = m + 1, (m + 1) * (m + 1)
    n0 = (I - g0) * sx + J * sy + K
    elems = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + 1, n0 + sx + 1, n0 + sx + sy + 1, n0 + sy + 1],
                     axis=1).astype(np.int32)
    tol = 1e-12
    mask = np.zeros(coords.shape[0], dtype=bool)
    for d in range(3):
        mask |= (np.abs(coords[:, d]) < tol) | (np.abs(coords[:, d] - 1.0) < tol)
    owned_elems = (min(part["owned_plane_hi"], mx) - part["owned_plane_lo"]) * m * m  # elements attributed to this rank
    return part, coords, elems, mask, owned_elems


def build_problem(m, rank, nranks):
    from autopdex_b200 import models, seeder, spaces
    part, coords, elems, mask, owned_elems = local_poisson_mesh(m, rank, nranks)
    weak = models.poisson_weak(lambda x, settings: 1.0, lambda x: 1.0)
    elem = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick,
                                                        *seeder.gauss_legendre_nd(dimension=3, order=2))
    static_settings = {"assembling mode": ("user element",), "solution structure": ("nodal imposition",),
                       "model": (elem,), "solver type": "newton", "solver backend": "b200", "solver": "cg",
                       "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"connectivity": (elems,), "node coordinates": coords, "dirichlet dofs": mask[:, None],
                "dirichlet conditions": np.zeros((coords.shape[0], 1))}
    if nranks > 1:
        settings["b200 partition"] = dict(owned_node_begin=part["owned_node_lo"] - part["node_lo"],
                                          owned_node_end=part["owned_node_hi"] - part["node_lo"],
                                          rank_lo=part["rank_lo"], rank_hi=part["rank_hi"],
                                          planes=(part["plane_lo"], part["plane_hi"], part["owned_plane_lo"],
                                                  part["owned_plane_hi"]))
    return settings, static_settings, owned_elems


# ---- CPU reference path (oracle + the reference's SciPy calls) --------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_step(m, solver="cg", rtol=1e-8, threads=None):
    """One Newton step of the same problem on an m^3 sample, reference loop semantics (solver.py:872-937) on the CPU:
    B-asm (NumPy restatement of the JAX assembly, element chunks on `threads` threads), B-dup (the reference's own SciPy
    calls), then B-cg = SciPy cg with the Jacobi preconditioner to the GPU arm's tolerance (solver='cg', the CPU analogue
    of solver.linear_solve_jax, solver.py:1093-1126) or B-direct = spsolve (solver='lapack', solver.py:1520), the update
    and the second residual assembly + norm.  Returns (seconds, elements, breakdown dict)."""
    from oracle import assemble as oasm
    from oracle import solve as osolve
    from tests import problems
    threads = threads or host_threads()
    p = problems.poisson_hex(m)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    prob.threads = threads
    rows, cols = prob.coo()                                                     # index arrays: built once per mesh
    free = ~p["mask"].ravel()
    dofs = np.zeros(p["mask"].shape)
    t0, c0 = time.perf_counter(), time.process_time()
    dofs[p["mask"]] = p["values"][p["mask"]]
    R, data = oasm.assemble(p["sets"], p["coords"], dofs, {}, threads=threads)  # B-asm (restated JAX half)
    t1 = time.perf_counter()
    csr = oasm.scipy_assembling(data, rows, cols, dofs.size, free)             # B-dup (reference's own SciPy calls)
    t2 = time.perf_counter()
    import scipy.sparse.linalg as spla
    b = -R[free]
    iters = [0]
    if solver == "lapack":
        x = spla.spsolve(csr, b)                                               # B-direct (SuperLU, solver.py:1520)
    else:
        d = csr.diagonal()

        def count(_):
            iters[0] += 1
        x, _ = spla.cg(csr, b, M=spla.LinearOperator(csr.shape, matvec=lambda v: v / d), rtol=rtol, atol=0.0,
                       callback=count)                                         # B-cg
    t3 = time.perf_counter()
    dofs.ravel()[free] += x
    R2 = prob.residual(dofs)                                                   # 2nd residual assembly of the step
    rn = float(np.linalg.norm(R2[free]))
    t4 = time.perf_counter()
    n_el = p["sets"][0]["conn"].shape[0]
    busy = (time.process_time() - c0) / max(t4 - t0, 1e-9)   # threads busy on average: process CPU time / wall time
    return t4 - t0, n_el, {"assembly_s": t1 - t0, "dup_sum_s": t2 - t1, "solve_s": t3 - t2, "residual_s": t4 - t3,
                           "krylov_iterations": iters[0], "res_norm": rn, "nnz_reduced": int(csr.nnz),
                           "threads_busy": busy}


REF_SIZES = (128, 112, 96, 80, 64, 48, 32)


def pick_ref_size(n_steps, budget_s, cap, threads):
    """Largest sample size whose n_steps Newton steps fit budget_s, predicted from one calibration step at 32^3:
    assembly, duplicate summing and the residual scale with the elements, Jacobi-PCG with elements^(4/3)."""
    dt, n_el, br = cpu_reference_step(32, "cg", threads=threads)
    lin = br["assembly_s"] + br["dup_sum_s"] + br["residual_s"]
    for m in REF_SIZES:
        if m > cap:
            continue
        f = (m / 32.0) ** 3
        est = lin * f + br["solve_s"] * f ** (4.0 / 3.0)
        if est * n_steps <= budget_s or m == REF_SIZES[-1]:
            return m, est
    return 32, dt


def cpu_sample_note(m, size, n_el):
    return ("%d^3 hex8 Poisson sample (%d elements) of the %d^3 workload, one Newton step with the GPU arm's solver: NumPy "
            "restatement of the JAX assembly (oracle/, element chunks on all host threads) + the reference's scipy "
            "coo->csr->[:,free][free] + SciPy cg with Jacobi M to rtol 1e-8 + update + residual; the full 256^3 step does "
            "not fit the host (25.8 GB of COO before SciPy's copies, BASELINE.md section 4) nor the time bound"
            % (m, n_el, size))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n_steps = args.warmup + args.steps
    m, est = pick_ref_size(n_steps, args.ref_budget_s, args.ref_size, threads)
    times, br = [], None
    for i in range(n_steps):
        dt, n_el, br_i = cpu_reference_step(m, "cg", args.rtol, threads)
        if i >= args.warmup:
            times.append(dt)
            br = br_i if br is None or dt <= min(times) else br
    dt = float(np.mean(times))
    value = n_el / dt
    direct = None
    if args.ref_direct:    # B-direct at 32^3 (3-D SuperLU fill-in: minutes and tens of GB beyond 64^3)
        d_dt, d_el, d_br = cpu_reference_step(32, "lapack", threads=threads)
        direct = {"size": 32, "value": d_el / d_dt, "unit": UNIT, "breakdown_s": d_br,
                  "note": "the reference's default 'scipy'/'lapack' backend (spsolve) on a 32^3 sample"}
    cores = max(1, int(round(br["threads_busy"])))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "best_ms_per_step": min(times) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.size, args.rtol, 1),
                       "sample": "CPU arm times a %d^3 sample of it per step (largest of %s whose %d steps fit %d s on %d host "
                                 "threads); same element, same solver (Jacobi-PCG, rtol %.0e), same Newton-step semantics; the "
                                 "ratio of the two arms is therefore per element, not per 256^3 step (PCG iterations grow "
                                 "with the mesh: %d here, 350 at 256^3)"
                                 % (m, "/".join(str(v) for v in REF_SIZES), n_steps, args.ref_budget_s, threads, args.rtol,
                                    br["krylov_iterations"])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_note(m, args.size, n_el),
                             "breakdown_s": br, "host_cores": os.cpu_count(), "threads_available": threads,
                             "threads_used": "assembly: element chunks on %d threads; scipy coo->csr, cg (SpMV) are serial, as "
                                             "in the reference's 'scipy' backend; cores = process CPU time / wall time" % threads,
                             "direct_solve_variant": direct},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_name(m, rtol, world):
    return ("3D Poisson Q1 hex %d^3 (%d dofs), Newton + Jacobi-PCG rtol %.0e, slab-partitioned over %d GPU(s)"
            % (m, (m + 1) ** 3, rtol, world))


# ncu (profiles/r02a_k_elements_neohooke64_full.txt): the k_elements<3,3> of that capture executed 2436 DFMA + 1386 DMUL +
# 791 DADD thread instructions per elapsed cycle over 2.144 M cycles for 262 144 hex8 neo-Hooke elements = 57.7 kflop per
# element, ~51.4 k of them in the node-pair loop (64 pairs x 8 Gauss points x ~100 flop).  The kernel now integrates the
# 36 upper node pairs only (symmetric tangent, stored twice): 6.3 k + 51.4 k * 36 / 64 = 35.2 kflop per element -- DERIVED
# from the capture by the trip counts, not re-measured (no GPU time was left).
HEX8_NEOHOOKE_FLOP_EXECUTED = 35200.0


def vector_newton_figures(n):
    """Config-5 family through the public API at n^3 (neo-Hooke brick clamped at x = 0, traction face at x = 1, nf = 3):
    one Newton solve with the config's Jacobi-BiCGSTAB and one with the multigrid-preconditioned CG (SURVEY.md 8f row N4:
    the reference's pyamg / PETSc preconditioners) -- Newton counts, Krylov iterations, device time, distance of the two
    solutions."""
    from autopdex_b200 import solver
    from tests import problems
    from tests.multi_gpu_worker import neohooke_api_problem
    p = problems.neo_hooke_brick(n)
    base = neohooke_api_problem(p)
    out, sols, kept = {}, {}, set(solver._PLAN_CACHE)      # the headline plan (bench still reads it) stays
    for name, st in (("jacobi_bicgstab", dict(base, solver="bicgstab")),
                     ("multigrid_pcg", dict(base, solver="cg", **{"type of preconditioner": "multigrid"}))):
        settings = {"connectivity": tuple(s["conn"].astype(np.int32) for s in p["sets"]), "node coordinates": p["coords"],
                    "dirichlet dofs": p["mask"], "dirichlet conditions": p["values"]}
        if name == "multigrid_pcg":
            settings["b200 multigrid"] = {"n_elements": (n, n, n)}
        d0 = np.zeros(p["mask"].shape)
        solver.solver(d0, settings, st, tol=1e-8)                     # plan (and hierarchy) build
        sol, info = solver.solver(d0, settings, st, tol=1e-8)
        q = dict(solver.last_stats)
        sols[name] = np.asarray(sol)
        out[name] = {"newton_steps": int(info[0]), "res_norm": float(info[1]), "diverged": bool(info[2]),
                     "krylov_iterations": int(q["krylov_iters"]), "device_ms": float(q["total_ms"]), "krylov_ms": float(q["krylov_ms"]),
                     "assembly_tangent_ms": float(q["assembly_tangent_ms"])}
        for key in [k for k in solver._PLAN_CACHE if k not in kept]:
            solver._PLAN_CACHE.pop(key).destroy()
    a, b = sols["multigrid_pcg"].ravel(), sols["jacobi_bicgstab"].ravel()
    out["rel_l2_between_the_two_solutions"] = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    out["workload"] = "3D neo-Hooke Q1 hex %d^3 brick (%d dofs), clamped face + traction face, Newton to 1e-8, Krylov rtol 1e-8" % (n, p["mask"].size)
    return out


def vector_problem_figures(args, hbm, fp64_peak, n=64):
    """Bounded nf = 3 run (BASELINE config 5 family at n^3): hex8 neo-Hooke tangent pass (generic element kernel
    k_elements<3,3> + deterministic scatter) and the nf = 3 sliced-ELL SpMV, both through the C ABI."""
    from autopdex_b200 import backend, mesher, seeder
    coords, elems = mesher.structured_mesh((n, n, n), UNIT_CUBE, "brick")
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-12)[:, None], 3, axis=1)
    st = backend.SetSpec("domain", "neo_hooke", elems.astype(np.int32), family="quad_brick", gp=seeder.gauss_legendre_nd(3, 2),
                         mode="3d", params={"youngs_modulus": 100.0, "poisson_ratio": 0.3})
    plan = backend.Plan(3, coords.shape[0], 3, [st], mask)
    plan.set_coords(coords)
    rng = np.random.default_rng(0)
    d = backend.DeviceArray.from_host(rng.uniform(-1e-3, 1e-3, mask.size))
    r = backend.DeviceArray(mask.size)
    tan = []
    for i in range(5):
        plan.assemble(d, True, r)
        tan.append(plan.stats()["assembly_tangent_ms"])
    tan_ms = float(np.mean(tan[2:]))
    spmv_ms = plan.time_spmv(50)
    rows, nnz = plan.f1 - plan.f0, plan.nnz_reduced
    phys = plan.stats()["sell_bytes"] + rows * 24
    n_el = elems.shape[0]
    tfl = n_el * HEX8_NEOHOOKE_FLOP_EXECUTED / (tan_ms * 1e-3) / 1e12
    out = {"workload": "3D neo-Hooke Q1 hex %d^3 (%d dofs), bounded sample of BASELINE config 5" % (n, mask.size),
           "assembly": {"tangent_pass_ms": tan_ms, "elements_per_s": n_el / (tan_ms * 1e-3),
                        "flop_per_element_executed": HEX8_NEOHOOKE_FLOP_EXECUTED, "tflops_fp64": tfl,
                        "frac_of_fp64_peak": tfl / fp64_peak,
                        "note": "whole apdx_assemble pass (element kernel + scatter into full CSR and sliced-ELL + residual); flops "
                                "per element derived from ncu's sass_thread_inst_executed_op_d{fma,mul,add} of the 64-pair "
                                "kernel (57.7 k) scaled to the 36 node pairs the kernel integrates now"},
           "spmv": {"ms_per_launch": spmv_ms, "bytes_per_launch": phys, "achieved_gbs": phys / (spmv_ms * 1e-3) / 1e9,
                    "frac": phys / (spmv_ms * 1e-3) / 1e9 / hbm, "nnz_reduced": nnz, "n_free": plan.n_free,
                    "algorithmic_gbs": (nnz * 12 + rows * 16 + (rows + 1) * 4) / (spmv_ms * 1e-3) / 1e9}}
    plan.destroy()
    out["newton"] = optional_section(vector_newton_figures, n)
    return out


def multigrid_figures(args, sol_jacobi, rank=0, world=1):
    """The same Newton step with 'type of preconditioner': 'multigrid' (opt-in: the headline above
    is BASELINE's Jacobi-PCG): geometric V-cycle on the structured hierarchy M^3 -> (M/2)^3 -> ..., Chebyshev smoothing,
    re-discretised coarse operators.  Device-resident step time, end-to-end time through solver.solver, and the distance
    of its solution from the Jacobi-PCG one.  On several GPUs the hierarchy is partitioned like the finest level (slabs;
    halo exchange per level and transfer); times are the maximum over the ranks."""
    from autopdex_b200 import backend, solver
    solver.clear_plan_cache()                     # the Jacobi plan's buffers go first
    m = args.size
    settings, static_settings, _ = build_problem(m, rank, world)
    static_settings = dict(static_settings, **{"type of preconditioner": "multigrid"})
    settings["b200 multigrid"] = {"n_elements": (NX_OVERRIDE[0] or m, m, m)}
    gmax = lambda v: float(backend.comm_allreduce_host([float(v)], "max")[0])
    dofs0 = np.zeros((settings["node coordinates"].shape[0], 1))
    t = time.perf_counter()
    sol, info = solver.solver(dofs0, settings, static_settings, tol=args.rtol)
    t_build = time.perf_counter() - t
    for _ in range(2):
        sol, info = solver.solver(dofs0, settings, static_settings, tol=args.rtol)
    t = time.perf_counter()
    for _ in range(args.steps):
        sol, info = solver.solver(dofs0, settings, static_settings, tol=args.rtol)
    e2e_ms = gmax((time.perf_counter() - t) / args.steps * 1e3)
    st = dict(solver.last_stats)
    state = next(iter(solver._PLAN_CACHE.values()))
    opts = backend.KrylovOptions("cg", rtol=args.rtol, jacobi="multigrid")
    step, kry, its = [], [], []
    for i in range(args.warmup + args.steps):
        state.dofs_d.zero()
        state.plan.newton(opts, state.dofs_d, state.vals_d, 1e-8, 30, 1.0)
        q = state.plan.stats()
        if i >= args.warmup:
            step.append(q["total_ms"]); kry.append(q["krylov_ms"]); its.append(q["krylov_iters"])
    levels, stt = 0, state
    while stt is not None:
        levels, stt = levels + 1, stt.coarse_state
    dev_gb = state.plan.device_bytes_now() / 1e9
    pt = settings.get("b200 partition", {})
    own = slice(pt.get("owned_node_begin", 0), pt.get("owned_node_end", np.asarray(sol).shape[0]))
    a, b = np.asarray(sol)[own].ravel(), np.asarray(sol_jacobi)[own].ravel()
    d2, n2 = backend.comm_allreduce_host([float(np.dot(a - b, a - b)), float(np.dot(b, b))])
    diff = float(np.sqrt(d2 / n2))
    step_ms = gmax(np.mean(step))
    total = (NX_OVERRIDE[0] or m) * m * m
    out = {"workload": workload_name(m, args.rtol, world).replace("Jacobi-PCG", "multigrid-PCG"),
           "newton_step_ms": step_ms, "elements_per_s": total / (step_ms * 1e-3),
           "e2e_ms_per_step": e2e_ms, "krylov_ms": gmax(np.mean(kry)), "krylov_iterations": float(np.mean(its)),
           "levels": levels, "hierarchy_build_s_first_call": t_build, "newton_steps": info[0], "res_norm": info[1],
           "rel_l2_vs_jacobi_pcg_solution": diff, "h2d_bytes_per_step": int(st["h2d_bytes"]), "plan_device_gb": dev_gb,
           "note": "krylov_ms includes the per-step set-up (coarse operators re-assembled at the injected state, power "
                   "iteration for the Chebyshev bounds)"}
    solver.clear_plan_cache()
    return out


def cpu_baseline_section(args, m):
    threads = host_threads()
    mref, _ = pick_ref_size(1, 25.0, args.ref_size, threads)
    dt, n_el, br = cpu_reference_step(mref, "cg", args.rtol, threads)
    return {"value": n_el / dt, "unit": UNIT, "cores": max(1, int(round(br["threads_busy"]))),
            "host_cores": os.cpu_count(), "threads_available": threads, "kind": "port",
            "sample": cpu_sample_note(mref, m, n_el) + " (%.1f s)" % dt, "breakdown_s": br}


def optional_section(fun, *a):
    """The extra sections of the line (nf = 3 figures, multigrid variant, CPU baseline) must not take the headline
    measurement down with them: an exception is reported in place of the section."""
    try:
        return fun(*a)
    except Exception as e:          # noqa: BLE001 -- reported, not swallowed
        import traceback
        return {"error": "%s: %s" % (type(e).__name__, e), "traceback": traceback.format_exc().splitlines()[-4:]}


# ---- own arm -----------------------------------------------------------------------------------------------------
def run_b200(args):
    from autopdex_b200 import backend, solver
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if _device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the b200 backend has no CPU path")
    backend.set_device(local_rank)
    if world > 1:
        from torch.distributed import TCPStore  # plumbing only: ships the NCCL id between ranks
        store = TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1,
                         world, rank == 0)
        if rank == 0:
            store.set("apdx_nccl_id", backend.comm_unique_id())
        backend.comm_init(bytes(store.get("apdx_nccl_id")), rank, world)

    m = args.size
    t_mesh = time.perf_counter()
    settings, static_settings, owned_elems = build_problem(m, rank, world)
    t_mesh = time.perf_counter() - t_mesh
    n_local = settings["node coordinates"].shape[0]
    dofs0 = np.zeros((n_local, 1))
    total_elems = (NX_OVERRIDE[0] or m) * m * m

    def barrier():
        backend.comm_allreduce_host(np.zeros(1))

    # ---- end-to-end through the public API (host buffers in, host solution out) ----
    t_plan = time.perf_counter()
    sol, info = solver.solver(dofs0, settings, static_settings, tol=args.rtol)       # builds + caches the plan
    t_plan = time.perf_counter() - t_plan
    # the settings buffers are page-locked in place the second time they are uploaded (backend.pin_if_repeated):
    # at least two untimed calls after the plan-building one, whatever --warmup says, so that no timed call pays for it
    for _ in range(max(args.warmup - 1, 2)):
        sol, info = solver.solver(dofs0, settings, static_settings, tol=args.rtol)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    launches_e2e = 0
    for _ in range(args.steps):
        sol, info = solver.solver(dofs0, settings, static_settings, tol=args.rtol)
        launches_e2e += solver.last_stats["kernel_launches"]
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    e2e_s = float(backend.comm_allreduce_host([e2e_s], "max")[0])
    st_e2e = dict(solver.last_stats)

    # ---- device-resident steps: apdx_newton on HBM-resident inputs, CUDA-event timing inside the plan ----
    state = next(iter(solver._PLAN_CACHE.values()))
    plan = state.plan
    opts = backend.KrylovOptions("cg", rtol=args.rtol, jacobi=True)
    step_ms, asm_ms, res_ms, kry_ms, iters, launches = [], [], [], [], [], 0
    for i in range(args.warmup + args.steps):
        state.dofs_d.zero()
        if i == args.warmup:
            barrier()
        n_it, rn, div = plan.newton(opts, state.dofs_d, state.vals_d, 1e-8, 30, 1.0)
        s = plan.stats()
        if i >= args.warmup:
            step_ms.append(s["total_ms"]); asm_ms.append(s["assembly_tangent_ms"]); res_ms.append(s["assembly_residual_ms"])
            kry_ms.append(s["krylov_ms"]); iters.append(s["krylov_iters"]); launches += s["kernel_launches"]
    barrier()
    ms = float(backend.comm_allreduce_host([np.mean(step_ms)], "max")[0])
    spmv_ms = plan.time_spmv(50)
    spmv_ms = float(backend.comm_allreduce_host([spmv_ms], "max")[0])
    dev_gb_after_solve = plan.device_bytes_now() / 1e9
    comm_info = plan.comm_info() if world > 1 else None
    clocks = sampler.stop()

    # ---- parity guard on the full-size run: residual norm and a discrete maximum principle ----
    owned = slice(settings.get("b200 partition", {}).get("owned_node_begin", 0),
                  settings.get("b200 partition", {}).get("owned_node_end", n_local))
    sol_sum = float(backend.comm_allreduce_host([np.asarray(sol)[owned].sum()])[0])
    ok = (info[0] == 1) and (not info[2]) and info[1] < 1e-8

    nnz, nfree = plan.nnz_reduced, plan.n_free
    rows = plan.f1 - plan.f0
    hbm, which = peaks()
    # SpMV roofline.  PHYSICAL bytes per launch = what the kernel streams from / to HBM: the sliced-ELL arrays (stored
    # values, compressed indices, mirror tables, slice headers) + x read + y written + the dot-product operand read;
    # cross-checked by ncu's dram__bytes (profiles/r02a_k_spmv_sell_p256_full.txt: 2.227 GB against 2.327 GB computed at
    # 256^3 on one GPU).  The ALGORITHMIC figure of SURVEY.md 8d (CSR bytes nnz*12 + n*16 + (n+1)*4) is reported beside
    # it: it exceeds the peak because symmetric storage + per-slice index compression stream fewer bytes than CSR.
    spmv_alg_bytes = nnz * 12 + rows * 16 + (rows + 1) * 4
    spmv_bytes = plan.stats()["sell_bytes"] + rows * 24
    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    spmv_alg_gbs = spmv_alg_bytes / (spmv_ms * 1e-3) / 1e9
    mean_iters = float(np.mean(iters))
    # CG iteration: SpMV + the two vector kernels (k_cg_update: q r minv -> r; k_cg_p: p x r minv -> x p = 10 streams)
    cg_iter_bytes = spmv_bytes + 80 * rows
    cg_gbs = cg_iter_bytes * mean_iters / (np.mean(kry_ms) * 1e-3) / 1e9
    cg_alg_gbs = (spmv_alg_bytes + 72 * rows) * mean_iters / (np.mean(kry_ms) * 1e-3) / 1e9
    asm_elems = state.plan.sets[0].conn.shape[0]
    asm_gbs = asm_elems * 290.0 / (np.mean(asm_ms) * 1e-3) / 1e9
    fp64_peak = backend.measure_fp64_peak()                        # TFLOP/s, measured here (not in MEASURED_PEAKS.json)
    asm_tflops = asm_elems * HEX8_POISSON_FLOP / (np.mean(asm_ms) * 1e-3) / 1e12

    line = None if rank != 0 else {
        "metric": METRIC, "value": total_elems / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(m, args.rtol, world),
                   "l2": "inputs larger than L2 (reduced CSR %.2f GB per rank)" % (nnz * 12 / 1e9),
                   "parity_guard": {"newton_steps": info[0], "res_norm": info[1], "diverged": bool(info[2]), "ok": bool(ok),
                                    "solution_sum": sol_sum}},
        "clocks": clocks,
        "e2e": {"value": total_elems / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]),
                "plan_build_s_first_call": t_plan, "mesh_generation_s": t_mesh},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_spmv_sell<1> (sliced-ELL SpMV, compressed column indices, symmetric storage, fused p.Ap)",
                     "bound": "hbm", "achieved": spmv_gbs, "peak": hbm, "unit": "GB/s", "frac": spmv_gbs / hbm,
                     "frac_of_nominal_8TBs": spmv_gbs / 8000.0, "peak_source": which,
                     "bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms,
                     "traffic": NCU_TRAFFIC.get(m) if world == 1 else spmv_bytes,
                     "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, " + NCU_TRAFFIC_SOURCE)
                                       if (world == 1 and m in NCU_TRAFFIC) else
                                       "computed from the plan's array sizes (no ncu capture at this rank count / size)",
                     "note": "achieved = PHYSICAL bytes streamed per launch (sliced-ELL arrays + x + y + dot operand) / CUDA-event "
                             "time: an HBM utilisation.  algorithmic_* uses the CSR bytes nnz*12+n*16+(n+1)*4 of SURVEY.md 8d, which "
                             "the kernel does not move (symmetric storage: lower columns are read from their transposed position in "
                             "the L2; column offsets stored once per 64-row slice)",
                     "algorithmic_bytes_per_launch": spmv_alg_bytes, "algorithmic_gbs": spmv_alg_gbs,
                     "algorithmic_frac": spmv_alg_gbs / hbm, "sell": plan.sell_info()},
        "newton_step_ms": ms,
        "assembly": {"elements_per_s": asm_elems / (np.mean(asm_ms) * 1e-3), "ms": float(np.mean(asm_ms)),
                     "residual_only_ms": float(np.mean(res_ms)), "algorithmic_gbs": asm_gbs,
                     "frac_of_hbm": asm_gbs / hbm, "bytes_per_element": 290,
                     "flop_per_element": HEX8_POISSON_FLOP, "tflops_fp64": asm_tflops, "fp64_peak_measured_tflops": fp64_peak,
                     "frac_of_fp64_peak": asm_tflops / fp64_peak,
                     "note": "fused pass = element kernel (FP64-FMA bound) + deterministic segmented-reduction scatter (HBM bound)"},
        "cg": {"iterations": mean_iters, "ms_per_iteration": float(np.mean(kry_ms)) / max(mean_iters, 1),
               "streamed_gbs": cg_gbs, "frac_of_hbm": cg_gbs / hbm, "algorithmic_gbs": cg_alg_gbs,
               "algorithmic_frac": cg_alg_gbs / hbm, "krylov_ms": float(np.mean(kry_ms)),
               "note": "streamed = SpMV physical bytes + 80 B/row of vector streams per iteration"},
        "vector_problem": None,
        "comm": comm_info,      # N > 1: mailbox / NCCL all-reduce, peer-inbox / NCCL halo exchange and the set-up timing behind the choice
        "system": {"n_free": nfree, "nnz_reduced": nnz, "plan_device_gb": dev_gb_after_solve,
                   "plan_device_gb_note": "all device buffers of the plan after the solves (pattern, gather lists, element streams, sliced-ELL matrix, Krylov vectors)"},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = optional_section(cpu_baseline_section, args, m)     # host only; before the optional GPU sections
    # The optional sections run code that is younger than the headline path; a stall in one of them must not cost the
    # headline line: a watchdog prints the line without the section (rank 0 owns the line; on several GPUs the other ranks
    # follow a few seconds later) and ends the process.
    import threading

    def section_with_watchdog(name, fun, *a):
        def bail():
            if rank == 0:
                line[name] = {"error": "timed out after %.0f s (watchdog)" % args.mg_timeout_s}
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(args.mg_timeout_s + (0.0 if rank == 0 else 5.0), bail)
        dog.daemon = True
        dog.start()
        out = optional_section(fun, *a)
        dog.cancel()
        if rank == 0:
            line[name] = out

    if world == 1 and not args.no_vector:
        section_with_watchdog("vector_problem", vector_problem_figures, args, hbm, fp64_peak, args.vector_size)
    if not args.no_multigrid:
        # every rank takes part in the partitioned variant
        section_with_watchdog("multigrid", multigrid_figures, args, sol, rank, world)
    if rank != 0:
        return
    print(json.dumps(line))


def _device_count():
    from autopdex_b200 import _lib
    return _lib.device_count()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=256, help="elements per direction (BASELINE: 256)")
    ap.add_argument("--ref-size", type=int, default=128, help="largest CPU sample (elements per direction)")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="time bound of the whole --impl reference run")
    ap.add_argument("--ref-direct", action="store_true", help="--impl reference: also time spsolve on a 32^3 sample")
    ap.add_argument("--no-vector", action="store_true", help="skip the bounded 64^3 neo-Hooke figures")
    ap.add_argument("--vector-size", type=int, default=64, help="elements per direction of the bounded neo-Hooke sample")
    ap.add_argument("--no-multigrid", action="store_true", help="skip the multigrid-preconditioned variant of the step")
    ap.add_argument("--mg-timeout-s", type=float, default=240.0, help="watchdog of each optional section (nf = 3 figures, multigrid variant)")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nx", type=int, default=0, help="diagnostic: elements along the slowest index (default: --size)")
    args = ap.parse_args()
    NX_OVERRIDE[0] = args.nx
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
