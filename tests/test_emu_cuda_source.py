"""The product's CUDA source, executed on the host (tests/emu): the .cu translation units are compiled with g++ against a
functional stand-in for the CUDA runtime (threads of a block run as fibers; shuffles, votes and __syncthreads are real
rendezvous points; stream capture records and cudaGraphLaunch replays; device buffers sit in front of inaccessible
pages), and the -m gpu parity tests are run against that build through the unchanged C ABI.

This is test infrastructure for a container without a GPU: it checks indexing, buffer sizes, launch logic, capture
legality and numerics of the very source that nvcc compiles for sm_100a.  It is not a CPU path of the product (nothing in
autopdex_b200 refers to it; the library is only ever loaded through an explicit APDX_LIB) and it is no substitute for the
-m gpu run on a B200, which remains the parity gate.

The quick subset below runs in the CPU suite; `bash tools/emu_suite.sh` runs every -m gpu test (except the 256^3 ones) under
all three memory-guard modes and keeps the log under profiles/.
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the slowest emulated tests (20-40 s each: Cook's membrane load stepping, 1000+ Krylov iterations) stay in tools/emu_suite.sh
SLOW = [
    "tests/test_gpu_fullsize.py",
    "tests/test_gpu_api.py::test_cook_adaptive_load_stepping_golden",
    "tests/test_gpu_parity.py::test_g2_cook_load_stepping_golden",
    "tests/test_gpu_api.py::test_newton_maxiter_flags_divergence_like_reference",
    "tests/test_gpu_api.py::test_damped_newton_iteration_count_matches_oracle",
    "tests/test_gpu_api.py::test_tangent_solve_matches_reference_sensitivity_solve",
    "tests/test_gpu_api.py::test_cook_sensitivities_reproduce_the_reference_golden_values",
    "tests/test_zz_gpu_multigrid.py::test_poisson_hex_multigrid_matches_oracle",
    "tests/test_zz_gpu_multigrid.py::test_neo_hooke_brick_multigrid_newton_counts",
    "tests/test_zz_gpu_r02_dae.py::test_verbose_prints_one_line_per_newton_iteration",
    "tests/test_zz_gpu_r02_symmetric_tangent.py::test_cook_membrane_q1_multigrid_pcg_matches_jacobi_bicgstab",
]


def _build():
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    try:
        import build as emu_build
    finally:
        sys.path.pop(0)
    lib, n_sites = emu_build.build()
    assert n_sites >= 60          # every kernel launch of the product goes through the stand-in
    return lib


def test_gpu_suite_subset_on_emulated_cuda_source():
    lib = _build()
    env = dict(os.environ, APDX_LIB=lib, EMU_GUARD="1")
    cmd = [sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"]
    for s in SLOW:
        cmd += ["--deselect", s]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-4000:]
    assert " passed" in out and "failed" not in out
    summary = [l for l in out.splitlines() if l.startswith("[emu]")]
    assert summary, out[-2000:]
    # nothing leaked (every plan of the suite was destroyed), no shuffle / vote named an exited lane
    assert "live device blocks 0 " in summary[0] and "violations 0" in summary[0], summary[0]


def _run_ranks(n, worker_args, extra_env=None, timeout=1500):
    lib = _build()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, APDX_LIB=lib, APDX_NCCL_LIB=os.path.join(os.path.dirname(lib), "libfakenccl.so"), EMU_GUARD="1",
               APDX_CASE_TIMEOUT="300")
    env.pop("APDX_COMM", None)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")] + worker_args
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    out = r.stdout.decode()
    return r.returncode, [l for l in out.splitlines() if l.startswith("multi-gpu parity")], out


def test_multi_gpu_code_on_two_emulated_ranks():
    """dist.cu + the multi-rank Krylov loop (halo exchange of slab and list partitions, the collective partition check,
    CUDA graphs holding communication nodes, dot products through the peer-memory mailboxes = the default) on two host
    processes: tests/multi_gpu_worker.py -- the script that runs on 2/4/8 B200s -- with the emulated build, CUDA IPC
    emulated by POSIX shared memory and tests/emu/fake_nccl.cpp behind APDX_NCCL_LIB.
    {Poisson nf=1, neo-Hooke nf=3} x {CG, BiCGSTAB} x {slab, RCB} against the oracle."""
    rc, lines, out = _run_ranks(2, ["matrix", "8", "4"], {"APDX_HALO": "inbox", "APDX_TRACE": "1"})
    assert rc == 0 and len(lines) == 8 and all(l.endswith("-> OK") for l in lines), out[-4000:]
    assert "peer mapping unavailable" not in out      # the mailbox kernels ran (not the NCCL fall-back)
    # APDX_HALO=inbox: the halo exchange of the slab cases went through the peer inboxes (k_halo_push / k_halo_pull inside
    # the captured Krylov graphs) after their set-up self-test against the NCCL exchange; the default (auto) keeps
    # whichever is faster on the machine, which on emulated ranks is decided by noise
    assert "identical ghost entries -> inboxes" in out and "MISMATCH" not in out


def test_nccl_allreduce_mode_on_three_emulated_ranks():
    """APDX_COMM=nccl (ncclAllReduce + one-thread stage kernel) and an odd rank count."""
    rc, lines, out = _run_ranks(3, ["neohooke", "6", "bicgstab", "slab"], {"APDX_COMM": "nccl", "APDX_HALO": "nccl"})
    assert rc == 0 and len(lines) == 1 and lines[0].endswith("-> OK"), out[-4000:]


def test_partitioned_multigrid_on_emulated_ranks():
    """The multigrid-preconditioned CG with a hierarchy partitioned into slabs (halo exchange per level, restriction and
    prolongation across the interface, injected coarse state from the neighbour, all-reduced dot products): Poisson 16^3
    and the nonlinear neo-Hooke brick 8^3 (nf = 3, coarse tangents re-discretised in every Newton step) on 2 ranks against
    the oracle; the iteration count stays at the single-GPU level."""
    rc, lines, out = _run_ranks(2, ["mgmatrix", "16", "8"], {"APDX_HALO": "inbox"})
    assert rc == 0 and len(lines) == 3 and all(l.endswith("-> OK") for l in lines), out[-4000:]   # + the 2-D README problem
    its = int(lines[0].split("krylov_iters=")[1].split()[0])
    assert its <= 10, lines[0]


def test_time_stepping_manager_on_emulated_ranks():
    """dae.TimeSteppingManager (heat conduction, 'user residual' route; BackwardEuler, two-stage DIRK, Adams-Moulton 2) on two
    slabs, Jacobi- and multigrid-preconditioned CG (the coarse levels' 'dofs n' travels with the ghost-plane exchange),
    against the SciPy loops."""
    rc, lines, out = _run_ranks(2, ["dae", "8"])
    assert rc == 0 and len(lines) == 4 and all(l.endswith("-> OK") for l in lines), out[-4000:]


def test_readme_example_runs_on_the_emulated_build():
    """The usage example of README.md, at 8^3 instead of 64^3."""
    lib = _build()
    code = open(os.path.join(ROOT, "README.md")).read().split("```python")[1].split("```")[0].replace("(64, 64, 64)", "(8, 8, 8)")
    code += "\nprint('RESULT', n_newton, float(res_norm) < 1e-8, bool(diverged))\n"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, APDX_LIB=lib, EMU_GUARD="1"),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0 and "RESULT 1 True False" in r.stdout.decode(), r.stdout.decode()[-2000:]


def test_capture_rules_of_the_stand_in():
    """The stand-in must be as strict as the runtime where the product depends on it: an allocation or a synchronisation
    while a stream captures invalidates the capture (a scope-bound temporary freed inside the Krylov capture would be
    exactly this bug)."""
    import ctypes as C
    lib = C.CDLL(_build())
    assert lib.emu_selftest_capture() == 0
