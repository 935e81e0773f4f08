import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionfinish(session, exitstatus):
    """Runs of the -m gpu suite against the emulated build (tests/emu, APDX_LIB=.../libapdx_b200_emu.so): report what the
    CUDA stand-in counted, and fail the session if a *_sync intrinsic named a lane that had already exited."""
    lib_path = os.environ.get("APDX_LIB", "")
    if not lib_path.endswith("_emu.so") or not os.path.exists(lib_path):
        return
    import ctypes
    import gc
    if "autopdex_b200.solver" in sys.modules:        # plans kept by the plan cache are not leaks
        sys.modules["autopdex_b200.solver"].clear_plan_cache()
    gc.collect()
    lib = ctypes.CDLL(lib_path)
    for f in ("emu_kernel_launches", "emu_live_device_bytes", "emu_live_device_blocks", "emu_strict_violations"):
        getattr(lib, f).restype = ctypes.c_longlong
    bad_zones = lib.emu_check_red_zones()
    line = ("[emu] kernel launches %d, live device blocks %d (%d bytes), exited-lane *_sync violations %d, damaged red zones %d"
            % (lib.emu_kernel_launches(), lib.emu_live_device_blocks(), lib.emu_live_device_bytes(), lib.emu_strict_violations(), bad_zones))
    print("\n" + line)
    if lib.emu_strict_violations() or bad_zones:
        session.exitstatus = 1
