"""TEST INFRASTRUCTURE: build libapdx_b200_emu.so -- the product's CUDA translation units compiled with g++ against the
CUDA stand-in of tests/emu/include (kernels run as host fibers).  See include/cuda_runtime.h for what this does and does
not check.  Nothing in autopdex_b200 knows about this library: the tests load it by setting APDX_LIB explicitly.

    python tests/emu/build.py            # -> tests/emu/build/libapdx_b200_emu.so

The only source transformation is the launch syntax, which a host compiler cannot parse:
    k<T><<<grid, block, smem, stream>>>(args)   ->   emu::launch(emu::cfg(grid, block, smem, stream), k<T>, "k<T>")(args)
    extern __shared__ T name[];                 ->   T *name = (T *)emu::dyn_smem();
    asm volatile("prefetch...")                 ->   (dropped: a cache hint)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "autopdex_b200", "csrc")
# EMU_UBSAN=1: a second build with -fsanitize=undefined (alignment of the vector loads -- a misaligned double2 / int4
# access is a fault on the device and silently fine on x86 --, static array bounds, shifts, signed overflow, division by
# zero, float -> int conversions out of range); run with UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1
UBSAN = os.environ.get("EMU_UBSAN", "0") not in ("", "0")
# EMU_ASAN=1: a third build with AddressSanitizer on heap and globals only (--param asan-stack=0: the fibers switch stacks
# behind the sanitizer's back) -- host-side containers of api.cu / dist.cu / sell.cu, use-after-free across plan
# destruction, static __shared__ arrays (globals of the stand-in); run with
#   LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:abort_on_error=1
ASAN = os.environ.get("EMU_ASAN", "0") not in ("", "0")
OUT_DIR = os.path.join(HERE, "build_ubsan" if UBSAN else ("build_asan" if ASAN else "build"))
LIB = os.path.join(OUT_DIR, "libapdx_b200_emu_ubsan.so" if UBSAN else ("libapdx_b200_emu_asan.so" if ASAN else "libapdx_b200_emu.so"))
SAN_FLAGS = (["-fsanitize=undefined,float-cast-overflow", "-fno-sanitize=vptr"] if UBSAN else
             (["-fsanitize=address", "--param", "asan-stack=0"] if ASAN else []))
UNITS = ["api", "pattern", "elements", "elements_fast", "sell", "krylov", "multigrid", "dist"]
HEADERS = ["common.cuh", "krylov.cuh", "elements.cuh"]

_LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>();]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
_DYN_SMEM = re.compile(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;")
_PREFETCH = re.compile(r'asm\s+volatile\s*\(\s*"prefetch[^;]*;\s*"[^;]*\)\s*;')


def transform(text):
    def launch(m):
        name, cfg = m.group(1), " ".join(m.group(2).split())
        label = name.replace('"', "")
        return 'emu::launch(emu::cfg(%s), %s, "%s")(' % (cfg, name, label)

    text, n = _LAUNCH.subn(launch, text)
    text = _DYN_SMEM.sub(r"\1 *\2 = (\1 *)emu::dyn_smem();", text)
    text = _PREFETCH.sub("/* prefetch hint dropped */;", text)
    text = text.replace('"../../include/apdx_b200.h"', '"apdx_b200.h"')
    if "<<<" in text:
        raise RuntimeError("an untransformed kernel launch is left")
    if re.search(r"\basm\b", text):
        raise RuntimeError("inline PTX other than prefetch hints: teach tests/emu/build.py about it")
    return text, n


def newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(ROOT, "include", "apdx_b200.h"), os.path.abspath(__file__),
                                                        os.path.join(HERE, "include", "cuda_runtime.h"),
                                                        os.path.join(HERE, "include", "cub", "cub.cuh")]
    hdr_changed = False
    for h in HEADERS:
        src, dst = os.path.join(CSRC, h), os.path.join(OUT_DIR, h)
        text, _ = transform(open(src).read())
        if not os.path.exists(dst) or open(dst).read() != text:
            open(dst, "w").write(text)
            hdr_changed = True
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-omit-frame-pointer", "-Wall", "-Wno-unknown-pragmas",
             "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-sign-compare",
             "-I" + os.path.join(HERE, "include"), "-I" + OUT_DIR, "-I" + os.path.join(ROOT, "include"),
             "-include", "cuda_runtime.h"] + SAN_FLAGS
    objs = []
    jobs = []
    n_launch = 0
    for u in UNITS:
        src = os.path.join(CSRC, u + ".cu")
        cpp, obj = os.path.join(OUT_DIR, u + ".cpp"), os.path.join(OUT_DIR, u + ".o")
        text, n = transform(open(src).read())
        n_launch += n
        if not os.path.exists(cpp) or open(cpp).read() != text:
            open(cpp, "w").write(text)
        objs.append(obj)
        if force or hdr_changed or newer(cpp, obj) or any(newer(d, obj) for d in deps):
            jobs.append((u, subprocess.Popen(["g++"] + flags + ["-c", cpp, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    rt_src, rt_obj = os.path.join(HERE, "emu_runtime.cpp"), os.path.join(OUT_DIR, "emu_runtime.o")
    objs.append(rt_obj)
    if force or newer(rt_src, rt_obj) or newer(os.path.join(HERE, "include", "cuda_runtime.h"), rt_obj):
        jobs.append(("emu_runtime", subprocess.Popen(["g++"] + flags + ["-c", rt_src, "-o", rt_obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for name, p in jobs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("---- %s ----\n%s\n" % (name, out[-6000:]))
        elif verbose and out.strip():
            sys.stderr.write("---- %s ----\n%s\n" % (name, out[-3000:]))
    if failed:
        raise RuntimeError("emulated build failed")
    if jobs or not os.path.exists(LIB):
        subprocess.check_call(["g++", "-shared", "-o", LIB] + SAN_FLAGS + objs + ["-ldl"])
    # stand-in for libnccl.so.2 (the product dlopen()s NCCL; multi-rank emulated runs name this file in APDX_NCCL_LIB)
    nccl_src, nccl_lib = os.path.join(HERE, "fake_nccl.cpp"), os.path.join(OUT_DIR, "libfakenccl.so")
    if force or newer(nccl_src, nccl_lib):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", nccl_src, "-o", nccl_lib, "-ldl"])
    return LIB, n_launch


if __name__ == "__main__":
    lib, n = build(force="--force" in sys.argv, verbose=True)
    print("%s  (%d kernel launch sites rewritten)" % (lib, n))
