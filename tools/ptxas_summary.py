#!/usr/bin/env python
"""Per-kernel resource table from the ptxas logs of the last `make -C autopdex_b200/csrc` (registers, spills, stack,
static shared memory), demangled -- the static half of the occupancy story that needs no GPU.
    python tools/ptxas_summary.py > profiles/<round>_ptxas_resources.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"Compiling entry function '([^']+)' for '(sm_\w+)'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n[^\n]*Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?")


def main():
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "autopdex_b200", "csrc", "build", "*.ptxas.log"))):
        unit = os.path.basename(log)[:-len(".ptxas.log")]
        for m in PAT.finditer(open(log).read()):
            rows.append((unit,) + m.groups())
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("# unit  registers  stack_B  spill_store_B  spill_load_B  static_smem_B  arch  kernel")
    bad = 0
    for r, name in zip(rows, names):
        unit, _, arch, stack, ss, sl, regs, _bar, smem = r
        flag = " <-- spills" if int(ss) or int(sl) else ""
        bad += bool(flag)
        print("%-14s %4s %6s %6s %6s %7s  %s  %s%s" % (unit, regs, stack, ss, sl, smem or "0", arch, re.sub(r"\(.*", "", name), flag))
    print("# %d kernels, %d with spills" % (len(rows), bad))


if __name__ == "__main__":
    main()
