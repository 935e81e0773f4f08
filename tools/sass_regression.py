#!/usr/bin/env python
"""Device-code regression check without a GPU: build the CUDA sources of a git revision next to the working tree and
compare the SASS (instruction text + encodings, addresses and whitespace dropped) of every kernel that exists in both.

    python tools/sass_regression.py <git-rev>

Used in the CPU-only session 3 of round 1 to show that adding opt-in kernel variants left the 102 kernels of the measured
default path bit-identical.  SpMV instantiations of older revisions (no OCC parameter) are matched with OCC = 2."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = ["api", "pattern", "elements", "elements_fast", "sell", "krylov", "multigrid", "dist"]


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    d, name, buf = collections.OrderedDict(), None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                d[name] = buf
            name, buf = m.group(1), []
        elif name:
            buf.append(re.sub(r"\s+", "", re.sub(r"/\*[0-9a-f]{4}\*/", "", line, count=1)))
    if name:
        d[name] = buf
    return d


def main(rev):
    tmp = tempfile.mkdtemp(prefix="apdx_sass_")
    tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "autopdex_b200/csrc", "include"], capture_output=True, check=True)
    subprocess.run(["tar", "-x", "-C", tmp], input=tar.stdout, check=True)
    subprocess.run(["make", "-C", os.path.join(tmp, "autopdex_b200", "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", os.path.join(ROOT, "autopdex_b200", "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    key = lambda n: re.sub(r"(k_spmv_sellILi\dELi\dELb\dE)Li2E", r"\1", n)
    bad = total = 0
    for u in UNITS:
        old = kernels(os.path.join(tmp, "autopdex_b200", "csrc", "build", u + ".o"))
        new = {key(n): v for n, v in kernels(os.path.join(ROOT, "autopdex_b200", "csrc", "build", u + ".o")).items()}
        old = {key(n): v for n, v in old.items()}
        diff = [n for n in old if n in new and new[n] != old[n]]
        gone = [n for n in old if n not in new]
        total += len(old)
        bad += len(diff) + len(gone)
        print("%-14s kernels at %s: %3d  identical: %3d  different: %d  missing: %d  (now: %d)"
              % (u, rev, len(old), len(old) - len(diff) - len(gone), len(diff), len(gone), len(new)))
        for n in diff + gone:
            print("     ", subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:140])
    print("%d of %d kernels bit-identical" % (total - bad, total))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "HEAD"))
