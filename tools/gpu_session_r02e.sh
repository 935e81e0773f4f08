#!/usr/bin/env bash
# round 2, call E2 (1 GPU): multigrid at 256^3 -- parameter sweep, bench line with the multigrid section, GPU suite
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02e_pytest_gpu.txt 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02e_pytest_gpu.txt
rm -f $OUT/r02e_mg_sweep_p256.jsonl
# pre post coarsest ratio coarsest_ratio
for cfg in "2 2 12 4 40" "1 1 12 4 40" "3 3 12 4 40" "2 2 12 8 40" "2 2 12 3 40" "2 2 20 4 100" "1 2 12 4 40" "2 2 6 4 40"; do
  timeout 600 python tools/run_config.py poisson 256 multigrid $cfg 2>> $OUT/r02e_sweep.err | tail -1 >> $OUT/r02e_mg_sweep_p256.jsonl
done
python - <<PY
import json
for l in open("$OUT/r02e_mg_sweep_p256.jsonl"):
    try:
        d = json.loads(l)
    except Exception:
        print("bad line", l[:200]); continue
    g = d["mg"]
    print("pre %s post %s coarsest %s ratio %s cratio %s | iters %d krylov_ms %.1f total_ms %.1f e2e_ms %.1f first_s %.1f spmv %d launches %d res %.1e"
          % (g["pre"], g["post"], g["coarsest"], g["ratio"], g["coarsest ratio"], d["krylov_iters"], d["krylov_ms"], d["total_ms"], d["e2e_ms"], d["first_call_s"], d["spmv_launches"], d["kernel_launches"], d["newton"][1]))
PY
timeout 900 python bench.py --no-cpu > $OUT/r02e_bench_n1.json 2> $OUT/r02e_bench_n1.err; echo "bench exit $?"; tail -3 $OUT/r02e_bench_n1.err
python - <<PY
import json
d = json.loads(open("$OUT/r02e_bench_n1.json").read().strip().splitlines()[-1])
print("jacobi step ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"]); print(json.dumps(d.get("multigrid"))[:1500])
PY
