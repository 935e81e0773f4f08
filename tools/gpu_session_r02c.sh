#!/usr/bin/env bash
# round 2, call C (1 GPU): plane-grouped SpMV schedule A/B (APDX_SPMV_PLANES=1|2|4|8), GPU suite, new bench.py lines
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r02c_pytest_gpu.txt 2>&1; echo "pytest exit $?"; tail -2 $OUT/r02c_pytest_gpu.txt
for P in 1 2 4 8; do
  APDX_TRACE=1 APDX_SPMV_PLANES=$P python tools/time_spmv.py poisson 256 50 2>> $OUT/r02c_sched.err | sed "s/^{/{\"planes\": $P, /" >> $OUT/r02c_spmv_planes_p256.jsonl
  APDX_SPMV_PLANES=$P python tools/time_spmv.py neohooke 96 50 2>/dev/null | sed "s/^{/{\"planes\": $P, /" >> $OUT/r02c_spmv_planes_neohooke96.jsonl
done
APDX_SELL_SYM=0 APDX_SPMV_PLANES=8 python tools/time_spmv.py poisson 256 50 2>/dev/null | sed "s/^{/{\"planes\": 8, /" >> $OUT/r02c_spmv_planes_p256.jsonl
grep -h "spmv schedule" $OUT/r02c_sched.err | sort | uniq
python - <<PY
import json
for f in ("r02c_spmv_planes_p256.jsonl", "r02c_spmv_planes_neohooke96.jsonl"):
    for l in open("$OUT/" + f):
        d = json.loads(l); print(f, "planes", d["planes"], "sym", d["sell_sym"], "ms %.4f" % d["ms"], "impl GB/s %.0f" % d["implementation_gbs"], d["y_sha1"])
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_spmv_sell -s 12 -c 1 -f -o $OUT/r02c_spmv_p256 python tools/time_spmv.py poisson 256 5 > $OUT/r02c_spmv_p256.log 2>&1
ncu -i $OUT/r02c_spmv_p256.ncu-rep --page raw --csv > $OUT/r02c_spmv_p256_raw.csv 2>/dev/null
timeout 900 python bench.py > $OUT/r02c_bench_n1.json 2> $OUT/r02c_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --ref-budget-s 15 > $OUT/r02c_bench_ref.json 2> $OUT/r02c_bench_ref.err; echo "ref exit $?"
python tools/run_config.py neohooke 128 2 > $OUT/r02c_config5_n1.json 2> $OUT/r02c_config5_n1.err; echo "config5 exit $?"
cut -c1-3000 $OUT/r02c_bench_n1.json; cut -c1-600 $OUT/r02c_bench_ref.json; cut -c1-1200 $OUT/r02c_config5_n1.json
