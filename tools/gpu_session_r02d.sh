#!/usr/bin/env bash
# round 2, call D (1 GPU): plane-grouped SpMV schedule with the corrected plane stride, BCOO export tests
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r02d_pytest_gpu.txt 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02d_pytest_gpu.txt
rm -f $OUT/r02d_spmv_planes_*.jsonl
for P in 1 2 4 8; do
  APDX_TRACE=1 APDX_SPMV_PLANES=$P python tools/time_spmv.py poisson 256 50 2>> $OUT/r02d_sched.err | sed "s/^{/{\"planes\": $P, /" >> $OUT/r02d_spmv_planes_p256.jsonl
  APDX_TRACE=1 APDX_SPMV_PLANES=$P python tools/time_spmv.py neohooke 96 50 2>> $OUT/r02d_sched.err | sed "s/^{/{\"planes\": $P, /" >> $OUT/r02d_spmv_planes_neohooke96.jsonl
done
APDX_SELL_SYM=0 APDX_SPMV_PLANES=8 python tools/time_spmv.py poisson 256 50 2>/dev/null | sed "s/^{/{\"planes\": 8, /" >> $OUT/r02d_spmv_planes_p256.jsonl
grep -h "spmv schedule" $OUT/r02d_sched.err | sort | uniq
python - <<PY
import json
for f in ("r02d_spmv_planes_p256.jsonl", "r02d_spmv_planes_neohooke96.jsonl"):
    for l in open("$OUT/" + f):
        d = json.loads(l); print(f, "planes", d["planes"], "sym", d["sell_sym"], "ms %.4f" % d["ms"], "impl GB/s %.0f" % d["implementation_gbs"], d["y_sha1"])
PY
NCU="ncu --set full --clock-control none --import-source on"
APDX_SPMV_PLANES=8 timeout 600 $NCU -k regex:k_spmv_sell -s 12 -c 1 -f -o $OUT/r02d_spmv_p256_planes8 python tools/time_spmv.py poisson 256 5 > $OUT/r02d_spmv_p256.log 2>&1
ncu -i $OUT/r02d_spmv_p256_planes8.ncu-rep --page raw --csv > $OUT/r02d_spmv_p256_planes8_raw.csv 2>/dev/null
