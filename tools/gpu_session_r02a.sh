#!/usr/bin/env bash
# round 2, call A (1 GPU): GPU suite on the trimmed Krylov code, full ncu captures of the SpMV (scalar + nf = 3) and of
# the generic element kernel, assembly timings of the vector path
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r02a_pytest_gpu.txt 2>&1; echo "pytest exit $?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_spmv_sell -s 12 -c 1 -f -o $OUT/r02a_spmv_p256 python tools/time_spmv.py poisson 256 5 > $OUT/r02a_spmv_p256.log 2>&1
timeout 600 $NCU -k regex:k_spmv_sell -s 12 -c 1 -f -o $OUT/r02a_spmv_nf3_96 python tools/time_spmv.py neohooke 96 5 > $OUT/r02a_spmv_nf3_96.log 2>&1
timeout 600 $NCU -k regex:k_elements -s 2 -c 1 -f -o $OUT/r02a_k_elements_neohooke64 python tools/time_assembly.py neohooke 64 2 > $OUT/r02a_k_elements.log 2>&1
for f in r02a_spmv_p256 r02a_spmv_nf3_96 r02a_k_elements_neohooke64; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
done
python tools/time_assembly.py neohooke 64 5 > $OUT/r02a_asm_neohooke64.txt 2>&1
python tools/time_assembly.py neohooke 128 3 > $OUT/r02a_asm_neohooke128.txt 2>&1
python tools/time_spmv.py all poisson 256 50 > $OUT/r02a_spmv_p256.txt 2>&1
ls -la $OUT
timeout 600 python bench.py --no-cpu > $OUT/r02a_bench_n1.json 2> $OUT/r02a_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02a_launches_p256.csv python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/r02a_bench_ncu.log 2>&1
tail -3 $OUT/r02a_pytest_gpu.txt; cat $OUT/r02a_bench_n1.json
