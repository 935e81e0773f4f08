#!/usr/bin/env bash
# One gpurun call that settles everything written in the CPU-only session 3 of round 1 (no GPU minutes were left):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'
# Results land in gpurun_out/s3_*.  Steps are independent: a failure is logged and the script carries on.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
run() { local name=$1; shift; echo "=== $name: $*" | tee -a $OUT/s3_log.txt; ( timeout 900 "$@" ) > $OUT/s3_$name.txt 2>&1; echo "exit $?" | tee -a $OUT/s3_log.txt; }

# 1. the regular GPU suite (includes the new config-3 API test), then the opt-in variants
run pytest_gpu python -m pytest tests -m gpu -x -q
APDX_TEST_EXPERIMENTAL=1 run pytest_experimental python -m pytest tests -m gpu -q -k "occupancy"
# 2. SpMV A/B: storage mode x occupancy variant, checksum + time (P256 = the bench workload; nf = 3 at 96^3)
run spmv_ab_p256 python tools/time_spmv.py all poisson 256 50
run spmv_ab_neohooke96 python tools/time_spmv.py all neohooke 96 50
# 3. element kernel block size (same kernel, 8 warps per SM, blocks of 128 / 64 / 32 threads)
for b in 128 64 32; do APDX_ELEM_BLOCK=$b run asm_block_$b python tools/time_assembly.py poisson 256 5; done
# 4. the bench line with the default kernels and with each occupancy variant
run bench_default python bench.py --steps 3 --warmup 3
for b in 3 4; do APDX_SPMV_BPS=$b run bench_bps$b python bench.py --steps 3 --warmup 3 --no-cpu; done
APDX_SPMV_PIPE=1 run bench_pipe python bench.py --steps 3 --warmup 3 --no-cpu
# 5. launch list of the bench command (shares only: ncu serialises and cold-starts every launch)
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/s3_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu
tail -n 40 $OUT/s3_log.txt
