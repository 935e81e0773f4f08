#!/usr/bin/env bash
# Multi-GPU companion of tools/gpu_session.sh (N = 2 or 4 GPUs of one box):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_session_multi.sh 2'
# Slab parity (regression), then the list-based (recursive coordinate bisection) partition written in the CPU-only
# session 3 -- its device half (k_halo_pack + grouped ncclSend/ncclRecv per neighbour) has never run -- then the bench.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "=== $name: $*" | tee -a $OUT/s3m_log.txt; ( timeout 600 "$@" ) > $OUT/s3m_$name.txt 2>&1; echo "exit $?" | tee -a $OUT/s3m_log.txt; }
port=29700
for part in slab rcb; do
  for kry in cg bicgstab; do
    port=$((port + 10))
    run parity_${part}_${kry} $TR --master-port $port tests/multi_gpu_worker.py 20 $kry $part
  done
done
port=$((port + 10)); run parity_rcb_cg_33 $TR --master-port $port tests/multi_gpu_worker.py 33 cg rcb
port=$((port + 10)); run bench $TR --master-port $port bench.py --gpus $N --steps 3 --warmup 3
port=$((port + 10)); APDX_COMM=mbox run bench_mbox $TR --master-port $port bench.py --gpus $N --steps 3 --warmup 3
grep -h "multi-gpu parity" $OUT/s3m_parity_*.txt | tee -a $OUT/s3m_log.txt
tail -n 30 $OUT/s3m_log.txt
