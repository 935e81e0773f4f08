#!/bin/bash
# TEST INFRASTRUCTURE: every -m gpu test (except the 256^3 ones) against the emulated build of the CUDA source
# (tests/emu), once per memory-guard mode:
#   EMU_GUARD=1  every device block ends right in front of an inaccessible page  (over-runs fault, kernel/block/thread printed)
#   EMU_GUARD=2  every device block starts right behind one                        (under-runs fault)
#   EMU_GUARD=0  malloc'ed blocks between red zones, filled with 0xFF, freed blocks scribbled; 5 emulated SMs
#   EMU_ORDER=shuffle|reverse  blocks run in a shuffled / reversed order, threads of a block in reverse: results (several
#                tests assert bit-identical outputs) must not depend on the schedule
# The container has no GPU and no compute-sanitizer target; this is the memcheck that can run here.  Log -> profiles/.
set -u
cd "$(dirname "$0")/.."
python tests/emu/build.py || exit 1
LIB=$PWD/tests/emu/build/libapdx_b200_emu.so
OUT=${1:-profiles/r02g_emulated_cuda_source_suite.txt}
{
  echo "# $(date -u +%FT%TZ)  $(git rev-parse --short HEAD)  emulated CUDA source, -m gpu suite without tests/test_gpu_fullsize.py"
  for mode in "1 2 ascending" "2 2 ascending" "0 5 ascending" "1 2 shuffle" "1 2 reverse"; do
    set -- $mode
    echo "## EMU_GUARD=$1 EMU_SMS=$2 EMU_ORDER=$3 (order in which blocks, and threads inside a block, are executed)"
    EMU_ORDER=$3 EMU_GUARD=$1 EMU_SMS=$2 APDX_LIB=$LIB python -m pytest tests -q -m gpu --deselect tests/test_gpu_fullsize.py -p no:cacheprovider 2>&1 \
      | grep -E "^\[emu\]|passed|failed|error" 
  done
  for n in 2 4 8; do
    echo "## multi-rank: tests/multi_gpu_worker.py matrix on $n emulated ranks (fake NCCL over Unix sockets, CUDA IPC over POSIX shm: mailbox all-reduce, peer-inbox halo exchange forced, EMU_GUARD=1)"
    APDX_HALO=inbox EMU_GUARD=1 APDX_LIB=$LIB APDX_NCCL_LIB=$PWD/tests/emu/build/libfakenccl.so APDX_CASE_TIMEOUT=600 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) \
      tests/multi_gpu_worker.py matrix 12 8 2>&1 | grep "multi-gpu parity"
  done
  echo "## multi-rank: matrix on 2 emulated ranks, APDX_COMM=nccl APDX_HALO=nccl (ncclAllReduce + ncclSend/ncclRecv only)"
  APDX_COMM=nccl APDX_HALO=nccl EMU_GUARD=1 APDX_LIB=$LIB APDX_NCCL_LIB=$PWD/tests/emu/build/libfakenccl.so APDX_CASE_TIMEOUT=600 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 \
    tests/multi_gpu_worker.py matrix 12 8 2>&1 | grep "multi-gpu parity"
  for n in 2 3 4; do
    echo "## partitioned multigrid: tests/multi_gpu_worker.py mgmatrix on $n emulated ranks (peer-inbox halo exchange forced)"
    APDX_HALO=inbox EMU_GUARD=1 APDX_LIB=$LIB APDX_NCCL_LIB=$PWD/tests/emu/build/libfakenccl.so APDX_CASE_TIMEOUT=900 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29820 + n)) \
      tests/multi_gpu_worker.py mgmatrix 32 16 2>&1 | grep "multi-gpu parity"
    EMU_GUARD=1 APDX_LIB=$LIB APDX_NCCL_LIB=$PWD/tests/emu/build/libfakenccl.so APDX_CASE_TIMEOUT=900 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29840 + n)) \
      tests/multi_gpu_worker.py dae 8 2>&1 | grep "multi-gpu parity"
  done
} | tee "$OUT"
