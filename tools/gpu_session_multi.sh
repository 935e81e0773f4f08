#!/usr/bin/env bash
# multi-GPU session: gpurun --gpus N -- 'bash tools/gpu_session_multi.sh N [tag]'
#   parity matrix {poisson, neohooke} x {cg, bicgstab} x {slab, rcb} against the oracle, config 5 (128^3 neo-Hooke),
#   bench at N ranks (default = mailbox all-reduce, APDX_COMM=nccl, APDX_TRACE component times)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r02b}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/multi_gpu_worker.py matrix 24 12 > $OUT/${TAG}_parity_n$N.txt 2>&1; echo "parity exit $?"
grep "multi-gpu parity" $OUT/${TAG}_parity_n$N.txt
# partitioned multigrid hierarchy (written without hardware, validated on emulated ranks: DESIGN.md 3.6 / 6b)
timeout 300 $TR --master-port 29515 tests/multi_gpu_worker.py mgmatrix 64 32 > $OUT/${TAG}_mg_parity_n$N.txt 2>&1; echo "mg parity exit $?"
grep "multi-gpu parity" $OUT/${TAG}_mg_parity_n$N.txt
timeout 300 $TR --master-port 29517 tools/run_config.py neohooke ${CFG5_SIZE:-128} 2 -1 slab multigrid > $OUT/${TAG}_config5_mg_n$N.json 2> $OUT/${TAG}_config5_mg_n$N.err; echo "config5 multigrid exit $?"
cat $OUT/${TAG}_config5_mg_n$N.json | cut -c1-1500
timeout 300 $TR --master-port 29521 tools/run_config.py neohooke ${CFG5_SIZE:-128} 2 > $OUT/${TAG}_config5_n$N.json 2> $OUT/${TAG}_config5_n$N.err; echo "config5 exit $?"
cat $OUT/${TAG}_config5_n$N.json | cut -c1-1500
timeout 240 $TR --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench exit $?"
APDX_HALO=nccl timeout 300 $TR --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 --no-multigrid > $OUT/${TAG}_bench_n${N}_halo_nccl.json 2> $OUT/${TAG}_bench_n${N}_halo_nccl.err; echo "bench halo=nccl exit $?"
APDX_HALO=inbox timeout 300 $TR --master-port 29537 bench.py --gpus $N --steps 5 --warmup 3 --no-multigrid > $OUT/${TAG}_bench_n${N}_halo_inbox.json 2> $OUT/${TAG}_bench_n${N}_halo_inbox.err; echo "bench halo=inbox exit $?"
APDX_COMM=nccl timeout 240 $TR --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_n${N}_nccl.json 2> $OUT/${TAG}_bench_n${N}_nccl.err; echo "bench nccl exit $?"
APDX_TRACE=1 timeout 240 $TR --master-port 29551 bench.py --gpus $N --steps 1 --warmup 1 > /dev/null 2> $OUT/${TAG}_trace_n$N.txt
grep -h "apdx trace" $OUT/${TAG}_trace_n$N.txt | sort | uniq -c | head
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_n$N.json", "$OUT/${TAG}_bench_n${N}_halo_nccl.json", "$OUT/${TAG}_bench_n${N}_halo_inbox.json", "$OUT/${TAG}_bench_n${N}_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "comm", d.get("comm"), "step ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "cg ms/it", d["cg"]["ms_per_iteration"], "ok", d["config"]["parity_guard"]["ok"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
