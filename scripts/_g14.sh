set -u
out=gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT scripts/bench_sharded.py --qubits 36 --reps 2 2> $out/r03_sharded_36q_$tag.err | grep '^{' > $out/r03_sharded_36q_$tag.json; python - <<PY
import json
d=json.load(open("$out/r03_sharded_36q_$tag.json"))
print("$tag", d["seconds_per_circuit"], "exchange_s", d.get("exchange_seconds"), "GBps", d.get("nvlink_recv_GBps_per_gpu"), "passes", d.get("gate_passes"))
PY
}
PORT=29561 run seq1 TFQB_JIT_FWD_SEQ=1
PORT=29562 run seq2 TFQB_JIT_FWD_SEQ=2
PORT=29563 run nolift TFQB_JIT_LIFT=0
