set -u
out=gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > $out/r03_bench_n8.json 2> $out/r03_bench_n8.err
tail -c 300 $out/r03_bench_n8.err
python - <<PY
import json
d=json.loads([l for l in open("$out/r03_bench_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
print("n8", round(d["value"]), "e2e", round(d["e2e"]["value"]), "adj", round(d["adjoint"]["value"]))
for k in ("c4_strong","sharded_state","one_call_all_gpus"):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 scripts/bench_sharded.py --qubits 36 --reps 2 > $out/r03_sharded_36q_8gpu.jsonl 2> $out/r03_sharded_36q.err
tail -c 300 $out/r03_sharded_36q.err
cat $out/r03_sharded_36q_8gpu.jsonl
