set -u
out=gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > $out/r03b_bench_n2.json 2> $out/r03b_bench_n2.err
tail -c 400 $out/r03b_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open("$out/r03b_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
print("n2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "adj", round(d["adjoint"]["value"]))
print(json.dumps(d.get("one_call_all_gpus"))[:500])
PY
