set -u
out=gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_device.py -m gpu -x -q -k "two_gpus or multi or sharded" 2>&1 | tail -4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $out/r03_bench_n2.json 2> $out/r03_bench_n2.err
tail -c 600 $out/r03_bench_n2.err
python - <<PY
import json
d=json.loads(open("$out/r03_bench_n2.json").read().strip().splitlines()[-1])
print("n2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "adj", round(d["adjoint"]["value"]))
for k in ("c4_strong","sharded_state","one_call_all_gpus"):
    print(k, json.dumps(d.get(k))[:600])
PY
