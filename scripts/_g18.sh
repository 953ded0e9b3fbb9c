set -u
out=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 scripts/bench_sharded.py --qubits 36 --reps 3 --warmups 3 2> $out/r03_sharded_36q_final.err | grep '^{' > $out/r03_sharded_36q_8gpu_final.jsonl
tail -c 300 $out/r03_sharded_36q_final.err
cat $out/r03_sharded_36q_8gpu_final.jsonl
