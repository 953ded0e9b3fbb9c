set -u
out=gpurun_out
./scripts/micro/corun 1024 > $out/r03a_corun.jsonl 2>&1
cat $out/r03a_corun.jsonl
for lb in 4 5 6; do
  TFQB_GATE_LOW_BITS=$lb python scripts/bench_sharded.py --qubits 34 --reps 2 > $out/r03a_34q_lowbits$lb.json 2> $out/r03a_34q_lowbits$lb.err
  cat $out/r03a_34q_lowbits$lb.json
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
