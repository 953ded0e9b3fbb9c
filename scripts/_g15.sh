set -u
out=gpurun_out
for v in 8 1 2; do
  TFQB_JIT_FWD_SEQ=$v python scripts/bench_sharded.py --qubits 34 --reps 3 --warmups 3 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('34q seq$v', d['seconds_per_circuit'])"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specialised_pass_kernels_parity" 2>&1 | tail -3
