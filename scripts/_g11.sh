set -u
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/bench_configs.py --only c3 --c3-batch 64
TFQB_PASS_SEQ=1 python scripts/bench_configs.py --only c3 --c3-batch 64
python scripts/bench_configs.py --only c1,c5
TFQB_PASS_SEQ=1 python scripts/bench_configs.py --only c1,c5
TFQB_JIT=0 python bench.py --batch 512 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('interp seq8', round(d['value']), 'adj', round(d['adjoint']['value']))"
TFQB_PASS_SEQ=1 TFQB_JIT=0 python bench.py --batch 512 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('interp seq1', round(d['value']), 'adj', round(d['adjoint']['value']))"
