# sweep of the specialised forward-pass geometry: groups threads min_blocks tiles_per_cta
for cfg in "1 128 5 1" "1 128 2 2" "1 128 1 4" "1 64 3 4" "1 64 2 8" "1 256 1 2" "2 128 1 2" "1 128 3 2"; do
  set -- $cfg
  TFQB_JIT_FWD_GROUPS=$1 TFQB_JIT_FWD_THREADS=$2 TFQB_JIT_FWD_MINB=$3 TFQB_JIT_FWD_TILES=$4 python bench.py --no-cpu-baseline --no-adjoint --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg $cfg', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pass avg ms', round(d['roofline']['avg_launch_ms'],2), 'exp share', round(d['roofline']['expectation_kernel']['share_of_step'],3))"
done
