# sweep of the specialised adjoint-pass geometry: reg_bits threads min_blocks
for cfg in "3 256 2" "3 128 4" "3 128 5" "3 64 8" "4 256 1" "4 128 2" "4 128 3" "4 64 4" "4 64 6"; do
  set -- $cfg
  TFQB_ADJ_REGBITS=$1 TFQB_JIT_ADJ_THREADS=$2 TFQB_JIT_ADJ_MINB=$3 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['adjoint']; print('cfg $cfg', 'adjoint', round(a['value']), 'ms/step', round(a['ms_per_step'],1), 'frac', round(a['roofline']['frac'],3), 'fwd', round(d['value']))"
done
