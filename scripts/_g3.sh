set -u
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_at_size.py -m gpu -x -q -k "adjoint or adj or grad or c2 or c4 or specialised" 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs > $out/r03d_$tag.json 2> $out/r03d_$tag.err; python - <<PY
import json
d=json.load(open("$out/r03d_$tag.json"))
print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), "adj", round(d["adjoint"]["value"]), "frac", round(d["roofline"]["frac"],3))
PY
}
run default A=1
run nobatch TFQB_GRAD_BATCH=0
tmp=/tmp/ncu3; mkdir -p $tmp
# forward passes 0 and 1 of the second call (3 tfqb_jit_pass launches per call)
ncu --set full --clock-control none --import-source on -k regex:tfqb_jit_pass -s 3 -c 2 -o $tmp/fwd \
   python scripts/ncu_adjoint.py 128 expectation > $out/r03d_ncu_fwd.log 2>&1
ncu -i $tmp/fwd.ncu-rep --page raw --csv > $out/r03d_fwd_raw.csv 2>/dev/null
ncu -i $tmp/fwd.ncu-rep --page source --csv > $out/r03d_fwd_source.csv 2>/dev/null
# reverse pass 0 of the second adjoint call
ncu --set full --clock-control none --import-source on -k regex:tfqb_jit_pass -s 10 -c 1 -o $tmp/adj \
   python scripts/ncu_adjoint.py 128 adjoint > $out/r03d_ncu_adj.log 2>&1
ncu -i $tmp/adj.ncu-rep --page raw --csv > $out/r03d_adj_raw.csv 2>/dev/null
ncu -i $tmp/adj.ncu-rep --page source --csv > $out/r03d_adj_source.csv 2>/dev/null
# expectation pass 0
ncu --set full --clock-control none --import-source on -k regex:tfqb_jit_expect -s 2 -c 2 -o $tmp/exp \
   python scripts/ncu_adjoint.py 128 expectation > $out/r03d_ncu_exp.log 2>&1
ncu -i $tmp/exp.ncu-rep --page raw --csv > $out/r03d_exp_raw.csv 2>/dev/null
ncu -i $tmp/exp.ncu-rep --page source --csv > $out/r03d_exp_source.csv 2>/dev/null
ls -la $tmp $out | tail -20
du -sh $out
