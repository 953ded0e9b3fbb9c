set -u
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_at_size.py -m gpu -x -q 2>&1 | tail -5
run() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs > $out/r03c_$tag.json 2> $out/r03c_$tag.err; python - <<PY
import json
d=json.load(open("$out/r03c_$tag.json"))
print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), "adj", round(d["adjoint"]["value"]), "frac", round(d["roofline"]["frac"],3), "exp_share", round(d["roofline"]["expectation_kernel"]["share_of_step"],3), "parity", d["parity_vs_oracle"]["max_abs_err"], d["parity_vs_oracle"]["adjoint"]["max_abs_err"])
PY
}
run default A=1
run nolift TFQB_JIT_LIFT=0
run seq1 TFQB_JIT_FWD_SEQ=1 TFQB_JIT_ADJ_SEQ=1
run seq8 TFQB_JIT_FWD_SEQ=8 TFQB_JIT_ADJ_SEQ=8
run seq2 TFQB_JIT_FWD_SEQ=2 TFQB_JIT_ADJ_SEQ=2
run old TFQB_JIT_LIFT=0 TFQB_JIT_FWD_SEQ=1 TFQB_JIT_ADJ_SEQ=1
