set -u
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_at_size.py -m gpu -x -q 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs --no-adjoint > $out/r03h_$tag.json 2> $out/r03h_$tag.err; python - <<PY
import json
d=json.load(open("$out/r03h_$tag.json"))
r=d["roofline"]
print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"],3), "gate ms", round(r["fp32"]["gate_pass_ms_per_step"],2), "exp ms", round(r["expectation_kernel"]["share_of_step"]*d["ms_per_step"],2))
PY
}
run exp128 A=1
run exp256 TFQB_JIT_EXP_THREADS=256
