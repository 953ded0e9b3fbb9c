set -u
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs > $out/r03g_$tag.json 2> $out/r03g_$tag.err; python - <<PY
import json
d=json.load(open("$out/r03g_$tag.json"))
r=d["roofline"]
print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), "adj", round(d["adjoint"]["value"]), "frac", round(r["frac"],3), "gate ms", round(r["fp32"]["gate_pass_ms_per_step"],2), "exp ms", round(r["expectation_kernel"]["share_of_step"]*d["ms_per_step"],2))
PY
}
run default A=1
run minb6 TFQB_JIT_FWD_MINB=6
run minb4 TFQB_JIT_FWD_MINB=4
run adjminb3 TFQB_JIT_ADJ_MINB=3
run adj128 TFQB_JIT_ADJ_THREADS=128 TFQB_JIT_ADJ_MINB=4
python scripts/bench_configs.py --only c3 --c3-batch 64
