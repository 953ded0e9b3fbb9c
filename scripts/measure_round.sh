#!/usr/bin/env bash
# One measurement sweep for profiles/: bench lines (both arms), launch list,
# full ncu captures of the specialised kernels, secondary configs.
set -u
tag=${1:-r01j}
out=gpurun_out
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench_b256.csv \
    python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tfqb_jit" -s 10 -c 5 \
    -o $out/${tag}_fwd python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline \
    > $out/${tag}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tfqb_jit_pass|tfqb_jit_accum" -s 42 -c 7 \
    -o $out/${tag}_adj python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline \
    > $out/${tag}_ncu_c.log 2>&1
python scripts/bench_configs.py > $out/${tag}_secondary_configs.jsonl 2> $out/${tag}_configs.err
python scripts/bench_sharded.py --qubits 32 --reps 3 > $out/${tag}_c5_32q.jsonl 2> $out/${tag}_c5_32q.err
ls -la $out
