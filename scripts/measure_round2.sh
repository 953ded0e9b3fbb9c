#!/usr/bin/env bash
# One measurement sweep for profiles/ (round 2), one GPU: bench lines (both
# arms), launch list, full ncu captures (forward / expectation / adjoint /
# sampling kernels), secondary configs at the BASELINE batches, sanitizer.
set -u
tag=${1:-r02}
out=gpurun_out
export TFQB_JIT_CACHE_DIR=off          # cold NVRTC everywhere: honest first-call times
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
unset TFQB_JIT_CACHE_DIR
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench_b256.csv \
    python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs > $out/${tag}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tfqb_jit" -s 10 -c 5 \
    -o $out/${tag}_fwd python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs \
    > $out/${tag}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tfqb_jit_pass|tfqb_jit_accum" -s 44 -c 7 \
    -o $out/${tag}_adj python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs \
    > $out/${tag}_ncu_c.log 2>&1
python scripts/bench_configs.py --c3-batch 256 --c4-batch 2048 > $out/${tag}_secondary_configs.jsonl 2> $out/${tag}_configs.err
ncu --set full --clock-control none --import-source on -k regex:"tree_|sample_kernel|sort_rows" -c 8 \
    -o $out/${tag}_sampling python scripts/bench_configs.py --only c3 --c3-batch 8 \
    > $out/${tag}_ncu_d.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_noisy_ops.py -m gpu -q -x \
    -k "matches_oracle" > $out/${tag}_sanitizer_noisy.txt 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "peer_memory_exchange" > $out/${tag}_sanitizer_peer.txt 2>&1
ls -la $out | tail -20
