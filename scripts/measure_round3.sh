#!/usr/bin/env bash
# Measurement sweeps for profiles/ (round 2, second half: `r03*` files), one GPU.
#   measure_round3.sh bench <tag>   bench lines (both arms), secondary configs at
#                                   the BASELINE batches, sanitizer run, smoke
#   measure_round3.sh ncu <tag>     launch list + full ncu captures (raw and
#                                   per-instruction source pages as CSV; the
#                                   .ncu-rep files stay on the box)
set -u
what=${1:-bench}
tag=${2:-r03}
out=gpurun_out
if [ "$what" = bench ]; then
  export TFQB_JIT_CACHE_DIR=off          # cold NVRTC everywhere: honest first-call times
  python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
  python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
  unset TFQB_JIT_CACHE_DIR
  python scripts/bench_configs.py --c3-batch 256 --c4-batch 2048 > $out/${tag}_secondary_configs.jsonl 2> $out/${tag}_configs.err
  compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/memcheck_jit.py > $out/${tag}_sanitizer_jit.txt 2>&1
  python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1
else
  tmp=/tmp/ncu_$tag
  mkdir -p $tmp
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/${tag}_launches_bench_b256.csv \
      python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs > $out/${tag}_ncu_a.log 2>&1
  # forward passes 0-2 and the two expectation passes of the second call
  ncu --set full --clock-control none --import-source on -k regex:"tfqb_jit_pass|tfqb_jit_expect" -s 5 -c 5 \
      -o $tmp/fwd python scripts/ncu_adjoint.py 128 expectation > $out/${tag}_ncu_b.log 2>&1
  ncu -i $tmp/fwd.ncu-rep --page raw --csv > $out/${tag}_forward_expect_full_raw.csv 2>/dev/null
  # second adjoint call: forward x3, accumulate x2, reverse x4
  ncu --set full --clock-control none --import-source on -k regex:"tfqb_jit_pass|tfqb_jit_accum" -s 12 -c 6 \
      -o $tmp/adj python scripts/ncu_adjoint.py 128 adjoint > $out/${tag}_ncu_c.log 2>&1
  ncu -i $tmp/adj.ncu-rep --page raw --csv > $out/${tag}_adjoint_accum_full_raw.csv 2>/dev/null
  ls -la $tmp
fi
du -sh $out
