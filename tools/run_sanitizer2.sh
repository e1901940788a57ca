#!/bin/bash
# second sanitizer pass: memcheck over the whole small-index parity file (child processes = the CLI included), the multi-GPU / counter
# tests, the device EM and the index builder's small cases
S="compute-sanitizer --error-exitcode 9 --print-limit 20"
run() { local tag=$1; shift; local t0=$SECONDS; timeout 400 "$@" > gpurun_out/r2_san_$tag.log 2>&1; local rc=$?
        echo "== $tag rc=$rc $((SECONDS - t0)) s: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/r2_san_$tag.log | sort | uniq -c | tr '\n' ' ' | cut -c1-300)"; }
run memcheck_parity_all $S --tool memcheck --target-processes all python -m pytest tests/test_gpu_parity.py -q -m gpu -x
run memcheck_multi_em   $S --tool memcheck python -m pytest tests/test_gpu_multi.py tests/test_gpu_em.py -q -m gpu -x
run memcheck_build      $S --tool memcheck python -m pytest tests/test_gpu_build.py -q -m gpu -x -k "not bench_data_path and not many_sequences"
