#!/bin/bash
# compute-sanitizer over the small end of the GPU suite: memcheck on smoke() and on the parity tests that force every derived-table
# combination, the regeneration buffer, packed input, degenerate batches and the text operator's span handling; racecheck on smoke().
S="compute-sanitizer --error-exitcode 9 --print-limit 20"
run() { local tag=$1; shift; local t0=$SECONDS; timeout 420 "$@" > gpurun_out/r2_san_$tag.log 2>&1; local rc=$?
        echo "== $tag rc=$rc $((SECONDS - t0)) s: $(grep -E 'ERROR SUMMARY|passed|failed|smoke ok' gpurun_out/r2_san_$tag.log | tr '\n' ' ' | cut -c1-300)"; }
run memcheck_smoke   $S --tool memcheck python __graft_entry__.py --smoke
run memcheck_layouts $S --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "every_device_layout or packed_input or empty_and_degenerate or pipelined_and_resident"
run memcheck_text    $S --tool memcheck python -m pytest tests/test_gpu_text.py -q -m gpu -x
run racecheck_smoke  $S --tool racecheck python __graft_entry__.py --smoke
