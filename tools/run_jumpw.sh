python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_tests11.log; tail -4 gpurun_out/r2_tests11.log
for V in "CFB_KEEP_SHORT=0" "CFB_KEEP_SHORT=1"; do echo "== $V"; env $V python tools/ab_probe.py 2000000 2>&1 | tail -1; done > gpurun_out/r2_ab3.txt
cat gpurun_out/r2_ab3.txt
