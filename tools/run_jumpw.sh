python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_tests9.log; tail -3 gpurun_out/r2_tests9.log
for V in "CFB_JUMP_W=1 CFB_BIN_UNITS=0" "CFB_JUMP_W=1" "CFB_JUMP_W=2" "CFB_JUMP_W=3" "CFB_JUMP_W=4" "CFB_JUMP_W=1 CFB_FTABD=0"; do echo "== $V"; env $V python tools/ab_probe.py 2000000 2>&1 | tail -1; done > gpurun_out/r2_ab.txt
cat gpurun_out/r2_ab.txt
