python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_tests10.log; tail -3 gpurun_out/r2_tests10.log
for V in "CFB_JUMP_W=4" "CFB_JUMP_W=5" "CFB_JUMP_W=6" "CFB_JUMP_W=1"; do echo "== $V"; env $V python tools/ab_probe.py 2000000 2>&1 | tail -1; done > gpurun_out/r2_ab2.txt
cat gpurun_out/r2_ab2.txt
python tools/em_bench.py > gpurun_out/r2_em_bench.txt 2>&1; cat gpurun_out/r2_em_bench.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -3 gpurun_out/r2_bench5.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench5.json')); print(d['value'], d['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['host_ms_per_step_in_calls'], 'bf', d['e2e_byteform']['value'], d['e2e_byteform']['host_ms_per_step_in_calls'], 'text', d['e2e_text']['value'], d['roofline']['requests_per_unit'], d['roofline']['random_gather']['frac'], d['parity_check']['identical'])"
