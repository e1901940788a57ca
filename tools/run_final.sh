#!/bin/bash
# final round-2 lines of all four single-GPU configs + the ncu evidence of the default one
if [ "${CFB_FINAL_TESTS:-1}" = 1 ]; then
    timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_tests_final.log 2>&1 || { tail -30 gpurun_out/r2_tests_final.log; echo "GPU TESTS FAILED -- no benches run"; exit 1; }
    tail -2 gpurun_out/r2_tests_final.log
fi
export CFB_REGEN_STATS=1
bash tools/run_profiles.sh
CFB_CLI_GBP=9 timeout 600 python tools/cli_bench.py > gpurun_out/r2_cli_bench.txt 2>&1; tail -12 gpurun_out/r2_cli_bench.txt | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 --lens 75-300 > gpurun_out/r2_bench_mixed.json 2> gpurun_out/r2_bench_mixed.err
timeout 1200 python bench.py --steps 20 --warmup 3 --paired --rdlen 150 --index-gbp 17 > gpurun_out/r2_bench_pe150_17g.json 2> gpurun_out/r2_bench_pe150_17g.err
rm -rf /tmp/cfb200_bench/cid_g1700_* /tmp/cfb200_bench/cid_g900_*
timeout 1500 python bench.py --steps 20 --warmup 3 --index-gbp 26 > gpurun_out/r2_bench_26g.json 2> gpurun_out/r2_bench_26g.err
for f in final mixed pe150_17g 26g; do echo "== $f"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_$f.json")); print(d["metric"], d["value"], d["e2e"]["value"], d["e2e_text"]["value"], d["parity_check"]["identical"], d["cpu_baseline"]["value"], d["config"]["tables"]["walk8_rows"], d["kernel_ms"])
except Exception as e: print("no json:", e)
PY
done
grep -h "regenerated" gpurun_out/r2_bench_*.err | sort | uniq -c | head -20
