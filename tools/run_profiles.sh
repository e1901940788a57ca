python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests4.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_bench tools/gather_bench.cu && /tmp/gather_bench > gpurun_out/r2_gather_bench.txt 2>&1
export CFB_BENCH_PARITY_READS=2000
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --reads 2000000 --skip-arms e2e_byteform,cpu_baseline,gather > gpurun_out/r2_ncu_bench.json 2> gpurun_out/r2_ncu_bench.err
ncu --set full --clock-control none --import-source on -k regex:"^(k_search_t|k_prep|k_score)$" -s 15 -c 3 -o gpurun_out/r2_prof_main python bench.py --steps 1 --warmup 1 --reads 2000000 --skip-arms e2e_byteform,e2e_text,cpu_baseline,gather > /dev/null 2> gpurun_out/r2_ncu_full.err
CFB_GROUP=8 CFB_KEEP_SIDES=1 ncu --set full --clock-control none --import-source on -k regex:"^k_search$" -s 2 -c 1 -o gpurun_out/r2_prof_coop8 python tools/ab_probe.py 500000 > gpurun_out/r2_probe_coop8.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_search_t$" -s 2 -c 1 -o gpurun_out/r2_prof_t500k python tools/ab_probe.py 500000 > gpurun_out/r2_probe_t.txt 2>&1
tail -4 gpurun_out/r2_tests4.log; tail -3 gpurun_out/r2_bench3.err; cat gpurun_out/r2_probe_coop8.txt | tail -3; cat gpurun_out/r2_probe_t.txt | tail -3
