#!/bin/bash
# Round-2 evidence on one B200: the bench line, the launch list of the same command, one full ncu capture of the three
# per-batch kernels (a 2 M-read window each), results under gpurun_out/ (summaries are copied to profiles/ by hand).
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
export CFB_BENCH_PARITY_READS=2000
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --reads 2000000 --skip-arms e2e_byteform,cpu_baseline,gather > gpurun_out/r2_ncu_bench.json 2> gpurun_out/r2_ncu_bench.err
# launches of k_search_t / k_prep / k_score before the first timed-size window: parity check (text, byte form, packed) + two counter passes = 5 each
ncu --set full --clock-control none --import-source on -k regex:"^(k_search_t|k_prep|k_score)$" -s 15 -c 3 -o gpurun_out/r2_prof_main \
    python bench.py --steps 1 --warmup 1 --reads 2000000 --skip-arms e2e_byteform,e2e_text,cpu_baseline,gather > /dev/null 2> gpurun_out/r2_ncu_full.err
tail -3 gpurun_out/r2_bench_final.err | cut -c1-300
