#!/bin/bash
# Round-2 ncu evidence, final kernel revision (run under gpurun; outputs in gpurun_out/, summaries copied to profiles/).
set -x
export B200SEED_CLASS_STREAMS=0 B200SEED_CHUNK_STREAMS=1   # kernels back to back: every launch is timed on its own
# 1. launch list of the bench command (reduced legs so that ncu's serialisation stays short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2h_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --value-streams 1 --e2e-threads 1 --no-relaxed --no-latency --no-orthogonal --no-strips --parity-events 0 --no-cpu-baseline --no-oracle-counters > gpurun_out/r2h_bench_under_ncu.log 2>&1
python profiles/summarise_launches.py gpurun_out/r2h_bench_launches.csv > gpurun_out/r2h_bench_launch_summary.csv 2>&1
# 2. DRAM traffic + instructions of the seeding / doublet kernels of one 16-event step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none \
  -k regex:"k_seed_middles|k_doublets" --csv --log-file gpurun_out/r2h_traffic.csv python tools/stage_times.py 16 200 1 > /dev/null 2>&1
python tools/inst_count.py gpurun_out/r2h_traffic.csv > gpurun_out/r2h_traffic_summary.txt 2>&1
# 3. full captures: class 0 and class 1 seeding kernels (first chunk of a 16-event batch), the two doublet passes
ncu --set full --clock-control none --import-source on -k regex:k_seed_middles -s 6 -c 2 -o gpurun_out/r2h_seed python profiles/profile_driver.py --events 16 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_doublets -s 2 -c 2 -o gpurun_out/r2h_doublets python profiles/profile_driver.py --events 16 --reps 2 > /dev/null 2>&1
ncu -i gpurun_out/r2h_seed.ncu-rep --page details > gpurun_out/r2h_seed_details.txt 2>&1
ncu -i gpurun_out/r2h_doublets.ncu-rep --page details > gpurun_out/r2h_doublets_details.txt 2>&1
ncu -i gpurun_out/r2h_seed.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2h_seed_cs.csv 2>&1
ncu -i gpurun_out/r2h_seed.ncu-rep --page source --csv --print-source sass > gpurun_out/r2h_seed_sass.csv 2>&1
python profiles/phase_breakdown.py gpurun_out/r2h_seed_cs.csv gpurun_out/r2h_seed_sass.csv > gpurun_out/r2h_seed_phase_breakdown.txt 2>&1
# 4. the orthogonal seeder's kernels (launch list of an 8-event batch)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_orth_launches.csv python tools/orth_times.py 8 200 2 > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/r2h_orth_launches.csv > gpurun_out/r2h_orth_launch_summary.csv 2>&1
ls -la gpurun_out/ | tail -20
