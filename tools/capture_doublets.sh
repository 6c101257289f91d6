#!/bin/bash
# ncu --set full of the two doublet passes (count, fill) of a 16-event batch, kernels back to back
export B200SEED_CLASS_STREAMS=0 B200SEED_CHUNK_STREAMS=1
ncu --set full --clock-control none --import-source on -k regex:k_doublets -s 2 -c 2 -o gpurun_out/s11_doublets python profiles/profile_driver.py --events 16 --reps 2 > gpurun_out/s11_ncu.log 2>&1
tail -5 gpurun_out/s11_ncu.log
ncu -i gpurun_out/s11_doublets.ncu-rep --page details > gpurun_out/s11_doublets_details.txt 2>&1
ncu -i gpurun_out/s11_doublets.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/s11_doublets_cs.csv 2>&1
ls -la gpurun_out | tail -4
