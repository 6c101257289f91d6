#!/bin/bash
# register cap x threads per class of k_seed_middles (16-event batch, serial streams: seed_middles is the sum of its launches)
run() { echo "$1 threads=$3: $(B200SEED_LIB=$2 B200SEED_CLASS_THREADS=$3 B200SEED_CLASS_STREAMS=0 B200SEED_CHUNK_STREAMS=1 python tools/stage_times.py 16 200 3 2>&1 | grep 'rep 2' | sed 's/.*wall \([0-9.]*\).*seed_middles \([0-9.]*\).*seeds \([0-9]*\).*/wall \1 middles \2 seeds \3/')"; }
run default acts_b200/libacts_b200_seeding.so 192,288,384,576,1024,1024
run regs48 acts_b200/variants/regs48.so 192,288,384,576,1024,1024
run regs48 acts_b200/variants/regs48.so 224,336,448,672,1024,1024
run regs40 acts_b200/variants/regs40.so 256,384,512,768,1024,1024
run regs64 acts_b200/variants/regs64.so 160,256,320,512,1024,1024
run regs64 acts_b200/variants/regs64.so 192,288,384,576,1024,1024
