run() { echo "$1: $(B200SEED_LIB=$2 B200SEED_CLASS_STREAMS=0 B200SEED_CHUNK_STREAMS=1 python tools/stage_times.py 8 200 3 2>&1 | grep 'rep 2' | sed 's/.*seed_middles \([0-9.]*\).*/middles \1/')"; }
run full acts_b200/libacts_b200_seeding.so
run noscan acts_b200/variants/noscan.so
