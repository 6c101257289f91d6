run() { echo "$1: $(B200SEED_LIB=$2 B200SEED_CLASS_STREAMS=0 B200SEED_CHUNK_STREAMS=1 python tools/stage_times.py 16 200 3 2>&1 | grep 'rep 2' | sed 's/.*doublet_fill \([0-9.]*\).*seed_middles \([0-9.]*\).*/fill \1 middles \2/')"; }
run fill4 acts_b200/libacts_b200_seeding.so
run fill3 acts_b200/variants/fill3.so
run fill5 acts_b200/variants/fill5.so
