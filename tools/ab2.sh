run() { echo "$1 cs=$2 ch=$3: $(B200SEED_LIB=$4 B200SEED_CLASS_STREAMS=$2 B200SEED_CHUNK_STREAMS=$3 python tools/stage_times.py 16 200 4 2>&1 | grep 'rep 3' | sed 's/.*wall \([0-9.]*\).*seed \([0-9.]*\)  compact.*doublet_fill \([0-9.]*\).*seed_middles \([0-9.]*\).*/wall \1 seed \2 fill \3 middles \4/')"; }
run flat 0 1 acts_b200/libacts_b200_seeding.so
run flat 1 2 acts_b200/libacts_b200_seeding.so
run flat 1 1 acts_b200/libacts_b200_seeding.so
run branchy 0 1 acts_b200/variants/branchy.so
run branchy 1 2 acts_b200/variants/branchy.so
