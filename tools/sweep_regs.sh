run() { echo "$1 threads $2: $(B200SEED_LIB=$3 B200SEED_CLASS_THREADS=$2 python tools/stage_times.py 8 200 3 2>&1 | grep 'rep 2' | sed 's/.*seed \([0-9.]*\).*seed_middles \([0-9.]*\).*/seed \1 middles \2/')"; }
run r64 "160,256,320,512,1024,1024" acts_b200/libacts_b200_seeding.so
run r56 "192,288,384,576,1024,1024" acts_b200/variants/r56.so
run r56 "160,256,352,576,1024,1024" acts_b200/variants/r56.so
run r48 "224,320,448,640,1024,1024" acts_b200/variants/r48.so
run r48 "192,320,416,640,1024,1024" acts_b200/variants/r48.so
run r48 "160,256,320,512,1024,1024" acts_b200/variants/r48.so
