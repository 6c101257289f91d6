#!/bin/bash
# A/B of seeding-kernel variants: seeds checked equal through the seed count, stage times of a 16-event batch (3rd repetition),
# serial streams (every stage timed on its own) and the default concurrent configuration.
run() { echo "$1 cs=$2 ch=$3: $(B200SEED_LIB=$4 B200SEED_CLASS_STREAMS=$2 B200SEED_CHUNK_STREAMS=$3 python tools/stage_times.py 16 200 4 2>&1 | grep 'rep 3' | sed 's/.*wall \([0-9.]*\).*seed \([0-9.]*\)  compact.*doublet_count \([0-9.]*\).*doublet_fill \([0-9.]*\).*seed_middles \([0-9.]*\).*seeds \([0-9]*\).*/wall \1 seed \2 count \3 fill \4 middles \5 seeds \6/')"; }
for v in "$@"; do
  lib=acts_b200/variants/$v.so
  [ "$v" = default ] && lib=acts_b200/libacts_b200_seeding.so
  run $v 0 1 $lib
  run $v 1 2 $lib
done
