#!/bin/bash
# Arena size (chunks per batch) against the stage times of a 16-event batch.
for mb in 2048 8192 24576; do
  echo "ARENA_MB=$mb: $(B200SEED_ARENA_MB=$mb python tools/stage_times.py 16 200 4 2>&1 | grep 'rep 3' | sed 's/.*wall \([0-9.]*\).*seed \([0-9.]*\)  compact.*doublet_fill \([0-9.]*\).*seed_middles \([0-9.]*\).*launches \([0-9]*\).*/wall \1 seed \2 fill \3 middles \4 launches \5/')"
done
