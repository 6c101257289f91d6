#!/bin/bash
# orthogonal seeder: subtree size below which the walk scans elements instead of descending
for v in default kd16 kd32 kd64; do
  lib=acts_b200/variants/$v.so; [ "$v" = default ] && lib=acts_b200/libacts_b200_seeding.so
  echo "$v: $(B200SEED_LIB=$lib python tools/orth_times.py 8 200 3 2>&1 | grep 'rep 2')"
done
