#!/bin/bash
# A/B of kernel variants built by tools/build_variant.py: seeding-stage times of an 8-event batch and, with INST=1,
# the executed-instruction totals of a 2-event run (ncu).  Usage: tools/ab.sh default m8 m16 ...
for v in "$@"; do
  if [ "$v" = default ]; then unset B200SEED_LIB; else export B200SEED_LIB=acts_b200/variants/$v.so; fi
  echo "== $v: $(python tools/stage_times.py 8 200 3 2>&1 | grep 'rep 2' | sed 's/.*seed \([0-9.]*\).*doublet_count \([0-9.]*\).*doublet_fill \([0-9.]*\).*seed_middles \([0-9.]*\).*/seed \1 count \2 fill \3 middles \4/')"
  if [ -n "$INST" ]; then
    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:"k_seed|k_doublets" --csv --log-file /tmp/inst_$v.csv python profiles/profile_driver.py --events 2 --reps 1 > /dev/null 2>&1
    python tools/inst_count.py /tmp/inst_$v.csv | head -4
  fi
done
