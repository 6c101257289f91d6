for v in rf4 rf28; do
  B200SEED_LIB=acts_b200/variants/$v.so B200SEED_CLASS_STREAMS=0 B200SEED_CHUNK_STREAMS=1 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"k_seed" --csv --log-file /tmp/inst_$v.csv python profiles/profile_driver.py --events 2 --reps 1 > /dev/null 2>&1
  echo "== $v"; python tools/inst_count.py /tmp/inst_$v.csv 2>/dev/null | head -3
done
