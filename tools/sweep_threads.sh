for t in "128,128,160,192,224,288,384,576,1024,1024" "96,128,160,192,224,288,384,576,1024,1024" "128,128,128,192,224,288,384,576,1024,1024" "128,160,160,192,224,288,384,576,1024,1024"; do
  echo "threads $t: $(B200SEED_CLASS_THREADS=$t B200SEED_CHUNK_STREAMS=1 python tools/stage_times.py 8 200 3 2>&1 | grep 'rep 2' | sed 's/.*seed \([0-9.]*\).*seed_middles \([0-9.]*\).*/seed \1 middles \2/')"
done
