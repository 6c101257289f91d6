#!/bin/bash
# e2e leg of bench.py (one event per b200seed_run call from 1 / 2 / 3 worker threads) for several builds of the library
for lib in "$@"; do
  B200SEED_LIB=$lib python bench.py --steps 3 --warmup 3 --e2e-threads 1,2,3 --no-relaxed --no-latency --no-orthogonal --no-strips --parity-events 0 --no-cpu-baseline --no-oracle-counters 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value'],1), d['e2e']['threads_sweep'])"
done
