#!/bin/bash
# A/B of the count->fill hand-off (windows + survivor masks): B200SEED_MASK_WORDS_PER_SP = 0 (off) / 192 (default)
for m in 0 192 0 192; do
  export B200SEED_MASK_WORDS_PER_SP=$m
  echo "mask words per sp = $m"
  bash tools/ab5.sh default
done
