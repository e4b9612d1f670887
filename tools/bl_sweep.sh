#!/bin/bash
for r in 4 8 16; do for b in 8 16 32 64; do
  RL_BL_REFILL=$r RL_BL_BURST=$b timeout 60 python tools/bl_time.py 2>&1 | grep "BL refill"
done; done
