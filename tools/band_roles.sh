#!/bin/bash
# per-role L2-resident throughput of the band kernel (each role alone on all SMs): kernel-only clocks from the stats
for rt in 1 2; do
for only in 1 2 3; do
  echo "=== OFA_BAND_RT=$rt OFA_BAND_ONLY=$only"
  OFA_BAND_STATS=1 OFA_BAND_RT=$rt OFA_BAND_ONLY=$only timeout 120 python tools/test_planar.py bandtime 2>&1 | grep -E "band stats" | awk 'NR%52==1 || NR%52==2 || NR%52==3 || NR%52==4' | grep -v " 0 CTAs" | head -40
done
done
