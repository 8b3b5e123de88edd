#!/bin/bash
# A/B two builds through tools/step_breakdown.py (batch and per-call steps): tools/ab_steps.sh <old.so> [n]
old=$1; n=${2:-6400}
new=bensolve_b200/libbslv_poly_b200.so
cp $new /tmp/new.so
for rep in 1 2; do
  for which in old new; do
    if [ $which = old ]; then cp $old $new; else cp /tmp/new.so $new; fi
    echo "== $which"; B200_PHASES=1 python tools/step_breakdown.py $n 2>&1 | grep "^[12] \|host us" | sed 's/.*launch=/launch=/' | cut -c1-170 | tail -4
  done
done
cp /tmp/new.so $new
