#!/bin/bash
# A/B two builds of the engine on the same box: tools/ab.sh <old.so> -- probe args...
old=$1; shift; shift
new=bensolve_b200/libbslv_poly_b200.so
cp $new /tmp/new.so
for rep in 1 2; do
  for which in old new; do
    if [ $which = old ]; then cp $old $new; else cp /tmp/new.so $new; fi
    echo -n "$which: "; B200_PHASES=1 python tools/probe.py "$@" 2>&1 | grep "host us" | sed 's/.*launch=/launch=/'
  done
done
cp /tmp/new.so $new
