for cand in 24 32; do for refill in 12 16 20 24 28; do
  r=$(B200_WAVE_CAND=$cand B200_WAVE_REFILL=$refill python tools/wave_probe.py --dim 6 --n 6400 --reserve 7000000 --repeat 3 2>/dev/null | python -c "
import sys,json
best=max(json.loads(l)['cuts_per_s'] for l in sys.stdin if l.startswith('{'))
print(round(best))")
  echo "cand=$cand refill=$refill best=$r"
done; done
for depth in 2 4 6; do r=$(B200_WAVE_DEPTH=$depth python tools/wave_probe.py --dim 6 --n 6400 --reserve 7000000 --repeat 3 2>/dev/null | python -c "
import sys,json
best=max(json.loads(l)['cuts_per_s'] for l in sys.stdin if l.startswith('{'))
print(round(best))"); echo "depth=$depth best=$r"; done
for ml in 5000 10000 40000; do r=$(B200_WAVE_MIN_LIVE=$ml python tools/wave_probe.py --dim 6 --n 6400 --reserve 7000000 --repeat 3 2>/dev/null | python -c "
import sys,json
best=max(json.loads(l)['cuts_per_s'] for l in sys.stdin if l.startswith('{'))
print(round(best))"); echo "min_live=$ml best=$r"; done
