for n in 12500000 10000000 100000000; do
for cfg in "0 0" "64 0" "64 128" "56 128" "64 96"; do
  set -- $cfg
  BN_B200_SPEC_LOG2=$1 BN_B200_SPEC_MIN_CHUNK=$2 timeout 300 python bench.py --n-total $n --steps 10 --no-cpu --no-grad --no-fp32 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('N $n log2 $1 min $2', round(j['ms_per_step'],4), repr(j['energy']), {k: round(v,3) for k,v in j['kernels_ms_per_step'].items() if k.startswith('it_')})
"
done
done
