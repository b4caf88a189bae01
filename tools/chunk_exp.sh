for t in 0 37888 48829 56832; do
  BN_B200_CHUNK_TARGET=$t timeout 300 python bench.py --n-total 12500000 --steps 10 --no-cpu --no-grad --no-fp32 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('target $t', j['ms_per_step'], {k: round(v,3) for k,v in j['kernels_ms_per_step'].items()})
"
done
