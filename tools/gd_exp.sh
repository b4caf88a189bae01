for cap in 0 2 3; do for rep in 1 2; do
  BN_B200_GD_CTAS=$cap timeout 200 python tools/bench_small_d.py 200000 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('cap $cap', {k:(round(v['filter_ms'],2), round(v['smoother_ms'],2)) for k,v in j.items() if k!='N'})"
done; done
