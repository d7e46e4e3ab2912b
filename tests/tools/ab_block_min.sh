# A/B of HMPC_BLOCK_MIN (developer tool, run on the GPU box)
for bm in 1 2 3; do
  echo "== BLOCK_MIN=$bm"
  HMPC_BLOCK_MIN=$bm python bench.py --no-cpu-baseline --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readlines()[-1])
cl=l.get('closed_loop',{})
print('value %.3e e2e %.3e closed %.3e changes %.2f' % (l['value'], l['e2e']['value'], cl.get('value',0), cl.get('mean_working_set_changes',0)), {k:round(v['value']) for k,v in l.get('other_configs',{}).items()})
"
done
python -m pytest tests/test_locomotion_host.py -q -m gpu 2>&1 | tail -3
