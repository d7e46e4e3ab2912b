# quick check of a kernel change on the GPU box: parity subset + timings of both classes (developer tool)
python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
HMPC_CFG=2 HMPC_B=1024 python tests/tools/gpu_check.py 2>&1 | python -c "
import json,sys
g=json.loads(sys.stdin.read())
print('walk1024 ms %.4f qps %.3e codes %s rel %.2e' % (g['kernel_ms'], g['qp_per_s'], g['codes'], g['rel_u0_max']))
print({k:v['med'] for k,v in g['stage_cycles'].items()}, g['cta_total_cycles'])
print({k:v for k,v in g['fine_cycles_med'].items() if k.startswith('s4')})
"
HMPC_CFG=1 HMPC_B=296 python tests/tools/gpu_check.py 2>&1 | python -c "
import json,sys
g=json.loads(sys.stdin.read())
print('stand296 ms %.4f qps %.3e codes %s rel %.2e' % (g['kernel_ms'], g['qp_per_s'], g['codes'], g['rel_u0_max']))
print({k:v['med'] for k,v in g['stage_cycles'].items()}, g['cta_total_cycles'])
"
python bench.py --no-cpu-baseline --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readlines()[-1])
cl=l.get('closed_loop',{})
print('value %.3e e2e %.3e closed %.3e' % (l['value'], l['e2e']['value'], cl.get('value',0)), {k:round(v['value']) for k,v in l.get('other_configs',{}).items()})
"
