for i in 1 2 3; do python bench.py --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'], d['e2e']['clocks'])"; done
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K=3: value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'], d['e2e']['clocks'])"
