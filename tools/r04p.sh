python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or compressed or retry or multi_device" 2>&1 | tail -3
for ms in 20 250 0; do python bench.py --no-extras --no-cpu --e2e-clock-ms $ms 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('period $ms: value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['clocks'], d['clocks'])"; done
