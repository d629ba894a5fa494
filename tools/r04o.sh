# host pipeline: upload-done event instead of stream sync, unit-code transport; parity + e2e
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or compressed or retry or multi_device" 2>&1 | tail -3
echo "== (value,row) transport"; MGQP_NO_UNIT_CODES=1 python tools/pack_probe.py 2>&1 | grep "e2e flags=0" 
echo "== unit codes"; python tools/pack_probe.py 2>&1 | grep "e2e flags=0"
python bench.py --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e'])"
