python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02z_bench2.json 2> gpurun_out/r02z_bench2.err; echo "rc=$?"; tail -3 gpurun_out/r02z_bench2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02z_bench2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'n_gpus',d['n_gpus'])
print('c5s',d.get('c5_sharded'))
print('multi',d.get('multi_device_library'))
P
