# two GPUs: the library's multi-device entry point (test that skips on one GPU), bench under torchrun
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r05s_bench2.json 2> gpurun_out/r05s_bench2.err; echo "rc=$?"
tail -2 gpurun_out/r05s_bench2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r05s_bench2.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'frac',d['roofline']['frac'])
print('c5s',d.get('c5_sharded'))
print('multi',d.get('multi_device_library'))
P
