python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02l_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_ref.json 2>/dev/null
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02l_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print('c5s',d.get('c5_sharded'))
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    if k=='C1_latency': print(k,{a:(b['median_us'],b['bit_identical_to_oracle'],b['cpu_one_thread']) for a,b in v.items() if a!='what'}); continue
    print(k, '%.3g QP/s'%v['value'], 'frac %.3f'%v['roofline']['frac'], v['parity_check'], v['status_counts'], 'e2e %.3g'%v.get('e2e',{}).get('value',0) if 'e2e' in v else '')
    print('   cpu', v['cpu_baseline'])
P
