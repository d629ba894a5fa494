python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02y_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print('c5s',d.get('c5_sharded',{}).get('value'))
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    if k=='C1_latency': print(k,{a:(b['median_us'],b['bit_identical_to_oracle']) for a,b in v.items() if a!='what'}); continue
    if k=='cascade': print(k,{a:(b['value'],b['e2e']['value'],b['parity_check']['tracking_bit_identical'],b['cpu_baseline']['value']) for a,b in v.items()}); continue
    print(k, '%.3g QP/s'%v['value'], 'frac %.3f'%v['roofline']['frac'], v['parity_check'], 'e2e %.3g'%v['e2e']['value'] if 'e2e' in v else '')
    if k=='C4': print('   fast',v['fast_math_dmma'],'dense',v['dense_part'], 'viol', v['primal_infeasible_answers'])
P
