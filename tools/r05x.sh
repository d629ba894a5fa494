# Full refresh after the lane-kernel division work: GPU tests, default bench (all sub-records), reference arm,
# ncu launch list of the headline command (durations + DRAM bytes), full capture of the lane kernel.
T=r05x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${T}_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_${T}_lane -f \
  python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/${T}_ncu_full.log 2>&1
python - <<'P'
import json
d=json.loads(open('gpurun_out/r05x_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('c5s',d.get('c5_sharded',{}).get('value'))
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    if k=='C1_latency': print(k,{a:(b['median_us'],b['bit_identical_to_oracle']) for a,b in v.items() if a!='what'}); continue
    if k=='cascade': print(k,{a:(b['value'],b['e2e']['value'],b['parity_check']['tracking_bit_identical'],b['cpu_baseline']['value']) for a,b in v.items()}); continue
    print(k, '%.3g QP/s'%v['value'], 'frac %.3f'%v['roofline']['frac'], v['parity_check'], 'e2e %.3g'%v['e2e']['value'] if 'e2e' in v else '')
    if k=='C4': print('   fast',v['fast_math_dmma'],'dense',v['dense_part'], 'viol', v['primal_infeasible_answers'])
r=json.loads(open('gpurun_out/r05x_bench_ref.json').read().strip().splitlines()[-1])
print('ref', r['value'], r.get('cpu_baseline'))
P
