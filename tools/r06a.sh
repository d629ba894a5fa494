python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r06a_bench.json 2> gpurun_out/r06a_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r06a_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'], d['e2e']['clocks'], 'launches', d['gpu_launches'])
v=d['configs']['C4']; print('C4 %.4g'%v['value'], 'fast %.4g'%v['fast_math_dmma']['value'], v['fast_math_dmma']['gate']['within_1e-9'], v['parity_check'], 'dense', v['dense_part']['tflops'], v['dense_part']['dmma']['tflops'], v['dense_part']['what'][:40])
for k in ('C5','C5D','C3','S0','S2'):
    v=d['configs'][k]; print(k,'%.3g'%v['value'], v['parity_check']['x_bit_identical'], 'e2e %.3g'%v['e2e']['value'])
P
