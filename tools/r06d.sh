python bench.py > gpurun_out/r06d_bench.json 2> gpurun_out/r06d_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r06d_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r06d_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
for k in ('C5','C5D','C3','S0','S2'):
    v=d['configs'][k]; print(k,'%.3g'%v['value'], v['parity_check']['x_bit_identical'], 'e2e %.3g'%v['e2e']['value'], v.get('compressed_input'))
v=d['configs']['C4']; print('C4 %.4g'%v['value'], 'fast %.4g'%v['fast_math_dmma']['value'], v['fast_math_dmma']['gate']['within_1e-9'])
print('cascade', {a:(b['value'], b['e2e']['value'], b['parity_check']['tracking_bit_identical']) for a,b in d['configs']['cascade'].items()})
P
