python - <<'P'
import sys, json; sys.path.insert(0,'.')
import torch, mgqp_b200, bench
dev=torch.device('cuda:0')
peak,_=bench.peaks()
rec=bench.c4_record(mgqp_b200, dev, peak, 10000)
print('C4 %.4g QP/s'%rec['value'], 'fast %.4g'%rec['fast_math_dmma']['value'], rec['fast_math_dmma']['gate'])
print('dense', {k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('ms','tflops','frac')}) for k,v in rec['dense_part'].items() if k in ('what','ms','tflops','frac','dmma')})
P
