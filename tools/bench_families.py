import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
ks={k['kernel']:k for k in d['roofline']['kernels']}
def g(n):
    k=ks.get(n); return "%.0f/%.3f"%(k['us_per_step'],k['frac']) if k else "-"
print(sys.argv[1].split('/')[-1], "%.2f ms"%d['ms_per_step'], d['clocks']['sm_mhz'], "fwd",g('pointwise/skip 1x1 fwd GEMM + BN stats'),"wgrad",g('pointwise/skip 1x1 wgrad GEMM'),"dgrad",g('pointwise/skip 1x1 dgrad (or eval fwd) GEMM'),"bn",g('BN backward (reduce + apply)'),"dwb19",g('depthwise 3x3 bwd 19x19'), "infer %.0f"%d['infer']['value'])
