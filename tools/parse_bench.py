import json, sys
for f in sys.argv[1:]:
    try:
        txt=[l for l in open(f).read().splitlines() if l.startswith('{')][-1]
        d=json.loads(txt)
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d.get('clocks'), d.get('gpu_launches'))
    if "kernels" in d: print("   ", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["kernels"].items()}, 'sum', round(sum(v["ms_per_step"] for v in d["kernels"].values()),3))
    if d.get('roofline'): print('   roofline', d['roofline']['kernel'], round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), 'render fwd+bwd GB/s', d['roofline'].get('render_fwd_bwd_GBps'))
    if 'cpu_baseline' in d: print('   cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
