import json, sys
for f in sys.argv[1:]:
    d=json.load(open(f))
    print(f, 'value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'device_ms',round(d['device_ms_per_proof'],3),'launches/proof',d['gpu_launches']//d['steps'], d['clocks'])
    print(' stages',{k[:18]:v for k,v in d['stages_ms'].items()})
    for k in d['kernels']: print('   %-24s %8.4f ms  x%-3d %s GB/s frac %s' % (k['name'],k['ms'],k['launches'],k['gbps'],k['frac']))
    print(' cpu',d.get('cpu_baseline'))
