import json, sys, glob
for f in sys.argv[1:] or sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f"{f}: {d['value']} {d['unit']}  ms/step {d['ms_per_step']}  e2e {d['e2e']['value']}  launches/step {d.get('launches_per_step')}  n_gpus {d['n_gpus']} clocks {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
    r=d['roofline']; print(f"   roofline: {r['kernel']} {r['bound']} {r['achieved']} {r['unit']} frac {r['frac']}")
    for k,v in d.get('kernels',{}).items(): print(f"     {k:10s} {v['ms_per_step']:9.4f} ms  x{v['launches']:2d}  {v['achieved']:9.1f} {v['unit']:8s} frac {v['frac']:.3f}  (alg {v['hbm_gbs']} GB/s)")
    if d.get('cpu_baseline'): print('   cpu:', d['cpu_baseline']['value'], d['cpu_baseline']['unit'], d['cpu_baseline']['cores'], 'cores')
