import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    r=d.get('roofline') or {}
    ex=d.get('extra') or {}
    print(f, 'C2 %.3f Gev/s (%.3f ms)'%(d['value']/1e9, d['ms_per_step']), 'frac', round(r.get('frac') or 0,3),
          'lds_peak', (r.get('microbench_ops_per_s') or {}).get('lds32_conflict_free'),
          '| C3', ex.get('dcf_n64_u127_aes',{}).get('value'), '| HT', ex.get('halftree_n32_aes',{}).get('value'),
          '| EA', ex.get('dpf_evalall',{}).get('value'), '| e2e', (d.get('e2e') or {}).get('value'), '| clk', d.get('clocks',{}).get('sm_mhz'))
