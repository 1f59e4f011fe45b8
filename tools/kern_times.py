import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), (d.get('e2e_dropin') or {}).get('value'))
for k in d.get('kernels',[]): print('   %-20s %7.2f us' % (k['kernel'], k['us']))
