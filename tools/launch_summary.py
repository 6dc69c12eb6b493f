import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    name=row['Kernel Name']
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    unit=row['Metric Unit']
    v = v/1e6 if unit=='ns' else v/1e3 if unit=='us' else v*1e3 if unit=='s' else v
    agg[re.sub(r'\(.*','',name)][0]+=1; agg[re.sub(r'\(.*','',name)][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print('%-64s n=%4d  %9.3f ms  %5.1f%%  avg %.3f ms'%(k[:64],v[0],v[1],100*v[1]/tot, v[1]/v[0]))
print('total %.3f ms over %d launches' % (tot, sum(v[0] for v in agg.values())))
