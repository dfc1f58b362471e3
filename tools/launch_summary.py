import csv,re,collections,sys
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
agg=collections.defaultdict(list)
for row in csv.DictReader(lines):
    name=re.sub(r'\(.*','',row['Kernel Name']).replace('void ','').replace('nadm::','')
    agg[name].append(float(row['Metric Value'].replace(',','')))
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:36s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us  min={min(v)/1e3:9.1f}  share={sum(v)/tot:.3f}")
