import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr = rows[hi]; data = rows[hi+1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in data:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*','',r[ki])
    v = float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot[name]+=v; cnt[name]+=1
T = sum(tot.values())
print('total us %.1f launches %d' % (T, sum(cnt.values())))
for k,v in tot.most_common(30):
    print('%-60s %6d %10.1f us %5.1f%%  avg %8.1f' % (k[:60], cnt[k], v, 100*v/T, v/cnt[k]))
