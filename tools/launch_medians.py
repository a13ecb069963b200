import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; iK=hdr.index('Kernel Name'); iV=hdr.index('Metric Value'); iG=hdr.index('Grid Size'); iB=hdr.index('Block Size')
d=collections.defaultdict(list)
for r in rows[1:]: d[r[iK].split('(')[0][-40:]+' '+r[iG]+r[iB]].append(float(r[iV].replace(',','')))
for k,v in d.items(): print(k, 'n=%d'%len(v), 'median %.1f us'%(sorted(v)[len(v)//2]/1e3), 'min %.1f'%(min(v)/1e3))
