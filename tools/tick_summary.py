import csv,re,collections,sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
rows=collections.OrderedDict()
for row in r:
    d=rows.setdefault(row['ID'],{'name':re.sub(r'\(.*','',row['Kernel Name']).replace('void ','')+row['Block Size'],'grid':row['Grid Size']})
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']; m=row['Metric Name']
    if m=='gpu__time_duration.sum': v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    if 'bytes' in m: v = v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
    d[m]=v
tot=collections.OrderedDict()
for k,d in rows.items():
    t=tot.setdefault(d['name'],[0,0,0,0,0,0]); us=d['gpu__time_duration.sum']
    t[0]+=us; t[1]+=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0); t[2]+=1
    t[3]+=us*d.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',0); t[5]+=us*d.get('sm__warps_active.avg.pct_of_peak_sustained_active',0)
T=sum(t[0] for t in tot.values())
for k,t in tot.items(): print("%-30s n=%3d  %8.0f us %5.1f%%  %6.1f GB  %5.0f GB/s  fp64 %4.1f%%  warps %4.1f%%"%(k,t[2],t[0],100*t[0]/T,t[1]/1e9,t[1]/max(t[0],1)/1e3,t[3]/max(t[0],1),t[5]/max(t[0],1)))
print("total us %.0f"%T)
if len(sys.argv)>2:
    for k,d in rows.items():
        if sys.argv[2] in d['name']: print(d['name'],d['grid'],"%.0f us"%d['gpu__time_duration.sum'],"%.2f GB"%((d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0))/1e9))
