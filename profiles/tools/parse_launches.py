import csv, sys
def load(path):
    with open(path) as f:
        lines=[l for l in f if not l.startswith('==')]
    rows=[]
    for row in csv.DictReader(lines):
        if row.get('Metric Name')=='gpu__time_duration.sum':
            v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
            ms = v/1e6 if u in('nsecond','ns') else (v/1e3 if u in ('usecond','us') else v)
            rows.append((int(row['ID']), row['Kernel Name'], ms, row['Grid Size'], row['Block Size']))
    return rows
rows=load(sys.argv[1])
starts=[i for i,x in enumerate(rows) if 'k_sample_cols' in x[1]]
seg=rows[starts[-1]:]
tot=0
for x in seg:
    name=x[1].replace('<unnamed>::','').replace('void ','')[:44]
    print(f"{x[0]:5d} {name:44s} {x[2]:9.4f} ms  {x[3]} {x[4]}"); tot+=x[2]
print('kernels', len(seg), 'sum ms', round(tot,4))
