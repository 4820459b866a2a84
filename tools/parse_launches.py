import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
n=int(sys.argv[2]) if len(sys.argv)>2 else 7
for r in rows[h+1:][-n:]:
    print(r[4][:45].split('(')[0], r[-1])
