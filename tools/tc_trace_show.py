import sys
ev=[tuple(map(int,l.split())) for l in open(sys.argv[1])]
cnt={2600:0,2700:0,3000:0}
for t,tag in ev:
    if tag in cnt:
        cnt[tag]+=1; continue
    extra=' '.join(f"{k}:{v}" for k,v in cnt.items() if v)
    for k in cnt: cnt[k]=0
    print(f"{t:8d} {tag:5d}   [{extra}]")
