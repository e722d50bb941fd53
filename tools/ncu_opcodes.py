import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
tot=collections.Counter(); samples=collections.Counter()
for r in data:
    src=r[ix['Source']].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
    op=m.group(2) if m else src[:10]
    base=op if len(sys.argv)>2 else op.split('.')[0]
    try: n=float(r[ix['Instructions Executed']]); s=float(r[ix['# Samples']])
    except: n=0; s=0
    tot[base]+=n; samples[base]+=s
T=sum(tot.values()); S=sum(samples.values())
cvc=7.07e6
print("total",T, "per cvc", T/cvc)
for k,v in tot.most_common(45):
    print(f"{k:22s} {v:14.0f} {v/T*100:5.1f}%  per cvc {v/cvc:6.2f}   samples {samples[k]/S*100:5.1f}%")
