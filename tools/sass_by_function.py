import csv,re,collections,sys
d=sys.argv[1]; units=float(sys.argv[2])
rows=list(csv.reader(open(d+'/src.csv'))); hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
ins=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    ins.append((int(r[idx['Address']],16), r[idx['Source']], int(r[idx['Instructions Executed']]), int(r[idx['Thread Instructions Executed']])))
base=ins[0][0]
line_of={}; cur=None
for l in open(d+'/lines.txt'):
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if m: line_of[int(m.group(1),16)]=cur
# function ranges from the source
import bisect
def ranges(path):
    out=[]; name=None
    for i,l in enumerate(open(path),1):
        m=re.match(r'(?:template.*\n)?__device__ __forceinline__ \S+ (\w+)\(',l) or re.match(r'__device__ __forceinline__ \S+ (\w+)\(',l)
        if m: out.append((i,m.group(1)))
    return out
R={'theta_core.cuh':ranges('/root/repo/theta_b200/csrc/theta_core.cuh'),'theta_n2.cuh':ranges('/root/repo/theta_b200/csrc/theta_n2.cuh')}
def cat(k):
    if not k: return 'unknown'
    f,l=k
    if f in R:
        rs=R[f]; j=bisect.bisect_right([a for a,_ in rs],l)-1
        return rs[j][1] if j>=0 else f
    if f=='theta_n2.cu': return 'kernel body'
    return f
agg=collections.Counter(); thr=collections.Counter(); ops=collections.defaultdict(collections.Counter)
for a,s,n,t in ins:
    c=cat(line_of.get(a-base)); agg[c]+=n; thr[c]+=t
    op=s.split(); o=op[0] if not op[0].startswith('@') else op[1]
    ops[c][o.split('.')[0]]+=n
print('total %.1f'%(sum(agg.values())/units))
for c,n in agg.most_common(16):
    print('%-18s %6.1f  act %.1f   %s'%(c,n/units,thr[c]/max(n,1),' '.join('%s:%.0f'%(o,k/units) for o,k in ops[c].most_common(10))))
