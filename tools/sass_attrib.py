import re,csv,collections,sys
dis=sys.argv[1]; srccsv=sys.argv[2]
lines=open(dis).read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('//--------------------- .text._ZN3hdb18DeblockTiledKernel')][0]
chain=[]; seq=[]; pending=[]
for l in lines[start+1:]:
    if l.startswith('//---------------------'): break
    if l.strip().startswith('//## File'):
        for f,ln in re.findall(r'"([^"]+)", line (\d+)', l):
            pending.append((f.split('/')[-1], int(ln)))
        continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m:
        if pending: chain=pending; pending=[]
        seq.append((chain,m.group(2)))
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; ia=hdr.index('Instructions Executed'); it=hdr.index('Thread Instructions Executed')
data=rows[2:]
core=[(139,163,'strong'),(165,200,'normal'),(202,225,'decision'),(123,131,'decision'),(226,240,'chroma_filter'),(248,285,'vert_pack'),(286,287,'hor_pack'),(290,310,'chroma_pack'),(102,118,'params'),(81,81,'params'),(316,334,'params'),(343,352,'bs'),(394,443,'lumablock_glue'),(444,500,'chromablock_glue')]
kern=[(77,100,'producer'),(124,150,'producer'),(152,158,'prologue'),(160,171,'item_setup'),(172,177,'params'),(178,191,'bs'),(192,193,'wait'),(194,209,'lds_sts'),(210,237,'store_out'),(238,246,'recycle')]
def classify(ch):
    for f,ln in ch:   # innermost first
        if f=='deblock_core.cuh':
            for a,b,n in core:
                if a<=ln<=b: return n
        if f=='deblock_kernels.cu':
            for a,b,n in kern:
                if a<=ln<=b: return n
        if f=='deblock_kernels.cuh': return 'producer'
    for f,ln in ch:
        if f=='tma_ptx.cuh': return 'tma_ptx'
    return 'other'
agg=collections.defaultdict(lambda:[0,0]); tot=0
for (ch,txt),r in zip(seq,data):
    a=int(r[ia]); t=int(r[it]); tot+=a
    k=classify(ch); agg[k][0]+=a; agg[k][1]+=t
for k,(a,t) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:16s} {a/tot*100:5.1f}%  lanes {t/max(a,1):4.1f}")
