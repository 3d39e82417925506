import sys, os, lzma, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import kf_oracle as o
from kf2vecfsw_b200 import engine
engine.init(0)
d='/root/repo/tests/golden/fna'
small=lzma.decompress(open(os.path.join(d,'G000830275sub.fna.xz'),'rb').read())
big=lzma.decompress(open(os.path.join(d,'G000830295.fna.xz'),'rb').read())
ref=o.canonical_counts_bytes(small,7)
for tag,bufs,kw in (("small alone",[small],{}),("small alone nolg",[small],{"no_linegrid":True}),("small x2",[small,small],{}),("big+small",[big,small],{}),("small alone again",[small],{})):
    c,f,t,st=engine.count_buffers(bufs,k=7,**kw)
    i=len(bufs)-1
    print(tag, int(t[i]), int(ref.sum()), np.array_equal(c[i],ref), st.tolist(), flush=True)
# a few synthetic sizes
import random
random.seed(1)
for L in (800, 8000, 80000, 800000):
    seq=''.join(random.choice('ACGT') for _ in range(L))
    data=('>h\n'+'\n'.join(seq[i:i+80] for i in range(0,L,80))+'\n').encode()
    c,f,t,st=engine.count_buffers([data],k=7)
    r=o.canonical_counts_bytes(data,7)
    print("synthetic",L,int(t[0]),int(r.sum()),np.array_equal(c[0],r), flush=True)
