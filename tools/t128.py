import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, os.path.join(R,'tests')]
import numpy as np, torch, time
import entropy_coders_b200 as E, oracle_lib as O
ctx=E.Context(0)
for kind,n,bs,tl in [("text",9*65536+4321,65536,0),("geo",5*131072+99,131072,0),("few",6*65536,65536,9),("uniform",4*65536,65536,12),("text",128*7+5,300,0),("text", 3000, 1000, 13)]:
    src=O.generate(kind,77,n)
    scratch,sizes,status=O.compress_blocks(src,bs,tl,128,threads=8)
    nb=len(sizes)
    off=np.zeros(nb+1,np.int64); off[1:]=np.cumsum(sizes)
    comp=np.concatenate([scratch[i,:int(s)] for i,s in enumerate(sizes)])
    ok_blocks=[i for i in range(nb) if status[i]==0]
    out,st=ctx.decompress_blocks(torch.from_numpy(comp).cuda(),comp.size,torch.from_numpy(off).cuda(),n,bs,tl,128)
    st=st.cpu().numpy(); out=out.cpu().numpy()
    good=all(np.array_equal(out[i*bs:(i+1)*bs],src[i*bs:(i+1)*bs]) and st[i]==0 for i in ok_blocks)
    print(kind,n,bs,tl,"blocks ok:",len(ok_blocks),"/",nb,"decode128", "OK" if good else "MISMATCH", st[:6])
