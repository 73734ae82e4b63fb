import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from exploring_flash_attention_b200 import ops
from oracle import reference
for (B,H,L,d,dt) in ((1,2,700,32,torch.float32),(1,2,129,64,torch.float32),(1,2,256,32,torch.float32),(1,2,700,32,torch.float16)):
    g=torch.Generator().manual_seed(42)
    Q,K,V=(((torch.rand((B,H,L,d),generator=g)*2-1)).to(dt).cuda() for _ in range(3))
    for causal in (False,True):
        O,lse=ops.flash_attention_v1_ex(Q,K,V,causal=causal,return_lse=True,sync=True)
        q,k,v=(x.float().cpu().numpy().reshape(-1,L,d) for x in (Q,K,V))
        ro,rl=reference.naive_attention_ex_f64(q[0],k[0],v[0],causal=causal)
        E=np.abs(O.float().cpu().numpy().reshape(-1,L,d)[0]-ro)
        El=np.abs(lse.cpu().numpy().reshape(-1,L)[0]-rl)
        rows=np.argsort(-E.max(1))[:5]
        print(dt,L,d,'causal',causal,'Oerr',E.max(),'rows',rows,E.max(1)[rows],'lse err',El.max(), 'at', El.argmax())
