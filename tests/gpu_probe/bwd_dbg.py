import sys
sys.path.insert(0, "/root/repo")
import torch
from exploring_flash_attention_b200 import ops
for (B,H,L,d) in ((1,1,128,128),(1,1,128,64),(1,1,256,128)):
    q,k,v,do = ((torch.rand((B,H,L,d), device="cuda")*2-1).bfloat16() for _ in range(4))
    o,lse = ops.flash_attention_v1_ex(q,k,v,return_lse=True, sync=True)
    try:
        r = ops.flash_attention_backward(q,k,v,o,do,lse, sync=True)
        print("ok", B,H,L,d, float(r[0].float().abs().mean()))
    except Exception as e:
        print("FAIL", B,H,L,d, str(e)[:300]); break
