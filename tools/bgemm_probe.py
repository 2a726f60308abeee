"""Diagnostic (not part of the product): accuracy (vs an fp64 einsum) and time of ops.bgemm256 - the per-sample tcgen05
GEMM of the loss path - at a few (B, D) shapes."""
import os
import sys,torch,time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import ops
torch.manual_seed(0)
for (B,D) in ((1,1),(2,130),(3,784),(5,3072),(64,3072),(80,300)):
    X=torch.rand(B,D,256,device='cuda')*torch.exp(4*torch.randn(B,D,1,device='cuda'))
    M=torch.rand(B,256,256,device='cuda')
    out=ops.bgemm256(X,M)
    ref=torch.einsum('bdk,bnk->bdn',X.double(),M.double())
    rel=((out.double()-ref).abs()/ref.abs().clamp_min(1e-30)).max().item()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.bgemm256(X,M,out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(B,D,'max rel err',rel,'ms',ms,'TFLOP/s',2*B*D*65536/ms/1e9)
    assert rel<1e-4
