import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops
ext = ops.require()
M = N = K = 8192
x = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
for bn in (2, 256):
    ext.tc_set_gemm_bn(bn)
    for _ in range(3): ext.gemm_tn(x, w, None, False)
torch.cuda.synchronize()
