"""Single-shape conv launcher for ncu captures and halo-descriptor experiments.
usage: conv_one.py <algo> [desc_mode]  (shape from env SHAPE="Cin,H,W,Cout,R,stride,pad", NB)"""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops
ext = ops.require()
algo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ext.tc_set_conv_algo(algo)
if len(sys.argv) > 2:
    ext.tc_set_halo_desc_mode(int(sys.argv[2]))
Cin, H, W, Cout, R, st, pad = [int(v) for v in os.environ.get("SHAPE", "64,32,32,64,3,1,1").split(",")]
NB = int(os.environ.get("NB", 1024))
x = (torch.randn(NB, Cin, H, W, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(Cout, Cin, R, R, device="cuda") / math.sqrt(Cin * R * R)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
ref = F.conv2d(x[:8].float(), w.float(), None, st, pad)
for _ in range(3):
    y = ext.conv_fprop(x, w, None, st, pad, False)
torch.cuda.synchronize()
d = (y[:8].float() - ref).abs()
print("algo", algo, "mode", sys.argv[2:] , "maxerr", d.max().item(), "tcerr", ext.tc_error())
if d.max().item() > 0.1:
    # which taps are wrong?  probe with one-hot filters
    for r in range(R):
        for s in range(R):
            w1 = torch.zeros_like(w); w1[:, :, r, s] = w[:, :, r, s]
            y1 = ext.conv_fprop(x[:8].contiguous(memory_format=torch.channels_last), w1.contiguous(memory_format=torch.channels_last), None, st, pad, False)
            e = (y1.float() - F.conv2d(x[:8].float(), w1.float(), None, st, pad)).abs().max().item()
            print("  tap", r, s, "err", round(e, 4))
