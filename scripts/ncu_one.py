"""One kernel, a few launches, for `ncu --set full` captures:  ncu_one.py <what>  with <what> in
halo2_l1 (fprop 64->64 @32x32, CTA-pair halo) | conv2_l1 (same layer, tap-table pair) | dgrad_l2s2 (strided dgrad of 64->128/s2) |
wgrad_l1 | wgrad_l4 | conv2_l3 (256->256 @8x8)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops  # noqa: E402

ext = ops.require()
what = sys.argv[1]
NB = int(os.environ.get("NB", 1024))
SH = {"halo2_l1": (64, 32, 64, 3, 1, 1), "conv2_l1": (64, 32, 64, 3, 1, 1), "dgrad_l2s2": (64, 32, 128, 3, 2, 1), "wgrad_l1": (64, 32, 64, 3, 1, 1),
      "wgrad_l4": (512, 4, 512, 3, 1, 1), "conv2_l3": (256, 8, 256, 3, 1, 1), "halo2_l2": (128, 16, 128, 3, 1, 1)}[what]
Cin, H, Cout, R, st, pad = SH
x = (torch.randn(NB, Cin, H, H, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(Cout, Cin, R, R, device="cuda") / math.sqrt(Cin * R * R)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
Ho = (H + 2 * pad - R) // st + 1
dy = (torch.randn(NB, Cout, Ho, Ho, device="cuda") * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
scratch = torch.zeros(Cout * R * R * Cin, device="cuda")
wt = ext.weight_transpose(w, False)
for _ in range(4):
    if what.startswith("halo2"):
        ext.conv_fprop(x, w, None, st, pad, False)
    elif what.startswith("conv2"):
        ext.conv2(x, w, None, st, pad, False)
    elif what.startswith("dgrad"):
        ext.conv2(dy, wt, None, st, pad, False, None, True, H, H)
    else:
        ext.conv_wgrad(dy, x, scratch, R, R, st, pad)
torch.cuda.synchronize()
print(what, "tcerr", ext.tc_error())
