"""Stem wgrad (3->64, 3x3, 32x32, batch 1024): dedicated kernel vs cuDNN.  CUDA events, 20 iterations after 5 warm-up; dY (134 MB) exceeds L2."""
import torch, torch.nn.functional as F
from ps_pytorch_b200 import ops
ext = ops.require()
torch.backends.cudnn.benchmark = True
x = torch.randn(1024, 3, 32, 32, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
dy = torch.randn(1024, 64, 32, 32, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = torch.randn(64, 3, 3, 3, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
sc = torch.zeros(64 * 27, device="cuda")
def t(f, n=20):
    for _ in range(5): f()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
ours = t(lambda: ext.stem_wgrad(dy, x, sc))
lib = t(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]))
gb = (dy.numel() + x.numel()) * 2 / 1e9
print(f"stem wgrad: ours {ours:.1f} us ({gb / ours * 1e6 / 1e3:.2f} TB/s of dY+x) | cuDNN {lib:.1f} us")
