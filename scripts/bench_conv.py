"""Per-layer convolution microbenchmark on one B200: tcgen05 per-tap vs halo kernels vs cuDNN, ResNet-18 shapes at batch 1024.
CUDA-event timing, 3 warm-up + 10 timed launches, a 256 MB buffer is written between launches to flush L2."""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops  # noqa: E402

ext = ops.require()
torch.backends.cudnn.benchmark = True
NB = int(os.environ.get("NB", 1024))
SHAPES = [  # Cin, H, W, Cout, R, stride, pad
    (64, 32, 32, 64, 3, 1, 1), (64, 32, 32, 128, 3, 2, 1), (128, 16, 16, 128, 3, 1, 1), (128, 16, 16, 256, 3, 2, 1),
    (256, 8, 8, 256, 3, 1, 1), (256, 8, 8, 512, 3, 2, 1), (512, 4, 4, 512, 3, 1, 1), (64, 32, 32, 128, 1, 2, 0)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(it):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / it * 1e3


rows = []
for (Cin, H, W, Cout, R, st, pad) in SHAPES:
    x = (torch.randn(NB, Cin, H, W, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, R, R, device="cuda") / math.sqrt(Cin * R * R)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    Ho = (H + 2 * pad - R) // st + 1
    flops = 2.0 * NB * Ho * Ho * Cout * Cin * R * R
    ref = F.conv2d(x, w, None, st, pad)
    res = {"shape": [Cin, H, W, Cout, R, st, pad], "gflop": flops / 1e9}
    res["cudnn_us"] = timeit(lambda: F.conv2d(x, w, None, st, pad))
    for name, algo, cl in (("tap", 0, 1), ("halo", 1, 1), ("halo2", 1, 2)):
        ext.tc_set_conv_algo(algo)
        ext.tc_set_halo_cluster(cl)
        try:
            y = ext.conv_fprop(x, w, None, st, pad, False)
            torch.cuda.synchronize()
            err = (y.float() - ref.float()).abs().max().item()
            res[name + "_us"] = timeit(lambda: ext.conv_fprop(x, w, None, st, pad, False))
            res[name + "_maxerr"] = err
            res[name + "_tcerr"] = ext.tc_error()
        except Exception as e:  # noqa: BLE001
            res[name + "_err"] = repr(e)[:200]
    ext.tc_set_conv_algo(-1)
    # weight gradient: cuDNN vs ours
    try:
        dyt = (torch.randn_like(ref) * 0.1).contiguous(memory_format=torch.channels_last)
        res["cudnn_wgrad_us"] = timeit(lambda: torch.ops.aten.convolution_backward(dyt, x, w, None, (st, st), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False)))
        if Cin % 64 == 0 and Cout % 64 == 0:
            scratch = torch.zeros(Cout * R * R * Cin, device="cuda")
            g0 = torch.ops.aten.convolution_backward(dyt, x, w, None, (st, st), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False))[1]
            g1 = ext.conv_wgrad(dyt, x, scratch, R, R, st, pad)
            res["wgrad_maxerr"] = (g1.float() - g0.float()).abs().max().item()
            res["wgrad_us"] = timeit(lambda: ext.conv_wgrad(dyt, x, scratch, R, R, st, pad))
    except Exception as e:  # noqa: BLE001
        res["wgrad_err"] = repr(e)[:200]
    for k in ("cudnn_us", "tap_us", "halo_us", "halo2_us"):
        if k in res:
            res[k.replace("_us", "_pflops")] = flops / res[k] / 1e9
    rows.append(res)
    print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_conv.json", "w"), indent=1)
