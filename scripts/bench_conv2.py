"""Per-layer convolution benchmark, all three passes, ResNet-18 shapes at batch 1024 on one B200: our kernels (halo pair / tap-table pair /
1-CTA per-tap / wgrad) vs cuDNN.  CUDA events, 3 warm-up + 10 timed launches, 256 MB written between launches to flush L2."""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops  # noqa: E402
from ps_pytorch_b200.ops import conv as C  # noqa: E402

ext = ops.require()
torch.backends.cudnn.benchmark = True
NB = int(os.environ.get("NB", 1024))
SHAPES = [  # Cin, H, W, Cout, R, stride, pad
    (64, 32, 32, 64, 3, 1, 1), (64, 32, 32, 128, 3, 2, 1), (64, 32, 32, 128, 1, 2, 0), (128, 16, 16, 128, 3, 1, 1), (128, 16, 16, 256, 3, 2, 1),
    (128, 16, 16, 256, 1, 2, 0), (256, 8, 8, 256, 3, 1, 1), (256, 8, 8, 512, 3, 2, 1), (256, 8, 8, 512, 1, 2, 0), (512, 4, 4, 512, 3, 1, 1)]
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


def attempt(res, name, fn, ref=None):
    try:
        y = fn()
        torch.cuda.synchronize()
        if ref is not None:
            res[name + "_maxerr"] = (y.float() - ref.float()).abs().max().item()
        res[name + "_us"] = timeit(fn)
        e = ext.tc_error()
        if e:
            res[name + "_tcerr"] = e
    except Exception as e:  # noqa: BLE001
        res[name + "_err"] = repr(e)[:200]


rows = []
for (Cin, H, W, Cout, R, st, pad) in SHAPES:
    x = (torch.randn(NB, Cin, H, W, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, R, R, device="cuda") / math.sqrt(Cin * R * R)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    Ho = (H + 2 * pad - R) // st + 1
    flops = 2.0 * NB * Ho * Ho * Cout * Cin * R * R
    ref = F.conv2d(x, w, None, st, pad)
    dy = (torch.randn_like(ref) * 0.1).contiguous(memory_format=torch.channels_last)
    res = {"shape": [Cin, H, W, Cout, R, st, pad], "gflop": flops / 1e9}
    # ---- forward
    res["f_cudnn_us"] = timeit(lambda: F.conv2d(x, w, None, st, pad))
    attempt(res, "f_pair", lambda: ext.conv2(x, w, None, st, pad, False), ref)
    if Cout >= 128:
        ext.tc_set_conv2_bn(128)
        attempt(res, "f_pair_bn128", lambda: ext.conv2(x, w, None, st, pad, False), ref)
        ext.tc_set_conv2_bn(-1)
    ext.tc_set_conv_algo(0)
    attempt(res, "f_tap1", lambda: ext.conv_fprop(x, w, None, st, pad, False), ref)
    ext.tc_set_conv_algo(-1)
    if C._halo_ok(Ho, Ho, R, R, st):
        attempt(res, "f_halo2", lambda: ext.conv_fprop(x, w, None, st, pad, False), ref)
        ext.tc_set_halo_pair(0)
        attempt(res, "f_halo1", lambda: ext.conv_fprop(x, w, None, st, pad, False), ref)
        ext.tc_set_halo_pair(1)
    # ---- data gradient
    gref = torch.ops.aten.convolution_backward(dy, x, w, None, (st, st), (pad, pad), (1, 1), False, (0, 0), 1, (True, False, False))[0]
    res["d_cudnn_us"] = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (st, st), (pad, pad), (1, 1), False, (0, 0), 1, (True, False, False)))
    attempt(res, "d_ours", lambda: C.dgrad(ext, dy, w, H, W, st, pad), gref)
    wt = ext.weight_transpose(w, False)
    attempt(res, "d_pair_kernel_only", lambda: ext.conv2(dy, wt, None, st, pad, False, None, True, H, W), gref)
    # ---- weight gradient
    g0 = torch.ops.aten.convolution_backward(dy, x, w, None, (st, st), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False))[1]
    res["w_cudnn_us"] = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (st, st), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False)))
    scratch = torch.zeros(Cout * R * R * Cin, device="cuda")
    attempt(res, "w_ours", lambda: ext.conv_wgrad(dy, x, scratch, R, R, st, pad), g0)
    for nm, al in (("w_tap", 0), ("w_halo", 1), ("w_pair", 2)):
        ext.tc_set_wgrad_algo(al)
        attempt(res, nm, lambda: ext.conv_wgrad(dy, x, scratch, R, R, st, pad), g0)
    ext.tc_set_wgrad_algo(-1)
    for k in list(res):
        if k.endswith("_us"):
            res[k.replace("_us", "_pflops")] = round(flops / res[k] / 1e9, 3)
    rows.append(res)
    print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_conv2.json", "w"), indent=1)
hdr = ["f_cudnn", "f_pair", "f_pair_bn128", "f_tap1", "f_halo2", "f_halo1", "d_cudnn", "d_ours", "d_pair_kernel_only", "w_cudnn", "w_ours", "w_tap", "w_halo", "w_pair"]
print("| shape | " + " | ".join(hdr) + " |")
for r in rows:
    print("| %s | " % "x".join(map(str, r["shape"])) + " | ".join(("%.1f" % r[h + "_us"]) if h + "_us" in r else (r.get(h + "_err", "-")[:20]) for h in hdr) + " |")
