"""tcgen05 GEMM throughput vs N tile (128 vs 256 columns per MMA) vs cuBLAS, large square-ish problems, CUDA events."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops
ext = ops.require()
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
rows = []
for (M, N, K) in [(8192, 8192, 8192), (16384, 4096, 4096), (65536, 512, 2304), (1048576, 64, 576)]:
    x = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    fl = 2.0 * M * N * K
    r = {"M": M, "N": N, "K": K}
    r["cublas_us"] = timeit(lambda: torch.nn.functional.linear(x, w))
    ref = torch.nn.functional.linear(x[:256], w).float()
    for bn in (128, 256):
        if bn == 256 and N < 256: continue
        ext.tc_set_gemm_bn(bn)
        y = ext.gemm_tn(x, w, None, False)
        r["bn%d_err" % bn] = (y[:256].float() - ref).abs().max().item()
        r["bn%d_us" % bn] = timeit(lambda: ext.gemm_tn(x, w, None, False))
        r["bn%d_pf" % bn] = fl / r["bn%d_us" % bn] / 1e9
    ext.tc_set_gemm_bn(-1)
    r["cublas_pf"] = fl / r["cublas_us"] / 1e9
    r["tcerr"] = ext.tc_error()
    rows.append(r); print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True); json.dump(rows, open("gpurun_out/bench_gemm.json", "w"), indent=1)
