"""CTA-pair (cta_group::2) GEMM bring-up: correctness on a ladder of shapes (with a diagnosis of which output quadrants are wrong), then
throughput vs the 1-CTA kernel and cuBLAS.  Every launch checks the kernel's timeout word; run under `timeout`."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200 import ops
ext = ops.require()
torch.manual_seed(0)

def check(M, N, K):
    x = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    ref = (x.float() @ w.float().t())
    ext.tc_set_gemm_bn(2)
    y = ext.gemm_tn(x, w, None, False).float()
    torch.cuda.synchronize()
    e = ext.tc_error()
    err = (y - ref).abs()
    out = {"M": M, "N": N, "K": K, "tc_error": e, "max_err": err.max().item(), "ref_max": ref.abs().max().item()}
    if err.max().item() > 0.05:
        q = {}
        for i in range(0, min(M, 256), 128):
            for j in range(0, min(N, 256), 128):
                q["r%d_c%d" % (i, j)] = round(err[i:i + 128, j:j + 128].max().item(), 3)
        out["quadrant_err"] = q
        # is a column-half swap the explanation?
        if N >= 256:
            ysw = torch.cat([y[:, 128:256], y[:, 0:128]], 1)
            out["err_if_col_halves_swapped"] = (ysw[:256] - ref[:256, :256]).abs().max().item()
    print(json.dumps(out), flush=True)
    return e == 0 and err.max().item() <= 0.05

def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3

ok = True
for shp in [(256, 256, 64), (256, 256, 512), (512, 768, 256), (1000, 520, 328), (4096, 4096, 1024)]:
    ok = check(*shp) and ok
    if not ok: break
rows = []
if ok:
    for (M, N, K) in [(8192, 8192, 8192), (16384, 4096, 4096), (65536, 512, 2304)]:
        x = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
        fl = 2.0 * M * N * K
        r = {"M": M, "N": N, "K": K, "cublas_us": timeit(lambda: torch.nn.functional.linear(x, w))}
        for name, bn in (("bn256", 256), ("pair", 2)):
            ext.tc_set_gemm_bn(bn)
            r[name + "_us"] = timeit(lambda: ext.gemm_tn(x, w, None, False))
            r[name + "_pf"] = fl / r[name + "_us"] / 1e9
        r["cublas_pf"] = fl / r["cublas_us"] / 1e9
        r["tc_error"] = ext.tc_error()
        rows.append(r); print(json.dumps(r), flush=True)
ext.tc_set_gemm_bn(-1)
os.makedirs("gpurun_out", exist_ok=True); json.dump({"ok": ok, "rows": rows}, open("gpurun_out/gemm2.json", "w"), indent=1)
