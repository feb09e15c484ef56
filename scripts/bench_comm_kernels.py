"""Isolated timing of the two PS hot-path kernels on one GPU (local slots): push and reduce_update, CUDA events,
L2 flushed between launches.  Reports achieved HBM bytes/s against MEASURED_PEAKS.json (6.57 TB/s copy)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200.optim import build_optimizer
from ps_pytorch_b200.parallel.arena import FlatBuffer, ParamLayout
from ps_pytorch_b200.parallel.nvl import NvlTransport
PEAK = 6570.9e9
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
rows = []
for T in (11173962 // 128 * 128 + 128, 128 * 1024 * 1024):
    for wire, opt in (("bf16", "sgd"), ("fp8", "sgd"), ("bf16", "adam")):
        lay = ParamLayout(["w"], [(T,)], [T], [0], [False], T)
        tr = NvlTransport(lay, torch.device("cuda"), grad_dtype=wire)
        master = FlatBuffer(lay, torch.float32, "cuda"); master.flat.normal_()
        o = build_optimizer(opt, [torch.nn.Parameter(master.flat)], lr=0.01, momentum=0.9)
        o.attach_flat(master)
        lowp = FlatBuffer(lay, torch.bfloat16, "cuda", storage=tr.ps_weight_storage()); lowp.flat.copy_(master.flat)
        grads = FlatBuffer(lay, torch.bfloat16, "cuda", storage=tr.worker_grad_storage(torch.bfloat16))
        tp = tr_ = 0.0
        n = 0
        for step in range(1, 9):
            tr.ps_publish(step, lowp); tr.worker_fetch(lowp)
            grads.flat.normal_()
            flush.fill_(1.0)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); tr.worker_push(step, grads); e[1].record()
            flush.fill_(2.0)
            e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e2[0].record(); tr.ps_reduce_update(step, 1, o, master, lowp=lowp); e2[1].record()
            torch.cuda.synchronize(); tr.check_error()
            if step > 3:
                tp += e[0].elapsed_time(e[1]); tr_ += e2[0].elapsed_time(e2[1]); n += 1
        wb = {"bf16": 2, "fp8": 1.03125}[wire]
        push_bytes = T * (2 + 2 + wb)                                    # read grad, zero grad, write slot
        ru_bytes = T * (wb + 4 + 4 + 4 + 4 + 2) if opt == "sgd" else T * (wb + 4 * 6 + 2)   # slot + w r/w + state r/w + bf16 replica
        r = {"T": T, "wire": wire, "opt": opt, "push_us": tp / n * 1e3, "reduce_update_us": tr_ / n * 1e3,
             "push_frac_hbm": push_bytes / (tp / n * 1e-3) / PEAK, "ru_frac_hbm": ru_bytes / (tr_ / n * 1e-3) / PEAK}
        rows.append(r); print(json.dumps(r), flush=True)
        del tr, master, o, lowp, grads
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_comm_kernels.json", "w"), indent=1)
