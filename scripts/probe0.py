"""First-contact probe on a B200 box: stock torch ResNet-18/CIFAR bf16 step time (eager + CUDA graph),
device attributes relevant to the symmetric arena (VMM / multicast / fabric), P2P facts."""
import json, os, subprocess, sys, time
import torch, torch.nn as nn, torch.nn.functional as F
out = {}
dev = torch.device("cuda:0")
p = torch.cuda.get_device_properties(0)
out["name"] = p.name; out["sms"] = p.multi_processor_count; out["mem_gb"] = p.total_memory / 2**30
out["ngpu"] = torch.cuda.device_count()
try:
    from cuda.bindings import driver as cu
    cu.cuInit(0)
    for nm in ["CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED",
               "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED",
               "CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_WITH_CUDA_VMM_SUPPORTED"]:
        err, v = cu.cuDeviceGetAttribute(getattr(cu.CUdevice_attribute, nm), 0)
        out[nm] = int(v)
except Exception as e:
    out["cuda_python_err"] = repr(e)
out["topo"] = subprocess.run("nvidia-smi topo -m | head -14", shell=True, capture_output=True, text=True).stdout
out["smi"] = subprocess.run("nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv", shell=True, capture_output=True, text=True).stdout

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

class BasicBlock(nn.Module):
    def __init__(s, i, o, st):
        super().__init__()
        s.c1 = nn.Conv2d(i, o, 3, st, 1, bias=False); s.b1 = nn.BatchNorm2d(o)
        s.c2 = nn.Conv2d(o, o, 3, 1, 1, bias=False); s.b2 = nn.BatchNorm2d(o)
        s.sh = nn.Sequential()
        if st != 1 or i != o:
            s.sh = nn.Sequential(nn.Conv2d(i, o, 1, st, bias=False), nn.BatchNorm2d(o))
    def forward(s, x):
        y = F.relu(s.b1(s.c1(x))); y = s.b2(s.c2(y)); return F.relu(y + s.sh(x))
class R18(nn.Module):
    def __init__(s):
        super().__init__()
        s.c = nn.Conv2d(3, 64, 3, 1, 1, bias=False); s.b = nn.BatchNorm2d(64)
        L = []; i = 64
        for o, st in [(64,1),(64,1),(128,2),(128,1),(256,2),(256,1),(512,2),(512,1)]:
            L.append(BasicBlock(i, o, st)); i = o
        s.l = nn.Sequential(*L); s.fc = nn.Linear(512, 10)
    def forward(s, x):
        x = s.l(F.relu(s.b(s.c(x)))); x = F.avg_pool2d(x, 4).flatten(1); return s.fc(x)

def timeit(fn, warm=5, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

B = 1024
for tag, dt, cl in [("bf16_cl", torch.bfloat16, True), ("bf16_nchw", torch.bfloat16, False), ("fp32_tf32", torch.float32, True)]:
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    m = R18().to(dev).to(dt)
    if cl: m = m.to(memory_format=torch.channels_last)
    x = torch.randn(B, 3, 32, 32, device=dev, dtype=dt); y = torch.randint(0, 10, (B,), device=dev)
    if cl: x = x.contiguous(memory_format=torch.channels_last)
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
    def step():
        opt.zero_grad(set_to_none=False)
        loss = F.cross_entropy(m(x).float(), y); loss.backward(); opt.step()
    def fwd():
        with torch.no_grad(): m(x)
    out[f"eager_step_ms_{tag}"] = timeit(step)
    out[f"eager_fwd_ms_{tag}"] = timeit(fwd)
    t0 = time.time(); 
    for _ in range(20): step()
    torch.cuda.synchronize(); out[f"eager_wall_ms_{tag}"] = (time.time() - t0) / 20 * 1e3
    # whole-step CUDA graph
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=False)
        with torch.cuda.graph(g):
            loss = F.cross_entropy(m(x).float(), y); loss.backward(); opt.step()
        out[f"graph_step_ms_{tag}"] = timeit(g.replay)
    except Exception as e:
        out[f"graph_err_{tag}"] = repr(e)[:300]
    del m, opt
    torch.cuda.empty_cache()
# how many kernels per eager step
from torch.profiler import profile, ProfilerActivity
m = R18().to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
x = torch.randn(B, 3, 32, 32, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last); y = torch.randint(0, 10, (B,), device=dev)
for _ in range(3):
    F.cross_entropy(m(x).float(), y).backward()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    F.cross_entropy(m(x).float(), y).backward(); torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(k.key, k.count, getattr(k, "device_time_total", 0.0)) for k in ka], key=lambda r: -r[2])[:40]
out["top_kernels"] = [(a[:110], b, round(c, 1)) for a, b, c in rows]
out["n_kernels"] = sum(k.count for k in ka)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe0.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k not in ("top_kernels", "topo")}, indent=1))
print(out["topo"])
for r in out["top_kernels"]: print(r)
