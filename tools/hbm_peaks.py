import torch, json
torch.cuda.init()
n = 1 << 30  # 4 GiB of f32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def best(fn, bytes_, reps=10):
    ts=[]
    for _ in range(3): fn()
    torch.cuda.synchronize()
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)*1e-3)
    return bytes_/min(ts)/1e9
print(json.dumps({
 "copy_GBps": best(lambda: y.copy_(x), 8*n),
 "memset_GBps": best(lambda: x.zero_(), 4*n),
 "fill_GBps": best(lambda: x.fill_(1.5), 4*n),
 "read_sum_GBps": best(lambda: x.sum(), 4*n),
 "add_inplace_GBps": best(lambda: x.add_(1.0), 8*n),
}))
