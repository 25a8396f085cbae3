import torch, time
x = torch.empty(2*1024**3, dtype=torch.float32, device='cuda')  # 8 GiB
x.normal_()
for name, fn in [("sum", lambda: x.sum()), ("max", lambda: x.max()), ("copy_half", lambda: x[:1024**3].copy_(x[1024**3:]))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    nbytes = x.numel()*4 if name != "copy_half" else x.numel()*4
    print(name, "%.3f ms" % ms, "%.0f GB/s" % (nbytes/ms/1e6))
