import torch, time
x = torch.zeros(1024, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(100): x.add_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(s)
    for _ in range(2000): x.add_(1)
    e1.record(s); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("eager tiny kernels: %.2f us/launch on device, %.2f us/launch CPU issue" % (e0.elapsed_time(e1) * 1e3 / 2000, t_issue * 1e6 / 2000))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(2000): x.add_(1)
    g.replay(); torch.cuda.synchronize()
    e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
    print("graph tiny kernels: %.2f us/kernel" % (e0.elapsed_time(e1) * 1e3 / 2000))
    big = torch.zeros(64 << 20, device="cuda", dtype=torch.uint8)
    e0.record(s)
    for _ in range(100): big.zero_()
    e1.record(s); torch.cuda.synchronize()
    print("64MB memset: %.2f us each" % (e0.elapsed_time(e1) * 1e3 / 100))
