"""Pure-write HBM bandwidth on this box (calibration for the K2 roofline): torch zero_() / fill_() / copy_."""
import torch
dev = torch.device("cuda")
for mb in (246, 1024, 4096):
    x = torch.empty(mb * (1 << 20) // 8, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    for name, fn, byt in (("zero_", lambda: x.zero_(), 1), ("fill_", lambda: x.fill_(1.5), 1), ("copy_", lambda: y.copy_(x), 2)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(10):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name:6s} {mb:5d} MiB: {best*1e3:8.1f} us  {byt * x.numel() * 8 / best / 1e6:8.1f} GB/s")
