"""Diagnostic: what a plain read-only torch reduction reaches on this GPU (context for the class-scan roofline)."""
import torch
dev = torch.device("cuda:0")
x = torch.randn(64, 80, 8400, device=dev)           # the class rows of one C2 batch: 172 MB
big = torch.randn(3, 64, 144, 8400, device=dev)      # 3 x 310 MB
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
i = [0]
def amax_rows():
    i[0] += 1
    return big[i[0] % 3][:, 64:].amax(1)
def amax_all():
    i[0] += 1
    return big[i[0] % 3].amax()
def copy():
    i[0] += 1
    return big[i[0] % 3].clone()
for name, fn, nbytes in (("amax over classes (strided rows)", amax_rows, 64*80*8400*4), ("amax over everything", amax_all, 64*144*8400*4), ("clone", copy, 2*64*144*8400*4)):
    ms = timeit(fn)
    print(f"{name}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s")
