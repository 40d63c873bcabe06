import torch, time
x = torch.randn(622607, 4, device="cuda")
idx = torch.arange(100000, 255000, device="cuda")
def t(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
print("index_select int64 :", t(lambda: x.index_select(0, idx)), "us")
print("x[idx] int64       :", t(lambda: x[idx]), "us")
i32 = idx.int()
print("index_select int32 :", t(lambda: x.index_select(0, i32)), "us")
print("x[idx] int32       :", t(lambda: x[i32]), "us")
xv = x.view(torch.complex128).reshape(-1) if False else None
print("take_along / gather:", t(lambda: torch.gather(x, 0, idx[:, None].expand(-1, 4))), "us")
