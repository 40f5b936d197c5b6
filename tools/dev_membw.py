import torch
x=torch.empty(4*1024**3//2, dtype=torch.bfloat16, device='cuda')
y=torch.empty_like(x)
def t(fn,n=5):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
nb=x.numel()*2
ms=t(lambda: x.zero_()); print('memset  %.2f ms  %.2f TB/s (write only)'%(ms, nb/ms/1e9))
ms=t(lambda: y.copy_(x)); print('copy    %.2f ms  %.2f TB/s (read+write)'%(ms, 2*nb/ms/1e9))
ms=t(lambda: x.sum()); print('reduce  %.2f ms  %.2f TB/s (read only)'%(ms, nb/ms/1e9))
