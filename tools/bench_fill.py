import torch
n=401*401*601
a=torch.empty(n,dtype=torch.float64,device='cuda'); b=torch.empty(n,dtype=torch.float64,device='cuda')
for name,fn in (("zero_",lambda t:t.zero_()),("fill_",lambda t:t.fill_(1.5)),("copy",lambda t:t.copy_(a if t is b else b))):
    for _ in range(3): fn(a); fn(b)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8): fn(a if i&1 else b)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/8
    print(name, ms, 'ms', n*8/ms/1e6*(2 if name=="copy" else 1),'GB/s')
