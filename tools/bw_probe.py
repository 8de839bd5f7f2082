import torch
n = 1<<29  # 512M bf16 = 1 GB
a = torch.empty(n, dtype=torch.bfloat16, device='cuda'); b = torch.randn(n, device='cuda', dtype=torch.bfloat16); c = torch.randn(n, device='cuda', dtype=torch.bfloat16)
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
def t(fn, nbytes, name):
    ts=[]
    for _ in range(5):
        flush.zero_()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms=sorted(ts)[2]; print(name, '%.3f ms %.0f GB/s' % (ms, nbytes/ms/1e6))
t(lambda: a.copy_(b), 2*n*2, 'copy 1R1W')
t(lambda: torch.add(b, c, out=a), 3*n*2, 'add 2R1W')
t(lambda: torch.add(b, c, out=b), 3*n*2, 'add inplace 2R1W')
t(lambda: (b*c).sum(), 2*n*2, 'mul-sum 2R (+tmp)')
t(lambda: torch.dot(b, c), 2*n*2, 'dot 2R')
