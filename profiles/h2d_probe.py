import torch, time
x = torch.empty(66_500_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in (1, 4):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): d.copy_(x, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('H2D GB/s', 8*x.numel()/e0.elapsed_time(e1)/1e6)
