import torch,time
x=torch.empty(256*1024*1024,dtype=torch.float32).pin_memory()
d=torch.empty_like(x,device='cuda')
for _ in range(3): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("H2D GB/s", x.numel()*4/dt/1e9, "ms per GiB", dt*1e3)
