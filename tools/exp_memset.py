import torch, time
B,A,N=65536,7,200
mask=torch.zeros(B,A,N,dtype=torch.bool,device='cuda'); nf=torch.zeros(B,N,A,device='cuda')
for _ in range(5): mask.zero_(); nf.zero_()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): mask.zero_(); nf.zero_()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/50
print("memset mask+nf ms", ms, "GB/s", (mask.numel()+nf.numel()*4)/ms/1e6)
src=torch.empty(512*1024*1024//4,device='cuda'); dst=torch.empty_like(src)
for _ in range(3): dst.copy_(src)
e0.record()
for _ in range(20): dst.copy_(src)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/20
print("copy 512MB ms", ms, "GB/s r+w", 2*src.numel()*4/ms/1e6)
