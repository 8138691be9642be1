import torch, time, numpy as np
dev=torch.device('cuda',0)
n=4<<30
cudart=torch.cuda.cudart()
def pinned(nbytes):
    a=np.zeros(nbytes,dtype=np.uint8); rc=cudart.cudaHostRegister(a.ctypes.data,nbytes,0); assert int(rc)==0; return torch.from_numpy(a)
h1=pinned(n); h2=pinned(n)
d1=torch.empty(n,dtype=torch.uint8,device=dev); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(fn, label, nbytes):
    torch.cuda.synchronize(); t0=time.perf_counter(); fn(); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(label, '%.1f GB/s'%(nbytes/dt/1e9), flush=True)
for _ in range(2):
    t(lambda: d1.copy_(h1,non_blocking=True), 'H2D alone', n)
    t(lambda: h2.copy_(d2,non_blocking=True), 'D2H alone', n)
def both():
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
t(both,'H2D+D2H concurrent (sum)', 2*n)
hp=torch.empty(n,dtype=torch.uint8,pin_memory=True)
t(lambda: hp.copy_(d2,non_blocking=True), 'D2H alone (torch pinned alloc)', n)
