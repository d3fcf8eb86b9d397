import torch, time
dev=torch.device('cuda',0)
h=torch.empty(176<<20,dtype=torch.uint8).pin_memory(); d=torch.empty_like(h,device=dev)
hb=torch.empty(31<<20,dtype=torch.uint8).pin_memory(); db=torch.empty_like(hb,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(both, chunks=1):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10):
        n=h.numel()//chunks; nb=hb.numel()//chunks
        for c in range(chunks):
            with torch.cuda.stream(s1): d[c*n:(c+1)*n].copy_(h[c*n:(c+1)*n],non_blocking=True)
            if both:
                with torch.cuda.stream(s2): hb[c*nb:(c+1)*nb].copy_(db[c*nb:(c+1)*nb],non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/10*1e3
for _ in range(2): run(True)
print('h2d only %.3f ms'%run(False)); print('h2d+d2h concurrent %.3f ms'%run(True)); print('8 chunks: h2d only %.3f, both %.3f'%(run(False,8),run(True,8)))
