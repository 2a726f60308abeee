import sys,time,torch
sys.path.insert(0,'/root/repo')
B,D=1024,3072
dev=torch.device('cuda')
hh=torch.empty((B,2*D),dtype=torch.float32,pin_memory=True); hx=torch.empty((B,D),dtype=torch.int32,pin_memory=True); ho=torch.empty((B,D),dtype=torch.int32,pin_memory=True)
dh=torch.empty((B,2*D),device=dev); dx=torch.empty((B,D),dtype=torch.int32,device=dev)
s1=torch.cuda.Stream()
for name,fn in (("h2d head",lambda: dh.copy_(hh,non_blocking=True)),("h2d x",lambda: dx.copy_(hx,non_blocking=True)),("d2h x",lambda: ho.copy_(dx,non_blocking=True))):
    for st in (None,s1):
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(10):
            if st is None: fn()
            else:
                with torch.cuda.stream(st): fn()
        torch.cuda.synchronize(); print(name,'stream' if st else 'default',(time.perf_counter()-t0)*100,'ms each')
