import torch, time
n=87*1024*1024
h=torch.empty(n,dtype=torch.int8).pin_memory()
d=torch.empty(n,dtype=torch.int8,device='cuda')
def run(k,reps=20):
    ss=[torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for r in range(reps):
        c=n//k
        for i,s in enumerate(ss):
            with torch.cuda.stream(s):
                d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c],non_blocking=True)
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/reps
    return n/dt/1e9
for k in (1,2,4,1,2):
    print(k, round(run(k),2),'GB/s')
# with concurrent D2H
h2=torch.empty(10*1024*1024,dtype=torch.int8).pin_memory(); d2=torch.empty(10*1024*1024,dtype=torch.int8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for r in range(20):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
print('h2d with concurrent d2h', round(n/dt/1e9,2))
