import sys, time
sys.path.insert(0,'/root/repo')
import bench, torch
from emdee_b200 import api
lib = api.load()
R,P,L = bench.make_workload(63)
s = bench.build_system(lib,R,P,L,1)
for _ in range(10): bench.md_step(s)
torch.cuda.synchronize()
t0=time.perf_counter()
tb1=td=tb2=0
for _ in range(50):
    a=time.perf_counter(); s.boost(1,0,0.0025); b=time.perf_counter(); s.displace(1,0,0.005); c=time.perf_counter(); s.boost(1,0,0.0025); d=time.perf_counter()
    tb1+=b-a; td+=c-b; tb2+=d-c
torch.cuda.synchronize()
print("per step ms", (time.perf_counter()-t0)/50*1e3, "boost1", tb1/50*1e3, "displace", td/50*1e3, "boost2", tb2/50*1e3, "builds", s.md.Builds)
s.finalize()
