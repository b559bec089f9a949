import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import quantit_b200 as qb
from quantit_b200 import workloads as wl
ctx = qb.default_context()
def timeit(A,B,da,db,reps=20):
    C = A.tensordot(B,da,db); ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream)
    with torch.cuda.stream(st):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): A.tensordot_(B,da,db,C)
        ev0.record(st)
        for _ in range(reps): A.tensordot_(B,da,db,C)
        ev1.record(st)
        ev1.synchronize()
    return ev0.elapsed_time(ev1)/reps
for name,cfg in [('T1',wl.T1),('T2',wl.T2),('T2_8K',wl.T2_8K)]:
    a,b,da,db = wl.tdot_pair(**cfg)
    A,B = qb.BTensor.from_host(**a), qb.BTensor.from_host(**b)
    info = A.tensordot_info(B,da,db)
    ms = timeit(A,B,da,db, 20 if name!='T2_8K' else 5)
    print(name, info, f'{ms:.4f} ms  {info["flops"]/ms/1e9:.2f} TFLOP/s', flush=True)
x = torch.randn(8192,8192,dtype=torch.float64,device='cuda'); y = torch.randn(8192,8192,dtype=torch.float64,device='cuda')
for _ in range(2): z = x@y
torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(5): z = x@y
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)/5
print(f'cuBLAS DGEMM 8192^3: {ms:.3f} ms {2*8192**3/ms/1e9:.2f} TFLOP/s')
