import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from surfacenetworks_b200 import _native as N
DEV='cuda'
def run(A,B,flags=1):
    R,M=A.shape; Nn=B.shape[1]
    G=torch.full((M,Nn),-7.0,device=DEV); wsb=N.lib.sn_gemm_tn_tf32_ws_bytes(R,Nn); ws=torch.empty(wsb,dtype=torch.uint8,device=DEV)
    N.call("sn_gemm_tn_tf32_f32",A.data_ptr(),A.stride(0),B.data_ptr(),B.stride(0),G.data_ptr(),G.stride(0),R,M,Nn,flags,ws.data_ptr(),wsb,torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); return G
R,Nn=32,256
torch.manual_seed(0)
A=torch.randint(-2,3,(R,128),device=DEV).float(); B=torch.randint(-2,3,(R,Nn),device=DEV).float()
ref=A.t()@B
for dbg in [0]:
    G=run(A,B,1|(dbg<<8))
    print('dbg',dbg,'nonzero frac',float((G!=0).float().mean()),'match frac',float((G==ref).float().mean()), 'G[0,:6]',G[0,:6].tolist(),'ref',ref[0,:6].tolist())
