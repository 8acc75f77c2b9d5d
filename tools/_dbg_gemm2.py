import sys, torch
sys.path.insert(0, '/root/repo')
from surfacenetworks_b200 import _native as N
dev = torch.device('cuda', 0)
st = torch.cuda.current_stream().cuda_stream
WS = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
M, Nn, K = 255168, 128, 256
A = torch.randn(M, K, device=dev); B = torch.randn(Nn, K, device=dev) / 16; bias = torch.randn(Nn, device=dev)
C = torch.empty(M, Nn, device=dev)
for i in range(3):
    N.call("sn_gemm_tf32_f32", A.data_ptr(), K, B.data_ptr(), K, bias.data_ptr(), 0, 0, 0, 0, 0, C.data_ptr(), Nn, M, Nn, K, 0, WS.data_ptr(), WS.numel(), st)
torch.cuda.synchronize()
