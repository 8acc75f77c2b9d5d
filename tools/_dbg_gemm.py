import sys, time, torch
sys.path.insert(0, '/root/repo')
from surfacenetworks_b200 import _native as N
dev = torch.device('cuda', 0)
st = torch.cuda.current_stream().cuda_stream
WS = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
for M, Nn, K in ((20000,128,256),(40000,128,256),(128000, 128, 256),(255168,256,128)):
    A = torch.randn(M, K, device=dev); B = torch.randn(Nn, K, device=dev) / 16; bias = torch.randn(Nn, device=dev)
    R = torch.randn(M, Nn, device=dev); C = torch.empty(M, Nn, device=dev)
    for flags in (0, 1):
        torch.cuda.synchronize(); t = time.time()
        for _ in range(5):
            N.call("sn_gemm_tf32_f32", A.data_ptr(), K, B.data_ptr(), K, bias.data_ptr(), R.data_ptr(), Nn, 0, 0, 0, C.data_ptr(), Nn, M, Nn, K, flags, WS.data_ptr(), WS.numel(), st)
        torch.cuda.synchronize()
        print(M, Nn, K, flags, 'ms per launch', (time.time() - t) / 5 * 1e3, flush=True)
    ref = torch.addmm(bias, A, B.t()) + R
    print('maxdiff', float((C - ref).abs().max()), flush=True)
