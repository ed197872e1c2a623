import torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from occnet_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for (M, N, K) in [(40000, 256, 256), (184950, 256, 256), (40000, 768, 256)]:
    A = (torch.randn(M, K, device='cuda') * 0.5).bfloat16()
    W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
    b = torch.randn(N, device='cuda')
    C = torch.empty((M, N), device='cuda')
    for _ in range(2):
        _lib.check(lib.occb200_gemm_bf16_tc(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(C), M, N, K, _lib.stream_ptr()))
        torch.cuda.synchronize()
