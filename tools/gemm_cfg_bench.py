import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
shapes = [("s3 proj+res", 50176, 384, 384, True, False), ("s3 fc2+res", 50176, 384, 1536, True, False),
          ("s3 qkv", 50176, 1152, 384, False, False), ("s3 fc1+gelu", 50176, 1536, 384, False, True),
          ("s4 proj+res", 12544, 768, 768, True, False), ("s4 fc2+res", 12544, 768, 3072, True, False),
          ("s2 proj+res", 200704, 192, 192, True, False), ("s2 fc2+res", 200704, 192, 768, True, False)]
for name, M, N, K, res, gelu in shapes:
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda", dtype=torch.bfloat16) if res else None
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(x, w, b, residual=r, gelu=gelu, out=y); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[4]
    print(f"cfg={os.environ.get('MVIT_GEMM_CFG','auto'):4s} {name:12s} M={M} N={N} K={K}: {t*1e3:7.1f} us  {2.0*M*N*K/t/1e9:6.0f} TF/s")
