import sys, torch
sys.path.insert(0, "/root/repo")
from valor_b200 import kernels as K
torch.manual_seed(0)
for (M, N, rows) in [(512, 1024, 3136), (512, 1024, 3200), (768, 512, 20000), (512, 1024, 6272), (256, 1024, 3136), (512, 512, 3136), (1024, 256, 3136), (512, 1024, 384), (512, 1024, 64*12)]:
    dy = (torch.randn(rows, M, device="cuda") * 0.1).bfloat16()
    x = torch.randn(rows, N, device="cuda").bfloat16()
    ref = dy.float().t() @ x.float()
    for fb in (1256, 2256, 0):
        for rep in range(3):
            out = torch.zeros(M, N, device="cuda")
            K.gemm(dy, x, a_kmajor=False, b_kmajor=False, out=out, accumulate=True, backend=K.BACKEND_TENSOR, force_bn=fb)
            torch.cuda.synchronize()
            err = (out - ref).abs().max().item() / ref.abs().max().item()
            nr = out.norm().item() / ref.norm().item()
            if rep == 0 or err > 1e-2:
                print(M, N, rows, "force_bn", fb, "rep", rep, "relmaxerr %.4g normratio %.5f" % (err, nr), flush=True)
