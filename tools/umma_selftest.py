"""tcgen05 operand conventions of f2_cnn.cu checked against torch on the device: 128 x N x K products,
K-major plane layout, descriptor start address advanced by `shift` rows (the convolution-tap trick)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from f2cnn_b200 import _native

L = _native.lib()
torch.manual_seed(0)
worst = 0.0
for N, K, shift, a_rows in ((32, 16, 0, 128), (32, 32, 0, 128), (64, 64, 0, 128), (32, 288, 0, 128), (32, 32, 1, 136),
                            (32, 32, 131, 264), (64, 32, 66, 200), (176, 64, 0, 128), (256, 32, 5, 136), (16, 16, 3, 136)):
    A = torch.randn(a_rows, K, device="cuda").to(torch.bfloat16).contiguous()
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16).contiguous()
    want = A[shift:shift + 128].float() @ B.float().t()
    for variant in (0,):  # variant 1 (offsets swapped) reads outside shared memory: illegal address, as it should
        D = torch.full((128, N), float("nan"), device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = L.f2_umma_selftest(A.data_ptr(), a_rows, B.data_ptr(), N, K, shift, variant, D.data_ptr(), status.data_ptr(), None)
        torch.cuda.synchronize()
        err = float((D - want).abs().max() / want.abs().max()) if rc == 0 else float("nan")
        print("N=%3d K=%3d shift=%3d variant=%d: rc=%d status=%d  max rel err %.3e" % (N, K, shift, variant, rc, int(status.item()), err), flush=True)
        if variant == 0:
            worst = max(worst, err if err == err else 1e9)
print("WORST variant-0 error %.3e -> %s" % (worst, "OK" if worst < 1e-5 else "MISMATCH"))
