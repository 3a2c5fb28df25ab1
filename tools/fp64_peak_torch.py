"""Library FP64 baselines on the box (cuBLAS DGEMM/ZGEMM, cuSOLVER/MAGMA batched inverse).
Used only to state the roofline denominator and a library baseline; never on the product path."""
import json, time, torch
def ev(f, n=3):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); b.synchronize(); best = min(best, a.elapsed_time(b))
    return best
out = {}
A = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); B = torch.randn_like(A)
ms = ev(lambda: A @ B); out["dgemm_8192_tflops"] = 2 * 8192**3 / ms / 1e9
A = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda"); B = torch.randn_like(A)
ms = ev(lambda: A @ B); out["zgemm_4096_tflops"] = 8 * 4096**3 / ms / 1e9
for N, bs in ((256, 256), (1024, 64)):
    A = torch.randn(bs, N, N, dtype=torch.complex128, device="cuda") + 3 * torch.eye(N, dtype=torch.complex128, device="cuda")
    ms = ev(lambda: torch.linalg.inv(A)); out[f"torch_inv_N{N}_b{bs}_ms"] = ms
    out[f"torch_inv_N{N}_tflops"] = 8 * N**3 * bs / ms / 1e9
    # rank-32 batched update like one elimination step
    P = torch.randn(bs, N, 32, dtype=torch.complex128, device="cuda"); W = torch.randn(bs, 32, N, dtype=torch.complex128, device="cuda")
    ms = ev(lambda: torch.baddbmm(A, P, W, alpha=-1)); out[f"baddbmm_k32_N{N}_tflops"] = 8 * N * N * 32 * bs / ms / 1e9
    P = torch.randn(bs, N, 64, dtype=torch.complex128, device="cuda"); W = torch.randn(bs, 64, N, dtype=torch.complex128, device="cuda")
    ms = ev(lambda: torch.baddbmm(A, P, W, alpha=-1)); out[f"baddbmm_k64_N{N}_tflops"] = 8 * N * N * 64 * bs / ms / 1e9
print(json.dumps(out))
