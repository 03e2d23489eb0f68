"""cuBLAS DGEMM on the cross-covariance shape (tall operand stack x shared data matrix):
what the vendor library reaches on [M x Kd] . [Kd x N] with a short contraction, output
stored.  Usage: python scripts/cublas_shape_probe.py [M] [N] [Kd]"""
import sys
import torch

M = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
Kd = int(sys.argv[3]) if len(sys.argv) > 3 else 208
dev = torch.device('cuda', 0)
a = torch.randn(M, Kd, dtype=torch.float64, device=dev)
x = torch.randn(Kd, N, dtype=torch.float64, device=dev)
out = torch.empty(M, N, dtype=torch.float64, device=dev)
for _ in range(2):
    torch.matmul(a, x, out=out)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    torch.matmul(a, x, out=out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print('cublas dgemm %dx%dx%d: %.2f ms, %.2f TFLOP/s, store %.2f TB/s' % (
    M, N, Kd, best, 2.0 * M * N * Kd / best / 1e9, 8.0 * M * N / best / 1e9))
