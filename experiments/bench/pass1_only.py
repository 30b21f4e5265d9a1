"""pass 1 alone at c3 shape (for ncu dram-bytes measurements under different knobs)."""
import sys
sys.path.insert(0, ".")
import torch
from gppvae_b200 import ops
N, Q, L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 4096, 256
dev = torch.device("cuda:0")
V = torch.randn(N, Q, device=dev) / 64; Z = torch.randn(N, L, device=dev)
for _ in range(2):
    GC = ops.gram_vtz(V, Q, Z, L, N, Q, L)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    GC = ops.gram_vtz(V, Q, Z, L, N, Q, L)
e1.record(); torch.cuda.synchronize()
print(f"pass1 {e0.elapsed_time(e1) / 3:.2f} ms")
