"""A*A on the 3-D 27-point Laplacian (rows of 729 products / 125 columns: the warp-per-row hash class)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A = g4s_b200.CSR.laplacian3d27(n)
for _ in range(4):
    g4s_b200.HashSpGEMM(A, A).make_empty()
torch.cuda.synchronize()
print("ok")
