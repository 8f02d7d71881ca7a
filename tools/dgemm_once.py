import torch
n = 4096
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize()
