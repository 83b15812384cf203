import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gplasdi_b200.engine import Engine
eng = Engine(0)
Z = torch.randn(8192, 1001, 5, device="cuda")
for _ in range(2):
    eng.fd_dzdt(Z, 1e-3, "sbp12"); eng.sindy_fit(Z, 1e-3, "sbp12", 1)
torch.cuda.synchronize()
print("ok")
