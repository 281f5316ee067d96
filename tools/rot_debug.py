import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200 import rotation
dev = torch.device('cuda')
for (Y, X, Z) in ((8, 64, 64), (8, 160, 160), (16, 256, 256)):
    tab = rotation.device_table([Y, X, Z], 0.7, dev)
    obj = torch.rand((Z, Y, X, 2), device=dev)
    try:
        rot = rotation.rotate_db(obj, tab)
        torch.cuda.synchronize()
        t = tab.long()
        ref = obj[t[..., 1], :, t[..., 0]].permute(0, 2, 1, 3)     # [Z, X, Y, 2] -> [Z, Y, X, 2]
        print('gather', (Y, X, Z), 'ok', bool(torch.equal(rot, ref)))
    except Exception as e:
        print('gather', (Y, X, Z), 'FAILED', str(e)[:200]); break
    try:
        g = torch.rand((Z, 3, Y, X, 2), device=dev)
        out = torch.zeros((Z, Y, X, 2), device=dev)
        rotation.rotate_db_adjoint_batch(g, [tab] * 3, out, accumulate=False)
        torch.cuda.synchronize()
        print('adjoint', (Y, X, Z), 'ran', float(out.sum()), float(g.sum()))
    except Exception as e:
        print('adjoint', (Y, X, Z), 'FAILED', str(e)[:200]); break
