import sys, numpy as np
sys.path.insert(0, '/root/repo')
import pydynet_b200 as pdn
import pydynet_b200.nn.functional as F
which = sys.argv[1]
N, C, H, W, O, k, stride, pad = 4, 20, 14, 14, 50, 3, 1, 1
rng = np.random.default_rng(0)
x = rng.standard_normal((N, C, H, W)).astype(np.float32)
w = rng.standard_normal((O, C, k, k)).astype(np.float32)
tx = pdn.Tensor(x, dtype=np.float32, device='cuda:0', requires_grad=(which != 'fwd'))
tw = pdn.Tensor(w, dtype=np.float32, device='cuda:0', requires_grad=(which == 'all'))
out = F.conv2d(tx, tw, pad, stride)
pdn.cuda.synchronize()
print('fwd ok', float(np.abs(out.numpy()).sum()))
if which != 'fwd':
    out.sum().backward()
    pdn.cuda.synchronize()
    print('bwd ok')
