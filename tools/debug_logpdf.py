import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import numpy as np, torch
from cpg_b200 import sampling
from oracle import class_sampling as oc
fx = np.load(os.path.join(ROOT, 'tests/golden/class_sampling.npz'))
dev = torch.device('cuda')
w, m, cv = fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs']
gmm = sampling.GmmDevice(w, m, cv, dev)
x = torch.from_numpy(fx['z']).to(dev)
lq = sampling.gmm_logpdf(gmm, x).cpu().numpy()
want = oc.gmm_logpdf(fx['z'], w, m, cv)
# direct-form fp64 on the GPU with torch
xd = x.double()
mt, pt = gmm.mean_t.t(), gmm.prec_t.t()       # [K, D]
q = (((xd[:, None, :] - mt[None]) ** 2) * pt[None]).sum(2)
lp = gmm.logw_norm[None] - 0.5 * q
t = torch.logsumexp(lp, 1).cpu().numpy()
# direct-form fp64 in numpy
prec = 1.0 / cv
qn = (((fx['z'].astype(np.float64)[:, None, :] - m[None]) ** 2) * prec[None]).sum(2)
lpn = gmm.logw_norm.cpu().numpy()[None] - 0.5 * qn
mx = lpn.max(1, keepdims=True); n = (mx + np.log(np.exp(lpn - mx).sum(1, keepdims=True)))[:, 0]
print('kernel vs oracle(expanded)', np.abs(lq - want).max())
print('kernel vs torch direct    ', np.abs(lq - t).max())
print('kernel vs numpy direct    ', np.abs(lq - n).max())
print('numpy direct vs expanded  ', np.abs(n - want).max())
