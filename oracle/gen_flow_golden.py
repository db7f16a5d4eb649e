"""Golden vectors of the flow transforms and of the soft sampling modes from the LIVE reference (models/flow.py and
RNN_VAE.sample_G imported unmodified from /root/reference): python -m oracle.gen_flow_golden  ->  tests/golden/flow.npz.  TEST INFRASTRUCTURE ONLY."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    spec = importlib.util.spec_from_file_location('ref_flow', '/root/reference/models/flow.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    fx = {'versions': np.array(['torch ' + torch.__version__])}
    for kind, n in (('planar', 3), ('radial', 2), ('alternating', 5)):
        torch.manual_seed(11)
        r = ref.build_flow(kind, n, 100)
        with torch.no_grad():
            for p in r.parameters():
                p.mul_(20.0)                       # the reference's 0.01-scale initialisation barely moves z
            if kind != 'planar':
                r.radial_beta[n - 1 if kind == 'radial' else 1].fill_(-30.0)     # exercises "maintain invertibility"
        for k, v in r.state_dict().items():
            fx['%s/param/%s' % (kind, k)] = v.clone().numpy()
        z = torch.randn(37, 100)
        fx[kind + '/z'] = z.numpy().copy()
        zt, loss = r(z.clone(), train=True)
        fx[kind + '/z_out'], fx[kind + '/loss'] = zt.detach().numpy(), np.float32(loss.item())
        for k, v in r.state_dict().items():        # parameters after the in-place maintenance
            fx['%s/param_after/%s' % (kind, k)] = v.clone().numpy()
    # soft sampling modes of sample_G (models/model.py:330-359) on the trained fixture weights, eval mode (no dropout)
    from . import refharness as rh
    pf = np.load(os.path.join(ROOT, 'tests', 'golden', 'params_trained_v24.npz'))
    model = rh.build_model(24)
    model.load_state_dict({k: torch.from_numpy(pf[k].copy()) for k in pf.files})
    model.eval()
    g = torch.Generator().manual_seed(21)
    z = torch.randn(16, 100, generator=g)
    c = torch.eye(2)[torch.randint(0, 2, (16,), generator=g)]
    fx['soft/z'], fx['soft/c'] = z.numpy(), c.numpy()
    for mode, temp in (('greedy_softmax', 1.0), ('greedy_softmax', 0.6), ('none_softmax', 1.0)):
        with torch.no_grad():
            ix, soft = model.sample_G(16, z, c, sample_mode=mode, temp=temp)
        tag = 'soft/%s_t%.1f' % (mode, temp)
        fx[tag + '/ix'], fx[tag + '/soft'] = ix.numpy(), soft.numpy()
    out = os.path.join(ROOT, 'tests', 'golden', 'flow.npz')
    np.savez_compressed(out, **fx)
    print('wrote', out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()
