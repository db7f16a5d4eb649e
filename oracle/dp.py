"""Oracle for the data-parallel decomposition of one WAE iteration.

TEST INFRASTRUCTURE ONLY.  The reference has no distributed code; the contract is
"N ranks on contiguous shards of the batch == one process on the whole batch".  This file restates
the two local phases the CUDA library implements (cpg_wae_step_phase1 / phase2) with torch-CPU ops so
that the exchange logic (what is all-reduced, and when) can be checked with gloo on CPU.
"""
import torch

from . import wae as ow


def phase1_coupled(p, tokens, noise, rf_dim=ow.RF_DIM, sigma=ow.MMD_SIGMA):
    """Local statistics that couple the batch: [n_tok, 0, 5 latent sums, 0, RF sums of z, of z_prior]."""
    mu, logvar = ow.encoder_forward(p, tokens)
    z = ow.reparameterize(mu, logvar, noise['eps'])
    tgt = torch.cat([tokens[:, 1:], torch.full((tokens.shape[0], 1), ow.PAD_IDX, dtype=tokens.dtype)], 1)
    out = torch.zeros(8 + 2 * rf_dim)
    out[0] = float((tgt != ow.PAD_IDX).sum())
    out[2] = 0.5 * (logvar.exp() + mu ** 2 - 1 - logvar).sum()
    out[3] = 0.5 * (logvar.exp() - 1 - logvar).sum()
    out[4] = logvar.abs().sum()
    out[5] = mu.abs().sum()
    out[6] = logvar.sum()
    out[8:8 + rf_dim] = ow.gaussian_rf(z, noise['rf_w'], noise['rf_b'], sigma, rf_dim).sum(0)
    out[8 + rf_dim:] = ow.gaussian_rf(noise['z_prior_rf'], noise['rf_w'], noise['rf_b'], sigma, rf_dim).sum(0)
    return out.detach()


def phase2_local_grads(p, tokens, noise, coupled, global_batch, beta=1.0, lambda_kl=1e-3, lambda_l1=0.0,
                       rf_dim=ow.RF_DIM, sigma=ow.MMD_SIGMA):
    """This rank's share of the global-batch gradient, given the ALL-REDUCED `coupled` vector:
    CE divided by the global token count, RF-MMD linearised around the global feature means,
    KL-type terms divided by the global batch.  Returns (grads dict, local nll sum)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items() if k in ow.UNIQUE_VAE_PARAMS}
    mu, logvar = ow.encoder_forward(leaves, tokens)
    z = ow.reparameterize(mu, logvar, noise['eps'])
    logits = ow.decoder_forward(leaves, ow.word_dropout(tokens, noise['word_drop']), z, noise['c'], noise['out_keep'])
    B, L, V = logits.shape
    tgt = torch.cat([tokens[:, 1:], torch.full((B, 1), ow.PAD_IDX, dtype=tokens.dtype)], 1)
    lsm = logits - torch.logsumexp(logits, 2, keepdim=True)
    nll = -(lsm.gather(2, tgt.unsqueeze(2)).squeeze(2)) * (tgt != ow.PAD_IDX)
    recon_share = nll.sum() / coupled[0]
    delta = (coupled[8:8 + rf_dim] - coupled[8 + rf_dim:]) / global_batch            # global mean1 - mean2
    phi_sum = ow.gaussian_rf(z, noise['rf_w'], noise['rf_b'], sigma, rf_dim).sum(0)
    rf_share = (2.0 * delta.detach() * phi_sum / global_batch).sum()
    klsm_share = 0.5 * (logvar.exp() - 1 - logvar).sum() / global_batch
    l1_share = logvar.abs().sum() / global_batch
    (recon_share + beta * rf_share + lambda_kl * klsm_share + lambda_l1 * l1_share).backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in leaves}
    grads['word_emb.weight'][ow.PAD_IDX] = 0.0
    return grads, float(nll.sum())


# --------------------------------------------------------------------------------------------------
# An `eng` for cpg_b200.parallel.dp_train_step made of the oracle math above: lets the PRODUCT's exchange
# sequence (what is all-reduced, in which order, what rides in the gradient tail) run on CPU under gloo.
class OracleState:
    """Same attributes as cpg_b200.engine.FlatState (flat grads followed by the all-reduce tail)."""

    def __init__(self, p, tail=8):
        self.p = {k: v.clone() for k, v in p.items()}
        self.total = sum(self.p[k].numel() for k in ow.UNIQUE_VAE_PARAMS)
        self.grads_ext = torch.zeros(self.total + tail)
        self.grads = self.grads_ext[:self.total]
        self.adam = {}
        self.step = 0


class OracleEngine:
    SC = dict(loss=0, recon=1, kl=2, mmd=3, mmdrf=4, logvar_l1=5, logvar_kl=6, z_mu_l1=7, z_logvar=8,
              beta=9, grad_norm=10, ntok=11, nll_sum=12)

    def __init__(self):
        import contextlib
        self._null = contextlib.nullcontext
        self._nll = 0.0

    def stats_stream(self, device):
        return self._null()

    def ntok_stream(self, device):
        return self._null()

    def step_phase1(self, st, tokens, noise, hp, p_out=0.3):
        self._z = ow.reparameterize(*ow.encoder_forward(st.p, tokens), noise['eps']).detach()
        return phase1_coupled(st.p, tokens, noise, rf_dim=hp.rf_dim, sigma=hp.mmd_sigma), self._z

    def step_phase2(self, st, tokens, noise, hp, coupled, p_out=0.3):
        grads, nll = phase2_local_grads(st.p, tokens, noise, coupled, hp.global_batch, beta=hp.beta,
                                        lambda_kl=hp.lambda_logvar_kl, lambda_l1=hp.lambda_logvar_l1,
                                        rf_dim=hp.rf_dim, sigma=hp.mmd_sigma)
        st.grads.copy_(torch.cat([grads[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS]))
        self._nll = nll
        Bg = hp.global_batch
        sc = torch.zeros(16)
        S = self.SC
        sc[S['ntok']] = coupled[0]
        sc[S['nll_sum']] = nll
        sc[S['recon']] = nll / float(coupled[0])
        sc[S['kl']], sc[S['logvar_kl']], sc[S['logvar_l1']] = coupled[2] / Bg, coupled[3] / Bg, coupled[4] / Bg
        R = hp.rf_dim
        d = (coupled[8:8 + R] - coupled[8 + R:8 + 2 * R]) / Bg
        sc[S['mmdrf']] = (d ** 2).sum()
        sc[S['loss']] = (sc[S['recon']] + hp.beta * sc[S['mmdrf']] + hp.lambda_logvar_l1 * sc[S['logvar_l1']]
                         + hp.lambda_logvar_kl * sc[S['logvar_kl']])
        return sc

    def dp_pack_tail(self, tail):
        tail.zero_()
        tail[0] = self._nll

    def dp_apply_tail(self, tail, sc):
        S = self.SC
        recon = tail[0] / sc[S['ntok']]
        sc[S['loss']] += recon - sc[S['recon']]
        sc[S['recon']] = recon
        sc[S['nll_sum']] = tail[0]

    def clip_adam(self, st, hp, out=None):
        grads, off = {}, 0
        for k in ow.UNIQUE_VAE_PARAMS:
            n = st.p[k].numel()
            grads[k] = st.grads[off:off + n].view_as(st.p[k]).clone()
            off += n
        gn = ow.clip_and_adam(st.p, grads, st.adam, lr=hp.lr, max_norm=hp.clip_norm)
        if out is not None:
            out[0] = gn
        return torch.tensor([float(gn)])

    def mmd_full(self, z, zp, sigma, out=None):
        v = torch.tensor([float(ow.mmd_full_kernel(z, zp, sigma))])
        if out is not None:
            out.copy_(v)
        return v
