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
