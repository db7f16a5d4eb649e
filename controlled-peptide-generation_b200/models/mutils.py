"""Token-id constants and small model utilities with the reference's names (models/mutils.py)."""
import torch
import torch.nn as nn

from utils import check_dir_exists

# vocabulary layout fixed by the data loader (reference models/mutils.py:5-8, dataset.py:269-270)
UNK_IDX, PAD_IDX, START_IDX, EOS_IDX = 0, 1, 2, 3


def save_model(model, fn):
    """Checkpoint = torch.save(state_dict) with the reference's key set."""
    check_dir_exists(fn)
    torch.save(model.state_dict(), fn)
    print('Saved model to ' + fn)


def onehot_embed(hardIx, vocabSize):
    """indices [mb] -> one-hot rows [mb, vocabSize]."""
    assert hardIx.dim() == 1, 'expecting 1D tensor: minibatch of indices.'
    out = torch.zeros(hardIx.size(0), vocabSize, device=hardIx.device)
    return out.scatter_(1, hardIx.unsqueeze(1), 1.0)


def soft_embed(embed, softIx):
    """soft one-hots [mb, vocab] @ embedding matrix [vocab, emb_dim]."""
    assert isinstance(embed, nn.Embedding), 'Expecting nn.Embedding'
    return softIx @ embed.weight


def check_mask_eos(sentence, model):
    """Index of the single <eos> (or the length); asserts that only <pad> follows it."""
    sentence = sentence.view(-1)
    pos = (sentence == EOS_IDX).nonzero().view(-1)
    assert pos.numel() <= 1, 'expecting NO or SINGLE occurence of eos'
    end = int(pos[0]) if pos.numel() else sentence.numel()
    assert bool((sentence[end + 1:] == PAD_IDX).all()), 'there should be nothing but padding behind eos'
    return end
