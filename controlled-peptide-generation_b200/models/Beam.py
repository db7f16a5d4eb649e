"""Host-side beam bookkeeping with the reference's interface (models/Beam.py) for callers that
drive a beam step by step.  The CLaSS decode path does NOT use this class: RNN_VAE.sample_G runs the
fused device beam search (csrc/decode.cu) which implements the same advance / finish / back-track
rules, with the tie rule "larger score first, then lower flat index"."""
import torch


class Beam(object):
    def __init__(self, size, pad, bos, eos, n_best=1, device=torch.device('cpu'), min_length=0):
        self.size, self.device = size, device
        self.scores = torch.zeros(size, device=device)
        self.all_scores, self.prev_ks = [], []
        first = torch.full((size,), pad, dtype=torch.long, device=device)
        first[0] = bos
        self.next_ys = [first]
        self._eos, self._bos = eos, bos
        self.eos_top = False
        self.finished = []
        self.n_best, self.min_length = n_best, min_length

    def get_current_state(self):
        return self.next_ys[-1]

    def get_current_origin(self):
        return self.prev_ks[-1]

    def done(self):
        return self.eos_top and len(self.finished) >= self.n_best

    def advance(self, word_probs):
        assert not self.done(), 'not expecting to advance once done'
        n_words = word_probs.size(1)
        if len(self.next_ys) < self.min_length:
            word_probs[:, self._eos] = -1e20
        word_probs[:, self._bos] = -1e20                        # never predict <start>
        if self.prev_ks:
            cand = word_probs + self.scores.unsqueeze(1)
            cand[self.next_ys[-1] == self._eos] = -1e20          # ended hypotheses have no children
        else:
            cand = word_probs[0]
        flat = cand.reshape(-1)
        # larger score first, then lower flat index (stable descending sort)
        order = torch.sort(flat, descending=True, stable=True).indices[:self.size]
        self.all_scores.append(self.scores)
        self.scores = flat[order]
        prev_k = torch.div(order, n_words, rounding_mode='floor')
        self.prev_ks.append(prev_k)
        self.next_ys.append(order - prev_k * n_words)
        for i in range(self.size):
            if self.next_ys[-1][i] == self._eos:
                self.finished.append((self.scores[i], len(self.next_ys) - 1, i))
        if self.next_ys[-1][0] == self._eos:
            self.all_scores.append(self.scores)
            self.eos_top = True

    def sort_finished(self, minimum=None):
        if minimum is not None:
            i = 0
            while len(self.finished) < minimum:
                self.finished.append((self.scores[i], len(self.next_ys) - 1, i))
                i += 1
        self.finished.sort(key=lambda a: -a[0])
        return [s for s, _, _ in self.finished], [(t, k) for _, t, k in self.finished]

    def get_hyp(self, timestep, k):
        hyp = []
        for j in range(timestep - 1, -2, -1):
            hyp.append(self.next_ys[j + 1][k])
            if j >= 0:
                k = self.prev_ks[j][k]
        return hyp[::-1]
