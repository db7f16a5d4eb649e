"""Kim-CNN attribute classifier container with the reference's names (models/classifier.py);
forward runs csrc/cnn.cu (eval mode).  Phase-2 training of the classifier is not shipped by the
reference and is out of scope here."""
import torch.nn as nn


def build_classifier(classifier_type, emb_dim, **C_args):
    if classifier_type != 'cnn':
        raise ValueError('Please use CNN classifier')
    return CNNClassifier(emb_dim, **C_args)


class CNNClassifier(nn.Module):
    def __init__(self, emb_dim, min_filter_width, max_filter_width, num_filters, dropout):
        super().__init__()
        if not (emb_dim == 150 and min_filter_width == 3 and max_filter_width == 5 and num_filters == 100):
            raise NotImplementedError('cpg_b200 CNN kernel is built for widths 3..5, 100 filters, emb 150')
        self.max_filter_width = max_filter_width
        self.conv_layers = nn.ModuleList(
            [nn.Conv2d(1, num_filters, (w, emb_dim)) for w in range(min_filter_width, max_filter_width + 1)])
        self.fc = nn.Sequential(nn.Dropout(dropout),
                                nn.Linear(num_filters * (max_filter_width - min_filter_width + 1), 2))

    def forward(self, x):
        raise RuntimeError('CNNClassifier is evaluated through RNN_VAE.forward_classifier (token ids -> cnn kernel)')
