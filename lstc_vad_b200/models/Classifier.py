from torch import nn

from .. import functional as Fn
from ._common import require_cuda, xavier_reset


class Classifier(nn.Module):
    """D -> 512 -> 32 -> 2 MLP with softmax output (probabilities) — reference models/Classifier.py:5-23.
    Layer 1 (Linear+ReLU+Dropout) is a tcgen05 GEMM with a fused epilogue; layers 2-3 + softmax are one kernel."""

    def __init__(self, input_feature_dim, dropout_rate=0.6, weight_init=True):
        super(Classifier, self).__init__()
        self.classifier = nn.Sequential(nn.Linear(input_feature_dim, 512), nn.ReLU(), nn.Dropout(dropout_rate),
                                        nn.Linear(512, 32), nn.Dropout(dropout_rate),
                                        nn.Linear(32, 2), nn.Softmax(dim=-1))
        if weight_init == True:  # noqa: E712
            self._reset_parameters()

    def _reset_parameters(self):
        xavier_reset(self)

    def forward(self, x):
        require_cuda(x, "Classifier")
        x = x.view([-1, x.shape[-1]])
        seq = self.classifier
        cfg = Fn.HeadConfig(sigmoid=False, drop1=Fn.next_dropout(seq[2].p, self.training),
                            drop2=Fn.next_dropout(seq[4].p, self.training))
        return Fn.HeadFn.apply(x, seq[0].weight, seq[0].bias, seq[3].weight, seq[3].bias, seq[5].weight,
                               seq[5].bias, cfg)
