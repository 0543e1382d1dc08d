from torch import nn

from .. import functional as Fn
from ._common import require_cuda, xavier_reset


class Regressor(nn.Module):
    """D -> hidden -> 32 -> 1 MLP with sigmoid output — reference models/Regressor.py:4-20."""

    def __init__(self, input_feature_dim, dropout_rate=0.6, hidden_dim=512, weight_init=True):
        super(Regressor, self).__init__()
        self.regressor = nn.Sequential(nn.Linear(input_feature_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout_rate),
                                       nn.Linear(hidden_dim, 32), nn.Dropout(dropout_rate),
                                       nn.Linear(32, 1), nn.Sigmoid())
        if weight_init == True:  # noqa: E712
            self._reset_parameters()

    def _reset_parameters(self):
        xavier_reset(self)

    def forward(self, x):
        require_cuda(x, "Regressor")
        x = x.view([-1, x.shape[-1]])
        seq = self.regressor
        cfg = Fn.HeadConfig(sigmoid=True, drop1=Fn.next_dropout(seq[2].p, self.training),
                            drop2=Fn.next_dropout(seq[4].p, self.training))
        return Fn.HeadFn.apply(x, seq[0].weight, seq[0].bias, seq[3].weight, seq[3].bias, seq[5].weight,
                               seq[5].bias, cfg)
