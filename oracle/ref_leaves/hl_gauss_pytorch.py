"""Stand-in for hl-gauss-pytorch HLGaussLayer, regression mode only (reference: e2_tts.py:1035-1040,
1107, 1111). Linear(dim,1) -> activation -> squeeze; loss = mse. Test infrastructure only.
UNPINNED detail: presence of the Linear bias (upstream version dependent)."""
import torch.nn.functional as F
from torch import nn


class HLGaussLayer(nn.Module):
    def __init__(self, dim, *, hl_gauss_loss=None, use_regression=False, regress_activation=None, **kw):
        super().__init__()
        assert use_regression and hl_gauss_loss is None, 'only regression mode is restated'
        self.to_pred = nn.Sequential(nn.Linear(dim, 1), regress_activation or nn.Identity())

    def forward(self, embed, target=None):
        pred = self.to_pred(embed).squeeze(-1)
        if target is None:
            return pred
        return F.mse_loss(pred, target)
