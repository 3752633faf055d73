"""reference ``geotransformer/modules/sinkhorn/learnable_sinkhorn.py:5-70``."""
import torch
import torch.nn as nn

from ... import functional as GF


class LearnableLogOptimalTransport(nn.Module):
    def __init__(self, num_iterations, inf=1e12):
        super().__init__()
        self.num_iterations = num_iterations
        self.register_parameter('alpha', torch.nn.Parameter(torch.tensor(1.0)))
        self.inf = inf

    def forward(self, scores, row_masks=None, col_masks=None):
        """scores (B, M, M) -> log-assignment (B, M+1, M+1); all iterations run inside one kernel."""
        return GF.sinkhorn(scores, row_masks, col_masks, self.alpha, self.num_iterations, self.inf)

    def __repr__(self):
        return self.__class__.__name__ + '(num_iterations={})'.format(self.num_iterations)
