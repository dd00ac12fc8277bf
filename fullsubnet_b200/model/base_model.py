"""Host-side pieces of audio_zen/model/base_model.py that the drop-in Model needs:
norm_wrapper (:356-372) name checking and weight_init (:374-439, CPU-side initialisation)."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.init as init


class BaseModel(nn.Module):
    NORM_TYPES = {"offline_laplace_norm": 0, "cumulative_laplace_norm": 1}
    _UPSTREAM_NORMS = ("offline_laplace_norm", "cumulative_laplace_norm", "offline_gaussian_norm",
                       "cumulative_layer_norm", "forgetting_norm")

    def __init__(self):
        super().__init__()

    def norm_wrapper(self, norm_type: str) -> int:
        if norm_type in self.NORM_TYPES:
            return self.NORM_TYPES[norm_type]
        if norm_type in self._UPSTREAM_NORMS:
            raise NotImplementedError(f"norm_type {norm_type!r} is not built into libfsn_b200 yet (SURVEY 8f)")
        raise NotImplementedError(
            "You must set up a type of Norm. e.g. offline_laplace_norm, cumulative_laplace_norm, forgetting_norm, etc.")

    def weight_init(self, m):
        """base_model.py:374-439 restricted to the module types this model contains."""
        if isinstance(m, nn.Linear):
            init.xavier_normal_(m.weight.data)
            init.normal_(m.bias.data)
        elif isinstance(m, (nn.LSTM, nn.GRU)):
            for param in m.parameters():
                if len(param.shape) >= 2:
                    init.orthogonal_(param.data)
                else:
                    init.normal_(param.data)
