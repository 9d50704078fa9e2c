"""Minimal stand-in for the `normflows` package (absent from this image, no network).

The reference imports it at module level in src/flows/discrete_flow.py:18 (and touches
`distributions.BaseDistribution`, `flows.MaskedAffineFlow`, `NormalizingFlow` at :72,79,319)
but the molecular pipeline never instantiates those classes: molecules use
ParticleConservingFlowSampler (pipeline.py:344-354).  Test infrastructure only."""
import types

import torch.nn as nn


class _Base(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


distributions = types.SimpleNamespace(BaseDistribution=_Base)
flows = types.SimpleNamespace(MaskedAffineFlow=_Base)


class NormalizingFlow(_Base):
    pass
