"""Mirror of the reference's ``src/models/denoiser_interp_levels_causal.py``: the Stage-2 denoiser with the additive
-inf upper-triangular attention mask (``causal=True`` at line 49 is the only difference from the bidirectional one)."""
from .denoiser_interp_levels import InterpLevelDenoiser


class InterpLevelCausalDenoiser(InterpLevelDenoiser):
    _causal = True
