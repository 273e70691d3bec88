"""Attack classes of MegaAdversarial/src/attacks (same names, constructor arguments and call
convention `attack(x, y, kwargs) -> (x_adv, y)`), written against the fused ODE blocks:
input-gradient attacks run the backward in dgrad-only mode (no weight gradients are formed)."""
from .attacks import (Attack, Attack2Ensemble, Clean, Clean2Ensemble, FGSM, FGSMRandom, FGSM2Ensemble, PGD,  # noqa: F401
                      ensemble_logits)
