"""Loss epilogue of the registration path.

``RegistrationReconstructionLoss`` keeps the name, constructor and call signature of
/root/reference/modules/loss/registration_losses.py:17-28,

    loss = 0.5 * MSE(tar, Sdef) / sigma^2 + w * sum(v * m) / numel(tar),

but when the prediction dict carries ``'registration_loss_terms'`` (P,2) - the per-pair sums
{sum (tar - Sdef)^2, sum v.m} the shooting kernel produced while the fields were on chip
(``shoot_warp_strain(..., loss_terms=True)`` / model config ``fused_loss_terms``) - it only adds 2 P numbers:
no re-read of Sdef, tar, v, m, and in the backward no gradient images (``b2_warp_sqerr_bwd`` recomputes Sdef
from the taps, the regularisation gradient is closed-form inside ``b2_shoot_bwd_loss``).
"""
from __future__ import annotations

import torch


def reconstruction_loss_from_terms(loss_terms, numel_tar, sigma=0.03, regularization_weight=0.1):
    """The reference's formula evaluated from the per-pair sums (P,2)."""
    tot = loss_terms.sum(dim=0)
    return 0.5 * (tot[0] / numel_tar) / (sigma * sigma) + regularization_weight * (tot[1] / numel_tar)


class RegistrationReconstructionLoss:
    def __init__(self, sigma, regularization_weight=1):
        self.sigma = sigma
        self.regularization_weight = regularization_weight

    def __call__(self, prediction, target):
        tar = target["registration_target"]
        terms = prediction.get("registration_loss_terms")
        if terms is not None:
            return reconstruction_loss_from_terms(terms, tar.numel(), self.sigma, self.regularization_weight)
        Sdef = prediction["deformed_source"]
        recon_loss = torch.mean((tar - Sdef) ** 2)
        regularization = (prediction["velocity"] * prediction["momentum"]).sum() / tar.numel()
        return 0.5 * recon_loss / (self.sigma * self.sigma) + regularization * self.regularization_weight
