"""Same module path and class name as the reference's losses/ddpm_deletion_loss.py; see compat/losses/__init__.py."""
from siss_b200.losses.ddpm_deletion_loss import DDPMDeletionLoss  # noqa: F401

__all__ = ["DDPMDeletionLoss"]
