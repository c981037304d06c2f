from .ddpm_deletion_loss import DDPMDeletionLoss  # noqa: F401
