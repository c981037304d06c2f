"""Import-path shim: put ``<repo>/compat`` in front of the reference checkout on PYTHONPATH and the
reference's own ``from losses.ddpm_deletion_loss import DDPMDeletionLoss`` (delete_celeb.py:45,
delete_tshirt.py, delete_sd.py) resolves to the sm_100a implementation — no source edit at all:

    PYTHONPATH=/path/to/siss-b200/compat:/path/to/siss-b200 python main.py --config-name=delete_celeb
"""
