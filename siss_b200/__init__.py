"""siss_b200 — B200-native (sm_100a) kernels for the SISS data-unlearning hot path.

Host side mirrors the reference's interface for this path:
  * ``siss_b200.losses.DDPMDeletionLoss``      — losses/ddpm_deletion_loss.py
  * ``siss_b200.scheduler.SissDDPMScheduler``  — the DDPMScheduler surface the tasks use (add_noise)
  * ``siss_b200.grad_combine.GradCombiner``    — the inline two-term gradient combine of delete_*.py
  * ``siss_b200.step.UnlearnStep``             — the fused fast path (nothing [B,D]-sized materialised)
  * ``siss_b200.optim.FusedCombineAdamW``      — combine + clip + AdamW (+EMA) in one pass, ZeRO-1 under data parallel
  * ``siss_b200.graph.CapturedStep``           — one CUDA-graph launch per optimiser step for the launch-bound shapes
All compute goes through libsiss_b200.so (C ABI in include/siss_b200.h), reached through the thin torch extension
``torch.ops.siss_b200.*`` (hot ops) or ctypes. No CPU fallback.
"""
from ._lib import SissLibraryError  # noqa: F401

__version__ = "0.1.0"
