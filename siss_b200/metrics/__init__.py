from .class_membership import MembershipLoss

__all__ = ["MembershipLoss"]
