from .discounted_return import DiscountedReturn, affine_scan, discounted_reward, gae

__all__ = ["DiscountedReturn", "discounted_reward", "gae", "affine_scan"]
