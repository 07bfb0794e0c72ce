"""cupix_b200 -- B200-native batched Px likelihood (hot path of igmhub/cupix)."""
__version__ = "0.1.0"
