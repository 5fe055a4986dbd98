node = None  # imported (unused) by resource_v3/reward.py:3
