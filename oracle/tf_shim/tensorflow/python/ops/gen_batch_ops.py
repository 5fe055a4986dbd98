def batch(*a, **k):  # imported (unused) by knapsack_v2/env.py:5
    raise NotImplementedError
