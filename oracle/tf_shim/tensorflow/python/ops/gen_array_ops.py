def gather(*a, **k):  # imported (unused) by resource_v3/env.py:5
    raise NotImplementedError
