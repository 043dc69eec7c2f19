"""Test-side alias of the synthetic workload generator (rsplash_b200/synthetic.py)."""
from rsplash_b200.synthetic import make_problem  # noqa: F401
