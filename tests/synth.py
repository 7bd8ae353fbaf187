"""Synthetic problems for the tests: re-exports of workloads.py (generator) and oracle/problems.py (oracle builder)."""
from oracle.problems import build_oracle  # noqa: F401
from workloads import _gram, make_problem, round_f32  # noqa: F401
