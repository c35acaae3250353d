"""cpflow_b200 — B200-native engine for cpflow's multi-start variational synthesis loop."""
from ._lib import CpflowError  # noqa: F401
