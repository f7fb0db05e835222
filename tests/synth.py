"""Re-export of the synthetic batch generator (lives in the package so bench.py can use it too)."""
from toist_b200.synth import make_args, make_batch, targets_to  # noqa: F401
