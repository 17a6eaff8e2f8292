"""Empty stand-in so the reference's `from matplotlib import pyplot` imports succeed
(test infrastructure only; plotting is out of scope)."""
