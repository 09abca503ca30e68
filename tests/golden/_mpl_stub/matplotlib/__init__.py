"""Empty stand-in so the reference (which imports matplotlib.pyplot at module top,
sgmethods/utils.py:8) can be imported where matplotlib is not installed. Used only by
tests/golden/make_golden.py."""
