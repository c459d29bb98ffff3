"""Test infrastructure: see tests/mplstub/matplotlib."""
