"""Thresholds of the reference (code/utils/constants.py:2,6)."""
MIN_N_VIEWS_PER_POINT = 2
MIN_N_POINTS_PER_VIEW = 8
