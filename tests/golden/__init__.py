"""Golden fixtures generated from the unmodified reference (make_golden.py) and their loader."""
