"""GPU and CPU test suites (see tests/conftest.py for the `gpu` marker)."""
