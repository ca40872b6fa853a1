"""Dummy matplotlib (TEST INFRASTRUCTURE): the reference imports it at module scope for plotting only."""
