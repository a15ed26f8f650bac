"""Import shim for the uninstalled dftd3 package (pydbspm/interactions.py:4)."""
