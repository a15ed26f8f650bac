"""Import shim so pydbspm/SPMGrid.py (which imports matplotlib.pyplot at module scope for its plotting helpers) can be
imported in a container without matplotlib.  TEST INFRASTRUCTURE ONLY (oracle/make_golden.py); nothing is plotted."""
