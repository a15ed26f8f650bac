# CODATA-2014 values as shipped by ase 3.22 (ase/units.py)
Bohr = 0.5291772105638411
Ha = 27.211386024367243
