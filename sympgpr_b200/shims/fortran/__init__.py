"""Stands in for python/fortran/ (the install location of the f2py `sympgpr` extension,
python/fortran/__init__.py:1-8): `from fortran.sympgpr import sympgpr`."""
