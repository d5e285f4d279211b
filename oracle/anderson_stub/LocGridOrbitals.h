// TEST INFRASTRUCTURE ONLY: stands in for src/LocGridOrbitals.h (see ExtendedGridOrbitals.h here).
