// TEST INFRASTRUCTURE ONLY: stands in for src/ExtendedGridOrbitals.h (-> Control.h -> Boost) so that
// src/AndersonMix.cc compiles with -DTESTING; AndersonMix<Solution> does not use it.
