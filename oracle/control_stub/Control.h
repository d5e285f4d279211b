// TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the reference's src/Control.h (which pulls Boost
// program_options and cannot be compiled in this image), in the same spirit as
// the serial MPI stub: it lets src/GridMask.cc compile UNMODIFIED.  GridMask.cc
// reads exactly one thing from Control: the Poisson boundary conditions used by
// the minimum-image distance of the mask construction
// (src/GridMask.cc:146,162,211).  oracle/Makefile feeds GridMask.cc to g++
// on stdin so that this header, not src/Control.h, is the "Control.h" found.
#ifndef MGB_ORACLE_CONTROL_STUB_H
#define MGB_ORACLE_CONTROL_STUB_H

#include <fstream>

class Control
{
public:
    short bcPoisson[3];
    static Control* instance()
    {
        static Control c;
        return &c;
    }

private:
    Control() { bcPoisson[0] = bcPoisson[1] = bcPoisson[2] = 1; }
};

#endif
