// TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the reference's src/Control.h (which pulls Boost
// program_options and cannot be compiled in this image), in the same spirit as
// the serial MPI stub: it lets src/GridMask.cc compile UNMODIFIED.  GridMask.cc
// reads exactly one thing from Control: the Poisson boundary conditions used by
// the minimum-image distance of the mask construction
// (src/GridMask.cc:146,162,211).  oracle/Makefile feeds GridMask.cc to g++
// on stdin so that this header, not src/Control.h, is the "Control.h" found.
// src/PCGSolver.h includes "Control.h" from a header that sits next to the real
// one; for it the Makefile pre-includes this file (-include), and the real
// header's own guard (CONTROL_H, defined here) turns its body off.  PCGSolver
// reads Control::lap_type (src/PCGSolver.h:74-75), the operator of the
// preconditioner's fine level.
#ifndef MGB_ORACLE_CONTROL_STUB_H
#define MGB_ORACLE_CONTROL_STUB_H
#ifndef CONTROL_H
#define CONTROL_H
#endif

#include <fstream>
#include <string>

class Control
{
public:
    short bcPoisson[3];
    short lap_type;
    // src/Species.cc and src/radial/RadialMeshFunction.cc (row f3: the sparse KB
    // projectors) read the verbosity, whether this is a restart (both only gate
    // printing) and the output file name helper
    short verbose;
    bool restart_run;
    std::string getFullFilename(const std::string& name) const { return name; }
    static Control* instance()
    {
        static Control c;
        return &c;
    }

private:
    Control() : lap_type(0), verbose(0), restart_run(false)
    {
        bcPoisson[0] = bcPoisson[1] = bcPoisson[2] = 1;
    }
};

#endif
