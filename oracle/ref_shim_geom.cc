// TEST INFRASTRUCTURE ONLY.
//
// pb::PEenv::geom (src/pb/PEenv.cc:335-611), the reference's choice of the px x py x pz
// decomposition, called directly: it is a private member that reads the task count from
// the object, so this file is compiled with -fno-access-control (oracle/Makefile) and
// sets n_mpi_tasks_ on a default-constructed PEenv (the serial MPI stub has one rank).
// The reference source is compiled unmodified; mgmol_b200/parallel.py::geom is pinned
// against it (tests/test_parallel_cpu.py).
#include <sstream>

#include <mpi.h>

#include "PEenv.h"

extern "C" int ref_geom(int nx, int ny, int nz, int ntasks, int bias, int out[3])
{
    std::ostringstream sink;
    pb::PEenv pe(MPI_COMM_WORLD, &sink);
    pe.n_mpi_tasks_ = ntasks;
    for (int i = 0; i < 3; i++)
        pe.n_mpi_tasks_dir_[i] = 1;
    const int nmpi = pe.geom(nx, ny, nz, bias);
    for (int i = 0; i < 3; i++)
        out[i] = pe.n_mpi_tasks_dir_[i];
    return nmpi;
}
