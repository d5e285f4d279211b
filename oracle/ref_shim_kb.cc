// TEST INFRASTRUCTURE ONLY.
//
// extern "C" face of the reference's OWN non-local projector code,
// compiled unmodified where it lies: src/KBprojectorSparse.cc, src/Species.cc (reads
// the pseudopotential files of /root/reference/potentials), src/Mesh.cc, src/radial/*.
// Built twice by oracle/Makefile -- ORBDTYPE = KBPROJDTYPE double (libmgmol_refkb_f64.so)
// and, with -DUSE_MP, float (libmgmol_refkb_f32.so, src/global.h:18-22,38) -- because the
// reference fixes that type at compile time.
//
// What it exposes is exactly what the path consumes and computes:
//   * the sparse projector of an ion at a given centre as KBprojectorSparse::setup builds
//     it (node list nlindex_, the value arrays in getProjectors order, kbcoeff * sign);
//     these private members are read by compiling THIS file with -fno-access-control
//     (oracle/Makefile) -- the reference sources themselves are not touched;
//   * <beta|psi>: KBprojector::registerPsi + vel * dotPsi(iloc, i), the body of
//     KBPsiMatrixInterface::computeLocalElement (src/KBPsiMatrixInterface.cc:20-60);
//   * vnlpsi: the loop of get_vnlpsi (src/get_vnlpsi.cc:24-87) over the ions -- memset,
//     axpySKet for a single projector, axpyKet otherwise -- and the caller's
//     MPaxpy(numpt, 1., vnlpsi, hpsi) (src/computeHij.cc:361-363).
// get_vnlpsi.cc itself needs Ions / KBPsiMatrixSparse / MGmol (the whole driver); its
// loop is restated here over the reference's own KBprojectorSparse objects.
#include <cassert>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <mpi.h>
#include <omp.h>

#include "KBprojectorSparse.h"

#include "MGmol_MPI.h"
#include "MPIdata.h"
#include "Mesh.h"
#include "Species.h"
#include "mputils.h"

namespace
{
Species* g_species = nullptr;
std::vector<std::unique_ptr<KBprojectorSparse>> g_ions;

// the driver does this once in main (src/main.cc:77)
void init_mpi_once()
{
    static bool done = false;
    static std::ostringstream sink;
    if (!done) MGmol_MPI::setup(MPI_COMM_WORLD, sink);
    done = true;
}
}

extern "C"
{

int refkb_orbdtype_bytes() { return (int)sizeof(ORBDTYPE); }

// Mesh (one rank, subdivx = 1) + one species read from `pseudo_file` the way
// Potentials::readAll and MGmol::initKBR do (src/Potentials.cc:545-548,
// src/MGmol.cc:1004).  info: nlradius, dim_nl, max_l, llocal, projectors per ion.
int refkb_setup(const int gdims[3], const double origin[3], const double ll[3], int lap_type,
    const char* pseudo_file, char filter_flag, double* info)
{
    init_mpi_once();
    g_ions.clear();
    delete g_species;
    g_species = nullptr;
    Mesh::deleteInstance();
    const unsigned ngpts[3] = { (unsigned)gdims[0], (unsigned)gdims[1], (unsigned)gdims[2] };
    Mesh::setup(MPI_COMM_WORLD, ngpts, origin, ll, lap_type);
    const pb::Grid& grid = Mesh::instance()->grid();
    std::ifstream probe(pseudo_file);
    if (!probe.good()) return -1;
    probe.close();
    // Species::Species prints one line to MPIdata::sout on pe0: keep it off stdout
    std::ostream* keep = MPIdata::sout;
    std::ostringstream sink;
    MPIdata::sout = &sink;
    g_species     = new Species(MPI_COMM_WORLD);
    g_species->read_1species(pseudo_file);
    g_species->set_dim_nl(grid.hmin());
    g_species->set_dim_l(grid.hmin());
    g_species->initPotentials(filter_flag, grid.hmax(), false);
    MPIdata::sout = keep;
    if (info)
    {
        KBprojectorSparse probe_kb(*g_species);
        info[0] = g_species->nlradius();
        info[1] = g_species->dim_nl();
        info[2] = g_species->max_l();
        info[3] = g_species->llocal();
        info[4] = probe_kb.nProjectors();
    }
    return 0;
}

// Ion::setup path for the projector: KBprojectorSparse(species) + setup(centre)
int refkb_add_ion(const double center[3])
{
    if (!g_species) return -1;
    g_ions.emplace_back(new KBprojectorSparse(*g_species));
    g_ions.back()->setup(center);
    return (int)g_ions.size() - 1;
}

int refkb_ion_size(int j) { return g_ions[j]->size_nl_[0]; }
int refkb_ion_nproj(int j) { return g_ions[j]->nProjectors(); }
int refkb_ion_single(int j) { return g_ions[j]->onlyOneProjector() ? 1 : 0; }

// node list, value arrays in getProjectors order (nproj x size_nl), kbcoeff * sign
int refkb_ion_data(int j, int* nlindex, ORBDTYPE* proj, double* coeff)
{
    KBprojectorSparse& kb = *g_ions[j];
    const int n           = kb.size_nl_[0];
    for (int i = 0; i < n; i++)
        nlindex[i] = kb.nlindex_[0][i];
    std::vector<const KBPROJDTYPE*> projectors;
    if (n > 0) kb.getProjectors(0, projectors);
    std::vector<short> signs;
    std::vector<double> coeffs;
    kb.getKBsigns(signs);
    kb.getKBcoeffs(coeffs);
    const int np = kb.nProjectors();
    if ((int)signs.size() != np || (int)coeffs.size() != np) return -1;
    if (n > 0 && (int)projectors.size() != np) return -2;
    for (int p = 0; p < np; p++)
    {
        coeff[p] = coeffs[p] * signs[p];
        for (int i = 0; i < n; i++)
            proj[(size_t)p * n + i] = projectors[p][i];
    }
    return 0;
}

// KBPsiMatrixInterface::computeLocalElement for one ion and one function:
// out[i] = vel * <projector i | psi>
int refkb_psi(int j, const ORBDTYPE* psi, double* out)
{
    KBprojectorSparse& kb = *g_ions[j];
    const int np          = kb.nProjectors();
    if (!kb.overlaps(0))
    {
        for (int i = 0; i < np; i++)
            out[i] = 0.;
        return 0;
    }
    kb.registerPsi(0, psi);
    const double vel = Mesh::instance()->grid().vel();
    for (short i = 0; i < np; i++)
        out[i] = vel * kb.dotPsi(0, i);
    return 0;
}

// get_vnlpsi's loop for one function (kbpsi_rows: the projections of that function, ion after
// ion, projector after projector) and, add != 0, the caller's MPaxpy into hpsi
int refkb_vnlpsi(const double* kbpsi_rows, ORBDTYPE* out, int add)
{
    const int numpt = Mesh::instance()->numpt();
    std::vector<ORBDTYPE> vpsi(numpt);
    memset(vpsi.data(), 0, numpt * sizeof(ORBDTYPE));
    int row = 0;
    for (size_t j = 0; j < g_ions.size(); j++)
    {
        KBprojectorSparse& kb = *g_ions[j];
        std::vector<short> signs;
        std::vector<double> kbcoeffs;
        kb.getKBsigns(signs);
        kb.getKBcoeffs(kbcoeffs);
        const short nprojs = kb.nProjectors();
        if (kb.onlyOneProjector())
        {
            const double coeff = kbpsi_rows[row] * kbcoeffs[0] * signs[0];
            kb.axpySKet(0, coeff, vpsi.data());
        }
        else
        {
            std::vector<double> coeff;
            for (short i = 0; i < nprojs; i++)
                coeff.push_back(kbpsi_rows[row + i] * kbcoeffs[i] * signs[i]);
            kb.axpyKet(0, coeff, vpsi.data());
        }
        row += nprojs;
    }
    if (add)
        LinearAlgebraUtils<MemorySpace::Host>::MPaxpy(numpt, 1., vpsi.data(), out);
    else
        memcpy(out, vpsi.data(), numpt * sizeof(ORBDTYPE));
    return row;
}

} // extern "C"
