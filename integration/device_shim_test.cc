// Built against the reference's own headers and linked with BOTH the compiled reference
// (oracle/_ref/libmgmol_ref.so: pb::Grid, pb::PEenv and the Host kernels) and
// libmgmol_b200.so.  For every FD kernel of src/pb/FDkernels.h it calls the reference's Host
// overload and the shim's Device overload with the same pb::Grid and the same arguments and
// requires BIT-IDENTICAL results (the library's literal kernels reproduce the reference's
// arithmetic).  Exit 0 = ok, 77 = no CUDA device (nothing is computed on the CPU instead).
#include <mpi.h>

#include "mgmol_b200_device.h"
#include "PEenv.h"

#include <cmath>
#include <cstring>
#include <vector>

template <typename T>
static int run()
{
    const unsigned ngpts[3] = { 16, 12, 20 };
    const double origin[3]  = { 0., 0., 0. };
    const double lattice[3] = { 4., 3., 5. };
    int fails = 0;
    for (short ghosts = 1; ghosts <= 4; ghosts++)
    {
        pb::PEenv pe(MPI_COMM_WORLD, ngpts[0], ngpts[1], ngpts[2]);
        pb::Grid grid(origin, lattice, ngpts, pe, ghosts, 0);
        const size_t nfunc_all = 3, n = grid.sizeg() * nfunc_all;
        std::vector<T> v(n), host_out(n, (T)0), dev_out(n, (T)0);
        unsigned long long s = 12345 + ghosts;
        for (size_t i = 0; i < n; i++)
        {
            s    = s * 6364136223846793005ULL + 1442695040888963407ULL;
            v[i] = (T)((double)(s >> 11) / 9007199254740992.0 - 0.5);
        }
        using DevMem = MemorySpace::Memory<T, MemorySpace::Device>;
        T* v_dev = DevMem::allocate((unsigned)n);
        T* b_dev = DevMem::allocate((unsigned)n);
        DevMem::copy_view_to_dev(v.data(), (unsigned)n, v_dev);
        struct Case
        {
            const char* name;
            int min_ghosts;
            int which;
        } cases[] = { { "FDkernelDel2_2nd", 1, 0 }, { "FDkernelDel2_4th", 2, 1 },
            { "FDkernelDel2_4th_Mehr", 1, 2 }, { "FDkernelDel2_6th", 3, 3 },
            { "FDkernelDel2_8th", 4, 4 }, { "FDkernelRHS_4th_Mehr1", 1, 5 } };
        for (const Case& c : cases)
        {
            // the ghost width GridFactory gives the operator (src/GridFactory.h:23-51); the
            // Mehrstellen pair and the 2nd order operator also on wider grids
            if (ghosts < c.min_ghosts || (c.which >= 3 && c.which <= 4 && ghosts != c.min_ghosts)) continue;
            std::fill(host_out.begin(), host_out.end(), (T)0);
            DevMem::set(b_dev, (unsigned)n, 0);
            // The reference's batched 6th / 8th order kernels start the x offset once, outside
            // the loop over the functions (src/pb/FDkernels.cc:296 / :391: `int iix = gpt *
            // incx;` before `for (ifunc ...)`), so every function after the first is read and
            // written dim0 planes too far (out of bounds).  They are compared on one function;
            // the library applies the stencil to each function of the block.
            const size_t nfunc = (c.which == 3 || c.which == 4) ? 1 : nfunc_all;
            switch (c.which)
            {
                case 0:
                    pb::FDkernelDel2_2nd(grid, v.data(), host_out.data(), nfunc, MemorySpace::Host());
                    pb::FDkernelDel2_2nd(grid, v_dev, b_dev, nfunc, MemorySpace::Device());
                    break;
                case 1:
                    pb::FDkernelDel2_4th(grid, v.data(), host_out.data(), nfunc, MemorySpace::Host());
                    pb::FDkernelDel2_4th(grid, v_dev, b_dev, nfunc, MemorySpace::Device());
                    break;
                case 2:
                    pb::FDkernelDel2_4th_Mehr(grid, v.data(), host_out.data(), nfunc, MemorySpace::Host());
                    pb::FDkernelDel2_4th_Mehr(grid, v_dev, b_dev, nfunc, MemorySpace::Device());
                    break;
                case 3:
                    pb::FDkernelDel2_6th(grid, v.data(), host_out.data(), nfunc, MemorySpace::Host());
                    pb::FDkernelDel2_6th(grid, v_dev, b_dev, nfunc, MemorySpace::Device());
                    break;
                case 4:
                    pb::FDkernelDel2_8th(grid, v.data(), host_out.data(), nfunc, MemorySpace::Host());
                    pb::FDkernelDel2_8th(grid, v_dev, b_dev, nfunc, MemorySpace::Device());
                    break;
                default:
                    pb::FDkernelRHS_4th_Mehr1(
                        grid, v.data(), host_out.data(), ghosts, nfunc, MemorySpace::Host());
                    pb::FDkernelRHS_4th_Mehr1(grid, v_dev, b_dev, ghosts, nfunc, MemorySpace::Device());
            }
            T* out = dev_out.data();
            DevMem::copy_view_to_host(b_dev, (unsigned)n, out);
            const bool same = std::memcmp(host_out.data(), dev_out.data(), n * sizeof(T)) == 0;
            std::printf("%-24s %s ghosts %d  %s\n", c.name, sizeof(T) == 8 ? "f64" : "f32", ghosts,
                same ? "bit-identical" : "DIFFERS");
            if (!same) fails++;
        }
        DevMem::free(v_dev);
        DevMem::free(b_dev);
    }
    return fails;
}

int main(int argc, char** argv)
{
    MPI_Init(&argc, &argv);
    setvbuf(stdout, nullptr, _IONBF, 0);
    if (mgb_device_count() == 0)
    {
        std::printf("no CUDA device (mgmol_b200 has no CPU fallback)\n");
        return 77;
    }
    const int fails = run<double>() + run<float>();
    std::printf(fails ? "FAILED\n" : "device shim ok\n");
    MPI_Finalize();
    return fails ? 1 : 0;
}
