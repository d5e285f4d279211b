// mgmol_b200_poisson.hpp -- the Poisson solvers of the Hartree potential
//  above the C ABI: the host control flow of
//   pb::SolverLap<Lap<T>,T>::solve = pb::Mgm over pb::Vcycle
//       (src/pb/SolverLap.cc:62-72, src/pb/Mgm.h:21-112, src/pb/Vcycle.h:29-250)
//   PCGSolver<Lap<T>,T>::solve with its float multigrid preconditioner
//       (src/PCGSolver.cc:50-252)
// on one scalar field, every grid operation a C-ABI call with nfunc = 1.
// Operators Laph4M (0), Laph2 (1), Laph4 (2); boundary conditions 0 and 1;
// single-rank boxes.  Header-only, C++11, no CUDA headers.
//
// The solvers are templates over the field type.  GridFunc<T> below is the
// device field; the CPU tests instantiate the same solver code with a host
// field whose operations are the oracle's, to pin the control flow against the
// compiled reference without a GPU (tests/cpp/test_poisson_mirror.cc).
#ifndef MGMOL_B200_POISSON_HPP
#define MGMOL_B200_POISSON_HPP

#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "mgmol_b200.hpp"

namespace mgmol_b200
{

// pb::GridFunc<T> on the device, as far as the Poisson solvers use it: one
// ghosted function plus the reductions of src/pb/GridFunc.cc.
template <typename T>
class GridFunc
{
public:
    typedef T value_type;
    // a no-ghost array in this field's memory space (what solve() takes)
    typedef DeviceMemory<T> buffer_type;
    // the same kind of field in another precision
    template <typename U>
    struct rebind
    {
        typedef GridFunc<U> type;
    };
    explicit GridFunc(const Grid& grid) : gfv_(grid, 1), have_weights_(false)
    {
        gfv_.set_updated_boundaries(false);
    }
    const Grid& grid() const { return gfv_.grid(); }
    T* data() { return gfv_.data(); }
    void resetData() { gfv_.resetData(); }
    void set_updated_boundaries(const bool f) { gfv_.set_updated_boundaries(f); }
    bool updated_boundaries() const { return gfv_.updated_boundaries(); }
    // no-ghost values in device memory <-> this field (precision converted)
    template <typename T2>
    void assign(const T2* noghost_dev)
    {
        gfv_.assign(noghost_dev, grid().size());
    }
    template <typename T2>
    void getValues(T2* noghost_dev)
    {
        gfv_.getValues(noghost_dev, grid().size());
    }
    // GridFunc<T>(const GridFunc<T2>&): the values of a field of another precision
    template <class Field2>
    void assignFrom(Field2& other)
    {
        DeviceMemory<typename Field2::value_type> tmp(grid().size());
        other.getValues(tmp.data());
        assign(tmp.data());
    }
    void trade_boundaries() { gfv_.trade_boundaries(); }
    void applyLap(const int type, GridFunc<T>& out) { gfv_.applyLap(type, out.gfv_); }
    void applyRHS(const int type, GridFunc<T>& out) { gfv_.applyRHS(type, out.gfv_); }
    void axpy(const double alpha, GridFunc<T>& x) { gfv_.axpy(alpha, x.gfv_); }
    void scal(const double alpha) { gfv_.scal(alpha); }
    void restrict3D(GridFunc<T>& coarse) { gfv_.restrict3D(coarse.gfv_); }
    void extend3D(GridFunc<T>& coarse) { gfv_.extend3D(coarse.gfv_); }
    // copy constructor / operator=: every stored value and the boundary flag
    void copy_from(GridFunc<T>& o)
    {
        MGB_CHECK(mgb_copy_dev(data(), o.data(), grid().sizeg() * sizeof(T), nullptr));
        set_updated_boundaries(o.updated_boundaries());
    }
    // GridFunc::gdot (src/pb/GridFunc.cc:2745-2798): interior points minus the
    // first layer of every non-periodic direction, double sum
    double gdot(GridFunc<T>& o)
    {
        weights();
        MGB_CHECK(mgb_gfv_pointwise_product(dtype_of<T>::value, grid().c(), data(),
            wdot_.data(), tmp_.data(), 1, nullptr));
        return reduce(tmp_.data(), o.data());
    }
    // GridFunc::norm2 (src/pb/GridFunc.cc:2856-2861)
    double norm2() { return std::sqrt(gdot(*this) * grid().vel()); }
    // GridFunc::get_average (src/pb/GridFunc.cc:2888-2927)
    double get_average()
    {
        weights();
        return reduce(data(), winner_.data()) / (double)grid().size();
    }
    // GridFunc::operator+=(T): every stored value
    void add_scalar(const double alpha)
    {
        weights();
        MGB_CHECK(mgb_axpy(
            dtype_of<T>::value, grid().sizeg(), alpha, ones_.data(), data(), nullptr));
    }

private:
    double reduce(const T* a, const T* b)
    {
        double r = 0.;
        MGB_CHECK(mgb_dot(dtype_of<T>::value, grid().sizeg(), a, b, scalar_.data(), nullptr));
        scalar_.copy_to_host(&r, 1);
        return r;
    }
    // the index sets of the reductions as 0/1 fields, built once per field
    void weights()
    {
        if (have_weights_) return;
        const Grid& gr   = grid();
        const mgb_grid* c = gr.c();
        const int g       = c->ghosts;
        const size_t ny = c->dim[1] + 2 * g, nz = c->dim[2] + 2 * g;
        std::vector<double> wd(gr.sizeg(), 0.);
        std::vector<T> wi(gr.sizeg(), (T)0), one(gr.sizeg(), (T)1);
        int lo[3];
        for (int d = 0; d < 3; d++)
            lo[d] = (c->bc[d] != 1 && c->coord[d] == 0) ? 1 : 0;
        for (int ix = 0; ix < c->dim[0]; ix++)
            for (int iy = 0; iy < c->dim[1]; iy++)
                for (int iz = 0; iz < c->dim[2]; iz++)
                {
                    const size_t k = ((size_t)(ix + g) * ny + (iy + g)) * nz + (iz + g);
                    wi[k]          = (T)1;
                    if (ix >= lo[0] && iy >= lo[1] && iz >= lo[2]) wd[k] = 1.;
                }
        wdot_.allocate(wd.size());
        winner_.allocate(wi.size());
        ones_.allocate(one.size());
        tmp_.allocate(wi.size());
        scalar_.allocate(1);
        wdot_.copy_to_dev(wd.data(), wd.size());
        winner_.copy_to_dev(wi.data(), wi.size());
        ones_.copy_to_dev(one.data(), one.size());
        MGB_CHECK(mgb_stream_sync(nullptr));
        have_weights_ = true;
    }

    GridFuncVector<T> gfv_;
    bool have_weights_;
    DeviceMemory<double> wdot_, scalar_;
    DeviceMemory<T> winner_, ones_, tmp_;
};

namespace poisson_detail
{
inline short minGhosts(const int lap_type) { return lap_type == 2 ? 2 : 1; }
// Laph4 hands its coarse levels to Laph2 (USE_LOWER_ORDER, src/pb/Vcycle.h:14,
// 181-200; Laph4::getLowerOrderOp)
inline int lowerOrder(const int lap_type) { return lap_type == 2 ? 1 : lap_type; }
inline void check(const Grid& grid, const int lap_type, const char* who)
{
    const mgb_grid* c = grid.c();
    bool ok           = lap_type >= 0 && lap_type <= 2;
    for (int d = 0; d < 3; d++)
        ok = ok && (c->bc[d] == 0 || c->bc[d] == 1) && c->nproc[d] == 1;
    if (!ok)
    {
        std::fprintf(stderr,
            "%s: operators 0/1/2, boundary conditions 0/1 and single-rank boxes only\n", who);
        ::mgmol_b200::fail("precondition");
    }
}
inline bool canCoarsen(const Grid& grid, const int ghosts)
{
    for (int d = 0; d < 3; d++)
        if ((grid.dim(d) & 1) || grid.dim(d) < 2 * ghosts) return false;
    return true;
}
inline bool fullyPeriodic(const Grid& grid)
{
    const mgb_grid* c = grid.c();
    return c->bc[0] == 1 && c->bc[1] == 1 && c->bc[2] == 1;
}
// Lap::jacobi (src/pb/Lap.cc:26-37): W = A x - B; x += scale W
template <class Field>
void jacobi(const int lap_type, Field& x, Field& rhs, Field& w, const double scale)
{
    x.applyLap(lap_type, w);
    w.axpy(-1., rhs);
    x.axpy(scale, w);
    x.set_updated_boundaries(false);
    w.set_updated_boundaries(false);
}
template <class Field>
double jacobiScale(const Grid& grid, const int lap_type)
{
    return -1. * Lap<typename Field::value_type>(grid, lap_type).jacobiFactor();
}
}

// pb::SolverLap<T, T2> (src/pb/SolverLap.h:18-77)
template <class Field>
class PoissonMG
{
public:
    typedef typename Field::value_type T;
    PoissonMG(const Grid& grid, const int lap_type)
        : grid_(grid.with_ghosts(poisson_detail::minGhosts(lap_type))), type_(lap_type),
          nu1_(2), nu2_(2), max_sweeps_(10), tol_(1.e-16), max_nlevels_(10), nb_sweeps_(0),
          final_residual_(-1.), final_relative_residual_(-1.), residual_reduction_(-1.)
    {
        poisson_detail::check(grid, lap_type, "PoissonMG");
    }
    void setup(const short nu1, const short nu2, const short max_sweeps, const double tol,
        const short max_nlevels, const bool /*gather_coarse_level*/ = true)
    {
        nu1_ = nu1, nu2_ = nu2, max_sweeps_ = max_sweeps, tol_ = tol, max_nlevels_ = max_nlevels;
    }
    short getNbSweeps() const { return nb_sweeps_; }
    double getFinalResidual() const { return final_residual_; }
    double getFinalRelativeResidual() const { return final_relative_residual_; }
    double getResidualReduction() const { return residual_reduction_; }

    // SolverLap::solve: vh (in: initial guess, out: solution) and rho without
    // ghosts, in the memory space of Field (device pointers for GridFunc<T>)
    bool solve(T* vh, const T* rho)
    {
        Field gf_vh(grid_), gf_rho(grid_), res(grid_), rhs(grid_), lhs(grid_), work1(grid_);
        gf_vh.assign(vh);
        gf_rho.assign(rho);
        res.copy_from(gf_rho);
        if (type_ == 0)
            res.applyRHS(0, rhs); // Laph4M::rhs -> rhs_4th_Mehr1
        else
            rhs.copy_from(res); // FDoper::rhs: B = A
        const double inv_rhs_norm = 1. / rhs.norm2();
        double init_residual_norm = 1.;
        bool converged            = false;
        nb_sweeps_                = 0;
        for (short i = 0; i < max_sweeps_; i++)
        {
            gf_vh.applyLap(type_, lhs);
            res.copy_from(rhs); // res.diff(rhs, lhs)
            res.axpy(-1., lhs);
            const double res_norm = res.norm2();
            if (i == 0) init_residual_norm = res_norm;
            if (res_norm * inv_rhs_norm < tol_)
            {
                final_residual_          = res_norm;
                final_relative_residual_ = res_norm * inv_rhs_norm;
                converged                = true;
                break;
            }
            work1.resetData();
            vcycle(type_, work1, res, 0);
            nb_sweeps_++;
            gf_vh.axpy(1., work1);
        }
        if (!converged)
        {
            gf_vh.applyLap(type_, lhs);
            lhs.axpy(-1., rhs);
            final_residual_          = lhs.norm2();
            final_relative_residual_ = final_residual_ * inv_rhs_norm;
        }
        residual_reduction_ = final_residual_ / init_residual_norm;
        if (poisson_detail::fullyPeriodic(grid_))
            gf_vh.add_scalar(-gf_vh.get_average()); // GridFunc::average0
        gf_vh.getValues(vh);
        return converged;
    }

private:
    // pb::Vcycle, x = 0 on entry; depth = -level of the grid
    void vcycle(const int lap_type, Field& x, Field& rhs, const int depth)
    {
        using namespace poisson_detail;
        const Grid& grid   = x.grid();
        const short g      = grid.ghost_pt();
        const double scale = jacobiScale<Field>(grid, lap_type);
        Field res(grid);
        for (short i = 0; i < nu1_; i++)
            jacobi(lap_type, x, rhs, res, scale);
        if (depth < max_nlevels_ && canCoarsen(grid, g))
        {
            x.applyLap(lap_type, res);
            res.axpy(-1., rhs);
            const short gc = std::max<short>(g - 1, 1);
            const Grid coarse_same(grid.coarse_grid());
            const Grid coarse_grid(coarse_same.with_ghosts(gc));
            Field rcoarse(coarse_grid), ucoarse(coarse_grid);
            if (gc == g)
                res.restrict3D(rcoarse);
            else
            {
                Field tmp(coarse_same);
                res.restrict3D(tmp);
                rcoarse.assignFrom(tmp);
            }
            rcoarse.set_updated_boundaries(false);
            vcycle(lowerOrder(lap_type), ucoarse, rcoarse, depth + 1);
            if (gc == g)
                res.extend3D(ucoarse);
            else
            {
                Field tmp(coarse_same);
                tmp.assignFrom(ucoarse);
                res.extend3D(tmp);
            }
            x.axpy(-1., res);
        }
        for (short i = 0; i < nu2_; i++)
            jacobi(lap_type, x, rhs, res, scale);
    }

    Grid grid_;
    int type_;
    short nu1_, nu2_, max_sweeps_;
    double tol_;
    short max_nlevels_, nb_sweeps_;
    double final_residual_, final_relative_residual_, residual_reduction_;
};

// PCGSolver<T, ScalarType> (src/PCGSolver.h:20-106): Field = the solver's
// precision, PField = POISSONPRECONDTYPE (float in the reference)
template <class Field, class PField>
class PoissonPCG
{
public:
    typedef typename Field::value_type T;
    PoissonPCG(const Grid& grid, const int lap_type)
        : grid_(grid.with_ghosts(poisson_detail::minGhosts(lap_type))), type_(lap_type),
          maxiters_(10), nu1_(2), nu2_(2), tol_(1.e-16), max_nlevels_(10), nlevels_(0),
          final_residual_(-1.), residual_reduction_(-1.)
    {
        poisson_detail::check(grid, lap_type, "PoissonPCG");
        setupPrecon();
    }
    void setup(const short nu1, const short nu2, const short max_sweeps, const double tol,
        const short max_nlevels)
    {
        maxiters_ = max_sweeps, nu1_ = nu1, nu2_ = nu2, tol_ = tol, max_nlevels_ = max_nlevels;
        setupPrecon();
    }
    double getFinalResidual() const { return final_residual_; }
    double getResidualReduction() const { return residual_reduction_; }

    // PCGSolver::solve (src/PCGSolver.cc:165-252), A u = rho
    bool solve(T* vh, const T* rho)
    {
        Field gf_phi(grid_), gf_rhs(grid_), lhs(grid_), res(grid_);
        gf_phi.assign(vh);
        gf_rhs.assign(rho);
        gf_phi.applyLap(type_, lhs);
        res.copy_from(gf_rhs);
        res.axpy(-1., lhs);
        const double init_rnorm = res.norm2();
        if (init_rnorm < 1.e-24) return true;
        double rnorm = init_rnorm;
        PField prec_z(grid_), prec_res(grid_);
        prec_res.assignFrom(res);
        prec_z.resetData();
        preconSolve(prec_z, prec_res, 0);
        Field z(grid_), p(grid_), ap(grid_);
        z.assignFrom(prec_z);
        p.assignFrom(prec_z);
        double rtz     = res.gdot(z);
        bool converged = false;
        for (int k = 0; k < maxiters_; k++)
        {
            p.applyLap(type_, ap);
            const double ptap = p.gdot(ap);
            const double alp  = rtz / ptap;
            gf_phi.axpy(alp, p);
            res.axpy(-alp, ap);
            rnorm = res.norm2();
            if (rnorm <= tol_ * init_rnorm)
            {
                converged = true;
                break;
            }
            prec_z.resetData();
            prec_res.assignFrom(res);
            preconSolve(prec_z, prec_res, 0);
            z.assignFrom(prec_z);
            const double rtz_new = res.gdot(z);
            const double bet     = rtz_new / rtz;
            p.scal(bet);
            p.axpy(1., z);
            rtz = rtz_new;
        }
        final_residual_     = rnorm;
        residual_reduction_ = rnorm / init_rnorm;
        if (poisson_detail::fullyPeriodic(grid_)) gf_phi.add_scalar(-gf_phi.get_average());
        gf_phi.getValues(vh);
        return converged;
    }

private:
    // setupPrecon (src/PCGSolver.cc:50-110): the coarse grids keep the fine
    // grid's ghost width, Laph2 below the fine level
    void setupPrecon()
    {
        using namespace poisson_detail;
        const short g = grid_.ghost_pt();
        grids_.clear();
        work_.clear();
        rcoarse_.clear();
        newv_.clear();
        grids_.push_back(grid_);
        nlevels_ = max_nlevels_;
        for (short ln = 1; ln <= max_nlevels_; ln++)
        {
            if (!canCoarsen(grids_.back(), g))
            {
                nlevels_ = ln - 1;
                break;
            }
            grids_.push_back(grids_.back().coarse_grid());
        }
        for (size_t l = 0; l < grids_.size(); l++)
        {
            work_.push_back(std::unique_ptr<PField>(new PField(grids_[l])));
            if (l > 0)
            {
                rcoarse_.push_back(std::unique_ptr<PField>(new PField(grids_[l])));
                newv_.push_back(std::unique_ptr<PField>(new PField(grids_[l])));
            }
        }
    }
    // preconSolve (src/PCGSolver.cc:112-162)
    void preconSolve(PField& v, PField& f, const short level)
    {
        using namespace poisson_detail;
        const bool last    = level == nlevels_;
        const short ncycl  = last ? std::max<short>(4, nu1_ + nu2_) : nu1_;
        const int lt       = level == 0 ? type_ : 1;
        const double scale = jacobiScale<PField>(grids_[level], lt);
        PField& work       = *work_[level];
        for (short it = 0; it < ncycl; it++)
            jacobi(lt, v, f, work, scale);
        if (last) return;
        PField& rcoarse = *rcoarse_[level];
        PField& newv    = *newv_[level];
        work.restrict3D(rcoarse);
        rcoarse.set_updated_boundaries(false);
        newv.resetData();
        preconSolve(newv, rcoarse, level + 1);
        work.extend3D(newv);
        v.axpy(-1., work);
        for (short it = 0; it < nu2_; it++)
            jacobi(lt, v, f, work, scale);
        const mgb_grid* c = grid_.c();
        if (c->bc[0] != 1 || c->bc[2] != 1) v.trade_boundaries(); // as at :161
    }

    Grid grid_;
    int type_;
    int maxiters_;
    short nu1_, nu2_;
    double tol_;
    short max_nlevels_, nlevels_;
    double final_residual_, residual_reduction_;
    std::vector<Grid> grids_;
    std::vector<std::unique_ptr<PField>> work_, rcoarse_, newv_;
};

// Hartree<T> (src/Hartree.h:18-46, src/Hartree.cc:27-112) with the state of its Poisson
// base (src/Poisson.h:31-89): the Hartree potential, kept between calls as the next
// initial guess, and the integrals of vh against the charges.  Field = the potential's
// precision (POTDTYPE), RField = the charges' (RHODTYPE), Solver = PoissonMG<Field>
// (Hartree) or PoissonPCG<Field, PField> (Hartree_CG).  Boundary conditions 0 / 1: no
// multipole boundary values.
template <class Field, class RField, class Solver>
class Hartree
{
public:
    typedef typename Field::value_type T;
    typedef typename RField::value_type RT;
    Hartree(const Grid& grid, const int lap_type)
        : grid_(grid.with_ghosts(poisson_detail::minGhosts(lap_type))), solver_(grid, lap_type),
          vh_(grid.size()), Int_vhrho_(0.), Int_vhrhoc_(0.), Int_vhrho_old_(0.)
    {
        resetVh();
    }
    void setup(const short nu1, const short nu2, const short max_sweeps, const double tol,
        const short max_nlevels, const bool gather_coarse_level = true)
    {
        (void)gather_coarse_level;
        solver_.setup(nu1, nu2, max_sweeps, tol, max_nlevels);
    }
    // the potential without ghosts, in Field's memory space
    T* vh() { return vh_.data(); }
    void resetVh()
    {
        Field zero(grid_);
        zero.resetData();
        zero.getValues(vh_.data());
    }
    template <typename T2>
    void set_vh(const T2* vh)
    {
        Field f(grid_);
        f.assign(vh);
        f.getValues(vh_.data());
    }
    double IntVhRho() const { return Int_vhrho_; }
    double IntVhRhoc() const { return Int_vhrhoc_; }
    double IntVhRho_old() const { return Int_vhrho_old_; }
    double getFinalResidual() const { return solver_.getFinalResidual(); }
    double getResidualReduction() const { return solver_.getResidualReduction(); }
    Solver& solver() { return solver_; }

    // Hartree::solve (src/Hartree.cc:27-112): rhs = 4 pi (rho - rhoc) in the solver's
    // precision, the Poisson solve from the kept vh, the integrals.  rho, rhoc: no-ghost
    // arrays of the charge precision in Field's memory space.
    bool solve(const RT* rho, const RT* rhoc)
    {
        Int_vhrho_old_ = vhDot(rho);
        RField work_rho(grid_), gf_rhoc(grid_);
        work_rho.assign(rho);
        gf_rhoc.assign(rhoc);
        work_rho.axpy(-1., gf_rhoc); // work_rho -= rhoc
        Field rhs(grid_);
        rhs.assignFrom(work_rho); // GridFunc<POTDTYPE> rhs(work_rho)
        rhs.scal(4. * M_PI);      // Hartree units
        typename Field::buffer_type rhs_values(grid_.size());
        rhs.getValues(rhs_values.data());
        const bool conv = solver_.solve(vh_.data(), rhs_values.data());
        Int_vhrho_      = vhDot(rho);
        Int_vhrhoc_     = vhDot(rhoc);
        return conv;
    }

private:
    // vel * vh_->gdot(charge): the charge converted to vh's precision
    double vhDot(const RT* charge)
    {
        Field a(grid_), b(grid_);
        a.assign(vh_.data());
        b.assign(charge);
        return grid_.vel() * a.gdot(b);
    }

    Grid grid_;
    Solver solver_;
    typename Field::buffer_type vh_;
    double Int_vhrho_, Int_vhrhoc_, Int_vhrho_old_;
};

} // namespace mgmol_b200

#endif // MGMOL_B200_POISSON_HPP
