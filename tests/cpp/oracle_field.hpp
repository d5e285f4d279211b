// TEST INFRASTRUCTURE ONLY.  A host field with the interface the Poisson
// solver templates of include/mgmol_b200_poisson.hpp expect, every grid
// operation done by the oracle's C restatement (oracle/libmgmol_oracle.so) of
// the reference kernel.  It lets the CPU suite run the C++ solver code without
// a GPU and compare it with the compiled reference solver.
#ifndef MGB_TEST_ORACLE_FIELD_HPP
#define MGB_TEST_ORACLE_FIELD_HPP

#include <cmath>
#include <vector>

#include "mgmol_b200.hpp"

extern "C"
{
void orc_trade_boundaries_f64(const int dims[3], int g, const int bc[3], double* u, int nfunc);
void orc_trade_boundaries_f32(const int dims[3], int g, const int bc[3], float* u, int nfunc);
void orc_del2_4th_Mehr_f64(const int d[3], int g, const double h[3], const double* v, double* u, int n);
void orc_del2_4th_Mehr_f32(const int d[3], int g, const double h[3], const float* v, float* u, int n);
void orc_del2_2nd_f64(const int d[3], int g, const double h[3], const double* v, double* u, int n);
void orc_del2_2nd_f32(const int d[3], int g, const double h[3], const float* v, float* u, int n);
void orc_del2_4th_f64(const int d[3], int g, const double h[3], const double* v, double* u, int n);
void orc_del2_4th_f32(const int d[3], int g, const double h[3], const float* v, float* u, int n);
void orc_rhs_4th_Mehr1_f64(const int d[3], int g, const double* v, double* rhs, int rg, int n);
void orc_rhs_4th_Mehr1_f32(const int d[3], int g, const float* v, float* rhs, int rg, int n);
void orc_restrict3D_f64(const int fd[3], int g, const double* fine, double* coarse, int n);
void orc_restrict3D_f32(const int fd[3], int g, const float* fine, float* coarse, int n);
void orc_extend3D_f64(const int fd[3], int g, const double* coarse, double* fine, int n);
void orc_extend3D_f32(const int fd[3], int g, const float* coarse, float* fine, int n);
}

namespace orc
{
inline void trade(const int d[3], int g, const int bc[3], double* u) { orc_trade_boundaries_f64(d, g, bc, u, 1); }
inline void trade(const int d[3], int g, const int bc[3], float* u) { orc_trade_boundaries_f32(d, g, bc, u, 1); }
inline void lap(int t, const int d[3], int g, const double h[3], const double* v, double* u)
{
    if (t == 0) orc_del2_4th_Mehr_f64(d, g, h, v, u, 1);
    if (t == 1) orc_del2_2nd_f64(d, g, h, v, u, 1);
    if (t == 2) orc_del2_4th_f64(d, g, h, v, u, 1);
}
inline void lap(int t, const int d[3], int g, const double h[3], const float* v, float* u)
{
    if (t == 0) orc_del2_4th_Mehr_f32(d, g, h, v, u, 1);
    if (t == 1) orc_del2_2nd_f32(d, g, h, v, u, 1);
    if (t == 2) orc_del2_4th_f32(d, g, h, v, u, 1);
}
inline void rhs(const int d[3], int g, const double* v, double* u) { orc_rhs_4th_Mehr1_f64(d, g, v, u, g, 1); }
inline void rhs(const int d[3], int g, const float* v, float* u) { orc_rhs_4th_Mehr1_f32(d, g, v, u, g, 1); }
inline void restrict3D(const int d[3], int g, const double* f, double* c) { orc_restrict3D_f64(d, g, f, c, 1); }
inline void restrict3D(const int d[3], int g, const float* f, float* c) { orc_restrict3D_f32(d, g, f, c, 1); }
inline void extend3D(const int d[3], int g, const double* c, double* f) { orc_extend3D_f64(d, g, c, f, 1); }
inline void extend3D(const int d[3], int g, const float* c, float* f) { orc_extend3D_f32(d, g, c, f, 1); }
}

template <typename T>
class OracleField
{
public:
    typedef T value_type;
    typedef std::vector<T> buffer_type; // a no-ghost array in this field's memory space
    template <typename U>
    struct rebind
    {
        typedef OracleField<U> type;
    };
    explicit OracleField(const mgmol_b200::Grid& grid)
        : grid_(grid), u_(grid.sizeg(), (T)0), upd_(false)
    {
        const mgb_grid* c = grid_.c();
        g_                = c->ghosts;
        for (int d = 0; d < 3; d++)
            dims_[d] = c->dim[d], bc_[d] = c->bc[d], h_[d] = c->h[d];
    }
    const mgmol_b200::Grid& grid() const { return grid_; }
    std::vector<T>& raw() { return u_; }
    void resetData()
    {
        std::fill(u_.begin(), u_.end(), (T)0);
        upd_ = true;
    }
    void set_updated_boundaries(const bool f) { upd_ = f; }
    bool updated_boundaries() const { return upd_; }
    size_t at(int ix, int iy, int iz) const
    {
        return ((size_t)(ix + g_) * (dims_[1] + 2 * g_) + (iy + g_)) * (dims_[2] + 2 * g_)
               + (iz + g_);
    }
    template <typename T2>
    void assign(const T2* noghost)
    {
        std::fill(u_.begin(), u_.end(), (T)0);
        size_t k = 0;
        for (int ix = 0; ix < dims_[0]; ix++)
            for (int iy = 0; iy < dims_[1]; iy++)
                for (int iz = 0; iz < dims_[2]; iz++)
                    u_[at(ix, iy, iz)] = (T)noghost[k++];
        upd_ = false;
    }
    template <typename T2>
    void getValues(T2* noghost)
    {
        size_t k = 0;
        for (int ix = 0; ix < dims_[0]; ix++)
            for (int iy = 0; iy < dims_[1]; iy++)
                for (int iz = 0; iz < dims_[2]; iz++)
                    noghost[k++] = (T2)u_[at(ix, iy, iz)];
    }
    template <class Field2>
    void assignFrom(Field2& other)
    {
        std::vector<typename Field2::value_type> tmp(grid_.size());
        other.getValues(tmp.data());
        assign(tmp.data());
    }
    void trade_boundaries()
    {
        if (upd_) return;
        orc::trade(dims_, g_, bc_, u_.data());
        upd_ = true;
    }
    void applyLap(const int type, OracleField<T>& out)
    {
        trade_boundaries();
        orc::lap(type, dims_, g_, h_, u_.data(), out.u_.data());
        out.upd_ = false;
    }
    void applyRHS(const int, OracleField<T>& out)
    {
        trade_boundaries();
        orc::rhs(dims_, g_, u_.data(), out.u_.data());
        out.upd_ = false;
    }
    void axpy(const double alpha, OracleField<T>& x)
    {
        for (size_t i = 0; i < u_.size(); i++)
            u_[i] += (T)(alpha * (double)x.u_[i]);
        upd_ = upd_ && x.upd_;
    }
    void scal(const double alpha)
    {
        if (alpha == 1.) return;
        for (size_t i = 0; i < u_.size(); i++)
            u_[i] = alpha == 0. ? (T)0 : (T)(alpha * (double)u_[i]);
    }
    void restrict3D(OracleField<T>& coarse)
    {
        trade_boundaries();
        orc::restrict3D(dims_, g_, u_.data(), coarse.u_.data());
        coarse.upd_ = false;
    }
    void extend3D(OracleField<T>& coarse)
    {
        coarse.trade_boundaries();
        orc::extend3D(dims_, g_, coarse.u_.data(), u_.data());
        upd_ = false;
    }
    void copy_from(OracleField<T>& o)
    {
        u_   = o.u_;
        upd_ = o.upd_;
    }
    double gdot(OracleField<T>& o)
    {
        double s = 0.;
        for (int ix = (bc_[0] != 1); ix < dims_[0]; ix++)
            for (int iy = (bc_[1] != 1); iy < dims_[1]; iy++)
                for (int iz = (bc_[2] != 1); iz < dims_[2]; iz++)
                    s += (double)u_[at(ix, iy, iz)] * (double)o.u_[at(ix, iy, iz)];
        return s;
    }
    double norm2() { return std::sqrt(gdot(*this) * grid_.vel()); }
    double get_average()
    {
        double s = 0.;
        for (int ix = 0; ix < dims_[0]; ix++)
            for (int iy = 0; iy < dims_[1]; iy++)
                for (int iz = 0; iz < dims_[2]; iz++)
                    s += (double)u_[at(ix, iy, iz)];
        return s / (double)grid_.size();
    }
    void add_scalar(const double alpha)
    {
        for (size_t i = 0; i < u_.size(); i++)
            u_[i] += (T)alpha;
    }

private:
    mgmol_b200::Grid grid_;
    std::vector<T> u_;
    bool upd_;
    int g_, dims_[3], bc_[3];
    double h_[3];
};

#endif
