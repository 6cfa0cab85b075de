// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// Curvilinear (coords::diagonal_coords) companion of ref_driver.cc: the reference's own CPU
// pde_algs::flux_div(basic) on stretched grids, for the CONVECTIVE functors only.
//
// Why a second library: the unmodified reference cannot instantiate flux_div on diagonal_coords — the
// parameter `const typename idx_t::value_type::value_type& idir` of calc_normal_vector
// (reference src/core/coord_system.h:255,274) names a type that does not exist for cell/face indices, so
// info::metric (src/omni/infos/info_metric.h:31) fails to compile. oracle/Makefile builds this file against a
// TEMPORARY copy of that one header in which the two declarations read `const int& idir` (a sed one-liner, the copy
// is deleted after the build; nothing else of the reference is touched and no reference source enters this
// repository). With that repair the convective path works: calc_jacobian (coord_system.h:295-302) at the
// computational cell centre (flux_div_basic.h:49-50), calc_normal_vector (coord_system.h:250-267) at the MAPPED
// position of each stencil cell (info_metric.h:31 passes grid.get_coords(idx), i.e. physical coordinates, into
// coord_deriv — kept exactly as the reference does it).
// The viscous / sensor path has no reference implementation on general coordinates
// (src/omni/infos/info_gradient.h:83 static_asserts coords::identity): not available here, "parity unpinned".
//
// Output: oracle/_ref/libspade_ref_curv.so (git-ignored, travels to the GPU box prebuilt).

#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <cstdlib>
#include <new>

#include "spade.h"

// deterministic domain-boundary flags: see ref_driver.cc
void* operator new(std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { return ::operator new(n); }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

using real_t = double;
using prim_t = spade::fluid_state::prim_t<real_t>;
using cons_t = spade::fluid_state::cons_t<real_t>;
using flux_t = spade::fluid_state::flux_t<real_t>;

extern "C"
{
    struct ref_cfg            // identical to ref_driver.cc
    {
        int    nblocks[3];
        int    ncells[3];
        int    ng;
        double bounds[6];
        int    periodic[3];
        int    scheme;        // 3 totani_lr, 5 fweno_t, 7 cent_keep<4>   (convective functors only)
        double gamma, R, mu, prandtl, sensor_eps;
        int    nranks;
        int    integrator;    // 0 rk4_t, 1 ssprk3_opt, 2 ssprk3_t, 3 rk2_t
        double sgs_cw, sgs_delta, sgs_prt;   // (unused here)
    };
    // one 1-D mapping per direction (reference src/core/coord_system.h:54-177)
    //   kind 0 identity_1D; 1 scaled_coord_1D(par[0]); 2 integrated_tanh_1D(par[0..3] = y0, y1, inflation, rate); 3 quad_1D
    struct ref_coords
    {
        int    kind[3];
        double par[3][4];
    };
}

namespace
{
    std::string g_last_error;

    // run-time choice between the reference's own 1-D mappings (one template instantiation of the grid)
    struct switch_1D
    {
        typedef real_t coord_type;
        int kind = 0;
        spade::coords::identity_1D<real_t>        idn;
        spade::coords::scaled_coord_1D<real_t>    scl;
        spade::coords::integrated_tanh_1D<real_t> tnh;
        spade::coords::quad_1D<real_t>            qud;
        real_t map(const real_t& x) const
        {
            switch (kind) { case 1: return scl.map(x); case 2: return tnh.map(x); case 3: return qud.map(x); default: return idn.map(x); }
        }
        real_t coord_deriv(const real_t& x) const
        {
            switch (kind) { case 1: return scl.coord_deriv(x); case 2: return tnh.coord_deriv(x); case 3: return qud.coord_deriv(x); default: return idn.coord_deriv(x); }
        }
    };
    switch_1D make_1d(const ref_coords& cd, int d)
    {
        switch_1D m;
        m.kind = cd.kind[d];
        if (m.kind == 1) m.scl = spade::coords::scaled_coord_1D<real_t>(cd.par[d][0]);
        if (m.kind == 2) m.tnh = spade::coords::integrated_tanh_1D<real_t>(cd.par[d][0], cd.par[d][1], cd.par[d][2], cd.par[d][3]);
        return m;
    }
    using coords_t = spade::coords::diagonal_coords<switch_1D, switch_1D, switch_1D>;

    template <typename func_t>
    void with_scheme(const ref_cfg& c, const func_t& func)
    {
        spade::fluid_state::ideal_gas_t<real_t> air(c.gamma, c.R);
        spade::convective::totani_lr tscheme(air);
        spade::convective::fweno_t<decltype(air)> wscheme(air);
        switch (c.scheme)
        {
            case 3: { func(tscheme); break; }
            case 5: { func(wscheme); break; }
            case 7: { func(spade::convective::cent_keep<4>(air)); break; }
            default: throw std::runtime_error("ref_driver_curv: convective schemes 3, 5, 7 only (the reference has no gradient on general coordinates)");
        }
    }

    template <typename func_t>
    void with_setup(const ref_cfg& c, const ref_coords& cd, const func_t& func)
    {
        int argc = 0; char** argv = nullptr;
        std::vector<int> devices(c.nranks, 0);
        spade::parallel::compute_env_t env(&argc, &argv, devices);
        env.exec([&](spade::parallel::pool_t& pool)
        {
            spade::ctrs::array<int, 3> num_blocks(c.nblocks[0], c.nblocks[1], c.nblocks[2]);
            spade::ctrs::array<int, 3> cells_in_block(c.ncells[0], c.ncells[1], c.ncells[2]);
            spade::ctrs::array<int, 3> exchange_cells(c.ng, c.ng, c.ng);
            spade::bound_box_t<real_t, 3> bounds;
            for (int d = 0; d < 3; ++d) { bounds.min(d) = c.bounds[2*d]; bounds.max(d) = c.bounds[2*d+1]; }
            coords_t coords(make_1d(cd, 0), make_1d(cd, 1), make_1d(cd, 2));
            spade::ctrs::array<bool, 3> periodic(bool(c.periodic[0]), bool(c.periodic[1]), bool(c.periodic[2]));
            spade::grid::cartesian_blocks_t blocks(num_blocks, bounds);
            spade::grid::cartesian_grid_t grid(cells_in_block, blocks, coords, pool);
            prim_t fill1 = 0.0;
            flux_t fill2 = 0.0;
            spade::grid::grid_array prim(grid, fill1, exchange_cells, spade::device::cpu);
            spade::grid::grid_array rhs (grid, fill2, exchange_cells, spade::device::cpu);
            auto handle = spade::grid::make_exchange(prim, periodic);
            const std::size_t per_block = prim.data.size()/std::max<std::size_t>(1, grid.get_num_local_blocks());
            std::size_t first_glob = grid.get_num_local_blocks() > 0
                ? grid.get_partition().to_global(spade::utils::tag[spade::partition::local](std::size_t(0))).value : 0;
            func(pool, grid, prim, rhs, handle, first_glob*per_block, prim.data.size());
        });
    }

    template <typename func_t>
    int guarded(const func_t& f)
    {
        try { f(); return 0; }
        catch (const std::exception& e) { g_last_error = e.what(); return 1; }
        catch (...) { g_last_error = "unknown exception"; return 2; }
    }
}

extern "C"
{
    const char* refc_last_error() { return g_last_error.c_str(); }

    // the 1-D mapping and its derivative as the reference evaluates them
    double refc_map(const ref_coords* cd, int d, double x)   { return make_1d(*cd, d).map(x); }
    double refc_deriv(const ref_coords* cd, int d, double x) { return make_1d(*cd, d).coord_deriv(x); }

    // rhs (+)= flux_div(q) on the stretched grid; q must have its ghosts filled by the caller
    int refc_flux_div(const ref_cfg* c, const ref_coords* cd, const double* q, double* rhs_io, int increment)
    {
        return guarded([&]
        {
            with_setup(*c, *cd, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                std::copy(rhs_io + off, rhs_io + off + cnt, rhs.data.begin());
                with_scheme(*c, [&](const auto& flux_func)
                {
                    if (increment) spade::pde_algs::flux_div(prim, rhs, flux_func, spade::algs::make_traits(spade::pde_algs::basic, spade::pde_algs::increment));
                    else           spade::pde_algs::flux_div(prim, rhs, flux_func, spade::algs::make_traits(spade::pde_algs::basic, spade::pde_algs::overwrite));
                });
                std::copy(rhs.data.begin(), rhs.data.end(), rhs_io + off);
            });
        });
    }

    // What the reference's geometry hands the flux functors, for every cell of global block `lb` (rank 0 of 1):
    //   jac[(k*n1 + j)*n0 + i]                    calc_jacobian at the computational cell centre (interior cells)
    //   nrm[((k*n1 + j)*n0 + i)*3 + dir]          info::metric of the CELL (i,j,k) for direction dir (its dir-component)
    //   xyz[((k*n1 + j)*n0 + i)*3 + d]            physical coordinates of the cell centre
    int refc_geometry(const ref_cfg* c, const ref_coords* cd, int64_t lb, double* jac, double* nrm, double* xyz)
    {
        return guarded([&]
        {
            with_setup(*c, *cd, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                if (pool.rank() != 0) return;
                const auto geom = grid.image(spade::device::cpu);
                for (int k = 0; k < c->ncells[2]; ++k)
                for (int j = 0; j < c->ncells[1]; ++j)
                for (int i = 0; i < c->ncells[0]; ++i)
                {
                    spade::grid::cell_idx_t ic(i, j, k, (int)lb);
                    const std::size_t o = (std::size_t(k)*c->ncells[1] + j)*c->ncells[0] + i;
                    const auto xc = geom.get_comp_coords(ic);
                    jac[o] = spade::coords::calc_jacobian(geom.get_coord_sys(), xc, ic);
                    const auto xp = geom.get_coords(ic);
                    for (int d = 0; d < 3; ++d)
                    {
                        xyz[3*o + d] = xp[d];
                        const auto n = spade::coords::calc_normal_vector(geom.get_coord_sys(), xp, ic, d);
                        nrm[3*o + d] = n[d];
                    }
                }
            });
        });
    }

    // nsteps of integrator_t::advance() with bc = exchange, rhs = flux_div(basic, overwrite) on the stretched grid
    int refc_advance(const ref_cfg* c, const ref_coords* cd, double* q, double dt, int nsteps)
    {
        return guarded([&]
        {
            with_setup(*c, *cd, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                spade::fluid_state::ideal_gas_t<real_t> air(c->gamma, c->R);
                with_scheme(*c, [&](const auto& flux_func)
                {
                    auto bc = [&](auto& qq, const auto& t) { handle.exchange(qq, pool); };
                    auto calc_rhs = [&](auto& rr, const auto& qq, const auto& t)
                    {
                        spade::pde_algs::flux_div(qq, rr, flux_func, spade::algs::make_traits(spade::pde_algs::basic, spade::pde_algs::overwrite));
                    };
                    cons_t transform_state;
                    spade::fluid_state::state_transform_t trans(transform_state, air);
                    spade::time_integration::time_axis_t axis(real_t(0.0), real_t(dt));
                    auto run = [&](const auto& alg)
                    {
                        spade::time_integration::integrator_data_t qd(std::move(prim), std::move(rhs), alg);
                        spade::time_integration::integrator_t ti(axis, alg, qd, calc_rhs, bc, trans);
                        pool.sync();
                        for (int n = 0; n < nsteps; ++n) ti.advance();
                        pool.sync();
                        const auto& sol = ti.solution();
                        std::copy(sol.data.begin(), sol.data.end(), q + off);
                    };
                    switch (c->integrator)
                    {
                        case 0: { run(spade::time_integration::rk4_t());   break; }
                        case 2: { run(spade::time_integration::ssprk3_t()); break; }
                        default: throw std::runtime_error("ref_driver_curv: integrator 0 or 2");
                    }
                });
            });
        });
    }
}
