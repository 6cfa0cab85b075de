// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// C-callable driver around the UNMODIFIED SPADE reference headers (/root/reference/src, included
// at build time by oracle/Makefile; no reference source is copied into this repository).
// Output: oracle/_ref/libspade_ref.so (git-ignored, travels to the GPU box prebuilt).
//
// It exposes the reference's own CPU implementation of the hot path
//   pde_algs::flux_div (basic)      reference: src/pde-algs/flux-div/flux_div_basic.h:17-77
//   arr_exchange_t::exchange        reference: src/grid/make_exchange.h:111-410
//   integrate_advance (fused RK)    reference: src/time-integration/advance.h:57-102,236-280,286-402
//   algs::transform_reduce          reference: src/algs/transform_reduce.h:53-191
//   exchange_config_t tables        reference: src/grid/exchange_config.h:286-419
// on plain double buffers in the reference's own memory order
//   off(v,i,j,k,lb) = v + 5*((i+g) + (n0+2g)*((j+g) + (n1+2g)*((k+g) + (n2+2g)*lb)))
// (reference: src/core/mem_map.h:484-496). "Ranks" are std::threads of one process
// (reference: src/parallel/compute_pool.h:497-514) with arrays on device::cpu; rank r owns a
// contiguous run of global blocks, so the global buffer is the concatenation of rank buffers.

#include <cstdint>
#include <cstring>
#include <chrono>
#include <vector>
#include <mutex>
#include <string>
#include <cstdlib>
#include <new>

#include "spade.h"

// The reference never initialises the "not a domain boundary" flags: bound_box_t<bool,3>() leaves
// its storage unset and cartesian_blocks_t only ever writes `true` (reference:
// src/core/bounding_box.h:20, src/grid/cartesian_blocks.h:53,58-65), so the non-periodic pruning
// in src/grid/exchange_config.h:336-342 reads heap garbage and the reference's exchange tables
// vary from run to run. Zero-filling every heap block handed to the reference (replaceable
// operator new, bound to this library only by -Wl,-Bsymbolic) gives the evidently intended
// "false" and a deterministic reference. Nothing in the reference itself is changed.
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
    struct ref_cfg
    {
        int    nblocks[3];
        int    ncells[3];
        int    ng;            // exchange (ghost) cells, same in each direction
        double bounds[6];     // xmin xmax ymin ymax zmin zmax
        int    periodic[3];
        int    scheme;        // see with_scheme
        double gamma, R, mu, prandtl, sensor_eps;
        int    nranks;        // number of thread-ranks
        int    integrator;    // 0 = rk4_t fused prim/cons, 1 = ssprk3_opt, 2 = ssprk3_t fused, 3 = rk2_t fused
        double sgs_cw, sgs_delta, sgs_prt;   // subgrid_scale::wale_t(gas, cw, delta, prt) of schemes 11, 12
    };
}

extern "C"
{
    // the other half of a channel run's boundary callback and its forcing (same layout as spo_bc)
    struct ref_bc
    {
        int    mask[6];      // xmin xmax ymin ymax zmin zmax
        int    kind;         // 0: mirror kernel ghost[v] = a[v]*image[v] + b[v]; 1: boundary::extrapolate<order> (order 1 or 2)
        int    order;
        double a[5], b[5];
        int    use_normal;
        double a_normal;
        double force[3];
    };
}

namespace
{
    std::string g_last_error;
    // refinement passes (lists of global block ids) applied by with_setup; empty = uniform cartesian lattice
    std::vector<std::vector<std::size_t>> g_amr_refine;

    // algs::boundary_fill (reference src/grid/boundary_fill.h:32-133) with the kernels of ref_bc
    template <typename arr_t>
    void apply_boundary_fill(arr_t& prim, const ref_bc& bc)
    {
        spade::boundary::identifier_t which(bool(bc.mask[0]), bool(bc.mask[1]), bool(bc.mask[2]), bool(bc.mask[3]), bool(bc.mask[4]), bool(bc.mask[5]));
        if (bc.kind == 1)
        {
            if (bc.order == 1)      spade::algs::boundary_fill(prim, which, spade::boundary::extrapolate<1>);
            else if (bc.order == 2) spade::algs::boundary_fill(prim, which, spade::boundary::extrapolate<2>);
            else throw std::runtime_error("ref_driver: extrapolation order 1 or 2");
            return;
        }
        const ref_bc b = bc;
        auto kern = [b](const prim_t& q_image, const int idir)
        {
            prim_t g;
            for (int v = 0; v < 5; ++v) g[v] = b.a[v]*q_image[v] + b.b[v];
            if (b.use_normal) g[2+idir] = b.a_normal*q_image[2+idir] + b.b[2+idir];
            return g;
        };
        spade::algs::boundary_fill(prim, which, kern);
    }

    // pde_algs::source_term (reference src/pde-algs/source_term.h:25-51) with a body force
    template <typename arr_t, typename rhs_t>
    void apply_source_term(const arr_t& prim, rhs_t& rhs, const ref_bc& bc)
    {
        const real_t fx = bc.force[0], fy = bc.force[1], fz = bc.force[2];
        auto src = [=](const prim_t& q)
        {
            flux_t out;
            out.continuity() = 0.0;
            out.energy()     = fx*q.u() + fy*q.v() + fz*q.w();
            out.x_momentum() = fx;
            out.y_momentum() = fy;
            out.z_momentum() = fz;
            return out;
        };
        spade::pde_algs::source_term(prim, rhs, src);
    }

    template <typename func_t>
    void with_scheme(const ref_cfg& c, const func_t& func)
    {
        spade::fluid_state::ideal_gas_t<real_t> air(c.gamma, c.R);
        spade::viscous_laws::constant_viscosity_t<real_t> vlaw(c.mu, c.prandtl);
        spade::convective::totani_lr tscheme(air);
        spade::convective::fweno_t<decltype(air)> wscheme(air);
        spade::viscous::visc_lr vscheme(vlaw, air);
        spade::state_sensor::ducros_t<real_t> ducr(c.sensor_eps);
        switch (c.scheme)
        {
            case 0: { func(spade::omni::compose(tscheme, vscheme)); break; }
            case 1:
            {
                spade::convective::hybrid_scheme_t hyb(tscheme, wscheme, ducr, spade::convective::full_flux);
                func(spade::omni::compose(hyb, vscheme));
                break;
            }
            case 2: { func(spade::omni::compose(spade::convective::cent_keep<4>(air), vscheme)); break; }
            case 3: { func(tscheme); break; }
            case 4: { func(vscheme); break; }
            case 5: { func(wscheme); break; }
            case 6:
            {
                spade::convective::hybrid_scheme_t hyb(spade::convective::cent_keep<4>(air), wscheme, ducr, spade::convective::full_flux);
                func(spade::omni::compose(hyb, vscheme));
                break;
            }
            case 7: { func(spade::convective::cent_keep<4>(air)); break; }
            case 8:
            {
                spade::convective::hybrid_scheme_t hyb(tscheme, wscheme, ducr, spade::convective::diss_flux);
                func(spade::omni::compose(hyb, vscheme));
                break;
            }
            case 13: { func(spade::omni::compose(spade::convective::cent_keep<6>(air), vscheme)); break; }
            case 14: { func(spade::omni::compose(spade::convective::cent_keep<8>(air), vscheme)); break; }
            case 15: { func(spade::convective::cent_keep<6>(air)); break; }
            case 16: { func(spade::convective::cent_keep<8>(air)); break; }
            case 11:
            {
                // LES closure: visc_lr over sgs_visc_t(constant_viscosity_t, wale_t) (viscous_laws.h:175-216, subgrid_scale.h:25-91)
                spade::subgrid_scale::wale_t eddy(air, real_t(c.sgs_cw), real_t(c.sgs_delta), real_t(c.sgs_prt));
                spade::viscous_laws::sgs_visc_t slaw(vlaw, eddy);
                spade::viscous::visc_lr svisc(slaw, air);
                func(spade::omni::compose(tscheme, svisc));
                break;
            }
            case 12:
            {
                spade::subgrid_scale::wale_t eddy(air, real_t(c.sgs_cw), real_t(c.sgs_delta), real_t(c.sgs_prt));
                spade::viscous_laws::sgs_visc_t slaw(vlaw, eddy);
                spade::viscous::visc_lr svisc(slaw, air);
                spade::convective::hybrid_scheme_t hyb(tscheme, wscheme, ducr, spade::convective::full_flux);
                func(spade::omni::compose(hyb, svisc));
                break;
            }
            case 9:
            {
                spade::convective::rusanov_t rus(air);
                func(spade::convective::weno_t(rus));
                break;
            }
            case 10:
            {
                spade::convective::rusanov_t rus(air);
                spade::convective::weno_t wrus(rus);
                spade::convective::hybrid_scheme_t hyb(tscheme, wrus, ducr, spade::convective::full_flux);
                func(spade::omni::compose(hyb, vscheme));
                break;
            }
            case 17: { func(spade::convective::fweno_t<decltype(air), spade::convective::disable_smooth>(air)); break; }
            case 18:
            {
                spade::convective::rusanov_t rus(air);
                func(spade::convective::weno_t<decltype(rus), spade::convective::disable_smooth>(rus));
                break;
            }
            case 19:
            {
                spade::convective::fweno_t<decltype(air), spade::convective::disable_smooth> wlin(air);
                spade::convective::hybrid_scheme_t hyb(tscheme, wlin, ducr, spade::convective::full_flux);
                func(spade::omni::compose(hyb, vscheme));
                break;
            }
            default: throw std::runtime_error("ref_driver: unknown scheme id");
        }
    }

    // Runs func(pool, grid, prim, rhs, handle, off, cnt) on every thread-rank, where
    // [off, off+cnt) is the rank's slice (in doubles) of the global buffer.
    template <typename func_t>
    void with_setup(const ref_cfg& c, const func_t& func)
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
            spade::coords::identity<real_t> coords;
            spade::ctrs::array<bool, 3> periodic(bool(c.periodic[0]), bool(c.periodic[1]), bool(c.periodic[2]));
            auto body = [&](auto& grid)
            {
                prim_t fill1 = 0.0;
                flux_t fill2 = 0.0;
                spade::grid::grid_array prim(grid, fill1, exchange_cells, spade::device::cpu);
                spade::grid::grid_array rhs (grid, fill2, exchange_cells, spade::device::cpu);
                auto handle = spade::grid::make_exchange(prim, periodic);
                const std::size_t per_block = prim.data.size()/std::max<std::size_t>(1, grid.get_num_local_blocks());
                std::size_t first_glob = grid.get_num_local_blocks() > 0
                    ? grid.get_partition().to_global(spade::utils::tag[spade::partition::local](std::size_t(0))).value : 0;
                func(pool, grid, prim, rhs, handle, first_glob*per_block, prim.data.size());
            };
            if (g_amr_refine.empty())
            {
                spade::grid::cartesian_blocks_t blocks(num_blocks, bounds);
                spade::grid::cartesian_grid_t grid(cells_in_block, blocks, coords, pool);
                body(grid);
            }
            else
            {
                // AMR (reference src/amr/amr_blocks.h, src/grid/cartesian_grid.h:331-368): the grid is built on the unrefined
                // tree and refined before any array exists; every listed global block is split in all directions
                spade::amr::amr_blocks_t blocks(num_blocks, bounds);
                spade::grid::cartesian_grid_t grid(cells_in_block, blocks, coords, pool);
                using refine_t = typename decltype(blocks)::refine_type;
                for (const auto& pass: g_amr_refine)
                    grid.refine_blocks(pass, periodic, refine_t{true, true, true}, spade::amr::constraints::factor2);
                body(grid);
            }
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
    const char* ref_last_error() { return g_last_error.c_str(); }

    // doubles in the global padded array (uniform lattice)
    int64_t ref_array_size(const ref_cfg* c)
    {
        int64_t n = 5;
        for (int d = 0; d < 3; ++d) n *= (int64_t)(c->ncells[d] + 2*c->ng)*c->nblocks[d];
        return n;
    }

    // AMR: subsequent calls build the grid on amr_blocks_t and apply these refinement passes (npass lists, each
    // counts[p] global block ids, concatenated in `ids`); npass = 0 returns to the uniform lattice
    void ref_set_amr(int npass, const int64_t* counts, const int64_t* ids)
    {
        g_amr_refine.clear();
        for (int p = 0; p < npass; ++p)
        {
            g_amr_refine.emplace_back(ids, ids + counts[p]);
            ids += counts[p];
        }
    }

    // number of global blocks and their bounding boxes [nblocks][6] (global block order) of the current grid
    int ref_block_boxes(const ref_cfg* c, int64_t* nblocks, double* boxes, int64_t cap)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                if (pool.rank() != 0) return;
                const std::size_t n = grid.get_num_global_blocks();
                *nblocks = (int64_t)n;
                for (std::size_t lb = 0; lb < n && (int64_t)lb < cap; ++lb)
                {
                    const auto& bx = grid.get_blocks().get_block_box(lb);
                    for (int d = 0; d < 3; ++d) { boxes[6*lb + 2*d] = bx.min(d); boxes[6*lb + 2*d + 1] = bx.max(d); }
                }
            });
        });
    }

    // Interpolation (fine <-> coarse) transaction tables of rank `rank`: patch_fill_t lists send_data[1] / recv_data[1]
    // (reference src/grid/transactions.h:136-234, get_transaction.h:100-263). 26 int64 per transaction: the 16 fields of
    // ref_exchange_tables (source box = donor region), then dest.size(0..2), i_coeff[3], i_incr[3], 0.
    // offs: per peer the 6 intrp_offsets entries in the order of ref_exchange_tables.
    int ref_interp_tables(const ref_cfg* c, int rank, int64_t* out_send, int64_t* out_recv, int64_t cap,
        int64_t* n_send, int64_t* n_recv, int64_t* offs)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                if (pool.rank() != rank) return;
                using namespace spade::udci;
                const auto& cfg = handle.config;
                const auto dump = [&](const auto& list, int64_t* out, int64_t* n)
                {
                    *n = (int64_t)list.size();
                    int64_t idx = 0;
                    for (const auto& pf: list)
                    {
                        if (idx >= cap) break;
                        const auto& tr = pf.patches;
                        int64_t* o = out + 26*idx;
                        o[0] = int64_t(pf.tag); o[1] = tr.rank_send; o[2] = tr.rank_recv;
                        o[3] = int64_t(tr.glob_source_blk); o[4] = int64_t(tr.glob_dest_blk);
                        for (int d = 0; d < 4; ++d) o[5+d]  = tr.source.min(d);
                        for (int d = 0; d < 3; ++d) o[9+d]  = tr.source.size(d);
                        for (int d = 0; d < 4; ++d) o[12+d] = tr.dest.min(d);
                        for (int d = 0; d < 3; ++d) o[16+d] = tr.dest.size(d);
                        for (int d = 0; d < 3; ++d) o[19+d] = pf.i_coeff[d];
                        for (int d = 0; d < 3; ++d) o[22+d] = pf.i_incr[d];
                        o[25] = 0;
                        ++idx;
                    }
                };
                dump(cfg.send_data[1_c], out_send, n_send);
                dump(cfg.recv_data[1_c], out_recv, n_recv);
                const auto& io = cfg.intrp_offsets;
                for (int p = 0; p < c->nranks; ++p)
                {
                    offs[6*p+0] = int64_t(io.send_message_size[p]);
                    offs[6*p+1] = int64_t(io.recv_message_size[p]);
                    offs[6*p+2] = int64_t(io.send_rank_offsets[p]);
                    offs[6*p+3] = int64_t(io.send_rank_sizes[p]);
                    offs[6*p+4] = int64_t(io.recv_rank_offsets[p]);
                    offs[6*p+5] = int64_t(io.recv_rank_sizes[p]);
                }
            });
        });
    }

    // rhs (+)= flux_div(q); q must have its ghosts filled by the caller
    int ref_flux_div(const ref_cfg* c, const double* q, double* rhs_io, int increment)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
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

    // ghost exchange in place
    int ref_exchange(const ref_cfg* c, double* q)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                pool.sync();
                handle.exchange(prim, pool);
                pool.sync();
                std::copy(prim.data.begin(), prim.data.end(), q + off);
            });
        });
    }

    // max over interior cells of sqrt(gamma R T) + |u|   (the CFL wavespeed reduction)
    int ref_reduce_umax(const ref_cfg* c, const double* q, double* out)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                const real_t gam = c->gamma, rgas = c->R;
                auto get_u = [=](const prim_t& val)
                {
                    return sqrt(gam*rgas*val.T()) + sqrt(val.u()*val.u() + val.v()*val.v() + val.w()*val.w());
                };
                const auto reduc = spade::algs::make_reduction(prim, get_u, spade::algs::max);
                const real_t umax = spade::algs::transform_reduce(prim, reduc);
                if (pool.rank() == 0) *out = umax;
            });
        });
    }

    // io::output_vtk(out_dir, basename, prim) (io/io_vtk.h:276-288) and io::binary_write (io/io_native.h:40-47) of the array
    // holding q: the reference's own files, for the byte-for-byte check of the drop-in's writers (tests/test_io.py)
    int ref_output_vtk(const ref_cfg* c, const double* q, const char* out_dir, const char* basename)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                spade::io::output_vtk(std::string(out_dir), std::string(basename), prim);
            });
        });
    }

    // nsteps of integrator_t::advance() with bc = exchange, rhs = flux_div(basic, overwrite).
    // q must enter with ghosts filled. If seconds != nullptr it receives the wall time of the
    // advance() loop (max over ranks, barriers on both sides).
    int ref_advance_channel(const ref_cfg* c, const ref_bc* bcd, double* q, double dt, int nsteps, double* seconds);
    int ref_advance(const ref_cfg* c, double* q, double dt, int nsteps, double* seconds)
    {
        return ref_advance_channel(c, nullptr, q, dt, nsteps, seconds);
    }

    // the same with boundary = exchange + boundary_fill and rhs = flux_div + source_term (bcd != nullptr)
    int ref_advance_channel(const ref_cfg* c, const ref_bc* bcd, double* q, double dt, int nsteps, double* seconds)
    {
        return guarded([&]
        {
            std::mutex mut;
            double tmax = 0.0;
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                spade::fluid_state::ideal_gas_t<real_t> air(c->gamma, c->R);
                with_scheme(*c, [&](const auto& flux_func)
                {
                    auto bc = [&](auto& qq, const auto& t) { handle.exchange(qq, pool); if (bcd) apply_boundary_fill(qq, *bcd); };
                    auto calc_rhs = [&](auto& rr, const auto& qq, const auto& t)
                    {
                        spade::pde_algs::flux_div(qq, rr, flux_func, spade::algs::make_traits(spade::pde_algs::basic, spade::pde_algs::overwrite));
                        if (bcd) apply_source_term(qq, rr, *bcd);
                    };
                    cons_t transform_state;
                    spade::fluid_state::state_transform_t trans(transform_state, air);
                    spade::time_integration::time_axis_t axis(real_t(0.0), real_t(dt));
                    auto run = [&](const auto& alg)
                    {
                        spade::time_integration::integrator_data_t qd(std::move(prim), std::move(rhs), alg);
                        spade::time_integration::integrator_t ti(axis, alg, qd, calc_rhs, bc, trans);
                        pool.sync();
                        auto t0 = std::chrono::steady_clock::now();
                        for (int n = 0; n < nsteps; ++n) ti.advance();
                        pool.sync();
                        auto t1 = std::chrono::steady_clock::now();
                        const double dtw = std::chrono::duration<double>(t1 - t0).count();
                        { std::lock_guard<std::mutex> lk(mut); tmax = std::max(tmax, dtw); }
                        const auto& sol = ti.solution();
                        std::copy(sol.data.begin(), sol.data.end(), q + off);
                    };
                    switch (c->integrator)
                    {
                        case 0: { run(spade::time_integration::rk4_t());   break; }
                        case 1: { run(spade::time_integration::ssprk3_opt); break; }
                        case 2: { run(spade::time_integration::ssprk3_t()); break; }
                        case 3: { run(spade::time_integration::rk2_t());   break; }
                        case 4: { run(spade::time_integration::ssprk34_t()); break; }
                        case 5: { run(spade::time_integration::rk38r_t());   break; }
                        default: throw std::runtime_error("ref_driver: unknown integrator id");
                    }
                });
            });
            if (seconds) *seconds = tmax;
        });
    }

    // the same through the GENERIC integrate_advance (advance.h:109-230): no state transform (identity_transform), the array
    // itself is integrated; integrator: 0 rk4_t, 2 ssprk3_t, 3 rk2_t; high_storage: rk2hs_t / ssprk3hs_t
    int ref_advance_generic(const ref_cfg* c, double* q, double dt, int nsteps, int high_storage)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                with_scheme(*c, [&](const auto& flux_func)
                {
                    auto bc = [&](auto& qq, const auto& t) { handle.exchange(qq, pool); };
                    auto calc_rhs = [&](auto& rr, const auto& qq, const auto& t)
                    {
                        spade::pde_algs::flux_div(qq, rr, flux_func, spade::algs::make_traits(spade::pde_algs::basic, spade::pde_algs::overwrite));
                    };
                    spade::time_integration::time_axis_t axis(real_t(0.0), real_t(dt));
                    auto run = [&](const auto& alg)
                    {
                        spade::time_integration::integrator_data_t qd(std::move(prim), std::move(rhs), alg);
                        spade::time_integration::integrator_t ti(axis, alg, qd, calc_rhs, bc);
                        pool.sync();
                        for (int n = 0; n < nsteps; ++n) ti.advance();
                        pool.sync();
                        const auto& sol = ti.solution();
                        std::copy(sol.data.begin(), sol.data.end(), q + off);
                    };
                    const int id = c->integrator + 10*(high_storage ? 1 : 0);
                    switch (id)
                    {
                        case 0:  { run(spade::time_integration::rk4_t());      break; }
                        case 2:  { run(spade::time_integration::ssprk3_t());   break; }
                        case 3:  { run(spade::time_integration::rk2_t());      break; }
                        case 12: { run(spade::time_integration::ssprk3hs_t()); break; }
                        case 13: { run(spade::time_integration::rk2hs_t());    break; }
                        default: throw std::runtime_error("ref_driver: generic advance: integrator 0, 2, 3 (high storage: 2, 3)");
                    }
                });
            });
        });
    }

    // algs::boundary_fill in place (q enters with whatever ghosts the caller set)
    int ref_boundary_fill(const ref_cfg* c, const ref_bc* bc, double* q)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                apply_boundary_fill(prim, *bc);
                std::copy(prim.data.begin(), prim.data.end(), q + off);
            });
        });
    }

    // rhs += S(q)
    int ref_source_term(const ref_cfg* c, const ref_bc* bc, const double* q, double* rhs_io)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                std::copy(q + off, q + off + cnt, prim.data.begin());
                std::copy(rhs_io + off, rhs_io + off + cnt, rhs.data.begin());
                apply_source_term(prim, rhs, *bc);
                std::copy(rhs.data.begin(), rhs.data.end(), rhs_io + off);
            });
        });
    }

    // Injection transaction tables of rank `rank` out of c->nranks.
    // Each transaction is 16 int64: tag, rank_send, rank_recv, glob_src, glob_dst,
    //   src.min(0..3) [4], src.size(0..2) [3], dst.min(0..3) [4]
    // out_send/out_recv have capacity cap transactions; counts returned in n_send/n_recv.
    // offs receives, for the injection tables, per peer p in [0,nranks):
    //   send_message_size[p], recv_message_size[p], send_rank_offsets[p], send_rank_sizes[p],
    //   recv_rank_offsets[p], recv_rank_sizes[p]      (6*nranks int64)
    int ref_exchange_tables(const ref_cfg* c, int rank, int64_t* out_send, int64_t* out_recv, int64_t cap,
        int64_t* n_send, int64_t* n_recv, int64_t* offs)
    {
        return guarded([&]
        {
            with_setup(*c, [&](auto& pool, auto& grid, auto& prim, auto& rhs, auto& handle, std::size_t off, std::size_t cnt)
            {
                if (pool.rank() != rank) return;
                using namespace spade::udci;
                const auto& cfg = handle.config;
                const auto dump = [&](const auto& list, int64_t* out, int64_t* n)
                {
                    *n = (int64_t)list.size();
                    int64_t idx = 0;
                    for (const auto& tr: list)
                    {
                        if (idx >= cap) break;
                        int64_t* o = out + 16*idx;
                        o[0] = int64_t(tr.tag); o[1] = tr.rank_send; o[2] = tr.rank_recv;
                        o[3] = int64_t(tr.glob_source_blk); o[4] = int64_t(tr.glob_dest_blk);
                        for (int d = 0; d < 4; ++d) o[5+d]  = tr.source.min(d);
                        for (int d = 0; d < 3; ++d) o[9+d]  = tr.source.size(d);
                        for (int d = 0; d < 4; ++d) o[12+d] = tr.dest.min(d);
                        ++idx;
                    }
                };
                dump(cfg.send_data[0_c], out_send, n_send);
                dump(cfg.recv_data[0_c], out_recv, n_recv);
                const auto& io = cfg.injec_offsets;
                for (int p = 0; p < c->nranks; ++p)
                {
                    offs[6*p+0] = int64_t(io.send_message_size[p]);
                    offs[6*p+1] = int64_t(io.recv_message_size[p]);
                    offs[6*p+2] = int64_t(io.send_rank_offsets[p]);
                    offs[6*p+3] = int64_t(io.send_rank_sizes[p]);
                    offs[6*p+4] = int64_t(io.recv_rank_offsets[p]);
                    offs[6*p+5] = int64_t(io.recv_rank_sizes[p]);
                }
            });
        });
    }

    // prim -> cons -> prim round trip pieces, for pinning convert_state
    // (reference: src/navier-stokes/fluid_state.h:103-135)
    void ref_prim2cons(double gamma, double R, const double* p, double* w)
    {
        spade::fluid_state::ideal_gas_t<real_t> air(gamma, R);
        prim_t q; cons_t cc;
        for (int i = 0; i < 5; ++i) q[i] = p[i];
        spade::fluid_state::convert_state(q, cc, air);
        for (int i = 0; i < 5; ++i) w[i] = cc[i];
    }
    void ref_cons2prim(double gamma, double R, const double* w, double* p)
    {
        spade::fluid_state::ideal_gas_t<real_t> air(gamma, R);
        prim_t q; cons_t cc;
        for (int i = 0; i < 5; ++i) cc[i] = w[i];
        spade::fluid_state::convert_state(cc, q, air);
        for (int i = 0; i < 5; ++i) p[i] = q[i];
    }
}
