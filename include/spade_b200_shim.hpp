// spade_b200_shim.hpp — C++20 host side of the drop-in: include AFTER "spade.h".
//
// It keeps SPADE's own template / operator API for the RHS hot path and forwards it to the C ABI of
// libspade_b200.so (include/spade_b200.h), i.e. to hand-written sm_100a kernels. An existing solver changes
//   (1) the flux_div trait tag:  algs::make_traits(pde_algs::b200, pde_algs::overwrite)
//   (2) grid::make_exchange  ->  b200::make_exchange            (same handle.exchange(array, group) call)
//   (3) nothing for the time integrator: integrator_t::advance() picks the overload below for device::gpu arrays
//   (4) algs::transform_reduce(...) -> b200::transform_reduce(array, b200::wavespeed{gas}, algs::max)   (CFL)
// and compiles unchanged otherwise. There is NO CPU fallback: CPU arrays and functor types outside the
// implemented set are compile-time errors.
//
// Reference interfaces replaced (file:line in the reference's src/):
//   pde_algs::flux_div                      pde-algs/flux-div/flux_div.h:23-41, tags.h:24-32, pde_traits.h:12-20
//   grid::make_exchange / arr_exchange_t    grid/make_exchange.h:98-421, exchange_config.h:286-419
//   time_integration::integrate_advance     time-integration/advance.h:236-280 (fused prim/cons), 359-402 (ssprk3_opt)
//   detail::transform_advance_to            time-integration/advance.h:57-102
//   algs::transform_reduce                  algs/transform_reduce.h:43-191
// Errors: a nonzero C-ABI return becomes except::sp_exception, the reference's convention after a failed CUDA call
// (dispatch/execute.h:86-96).
#pragma once
#include <map>
#include <mutex>
#include <tuple>
#include <tuple>
#include <vector>
#include <string>
#include <type_traits>
#include "spade_b200.h"

namespace spade::b200
{
    inline void check(int rc, const char* what)
    {
        if (rc != 0) throw except::sp_exception(std::string("spade_b200: ") + what + " failed (" + std::to_string(rc) + "): " + spb_last_error());
    }

    template <typename array_t> constexpr bool on_gpu = device::is_gpu<typename array_t::device_type>;

    // ---------------------------------------------------------------- functor recognition (closed set, compile time)
    template <typename T> struct always_false : std::false_type {};

    template <typename gas_t> inline void fill_gas(spb_flux_desc& d, const gas_t& gas)
    {
        static_assert(std::same_as<gas_t, fluid_state::ideal_gas_t<typename gas_t::value_type>>, "spade_b200: only fluid_state::ideal_gas_t is implemented");
        d.gamma = gas.get_gamma(); d.R = gas.get_R();
    }

    template <typename F> inline void fill(spb_flux_desc&, const F&)
    {
        static_assert(always_false<F>::value, "spade_b200: this flux functor type is not in the implemented set "
            "(totani_lr, cent_keep<2|4|6|8>, fweno_t / weno_t<rusanov_t> with enable_smooth or disable_smooth, hybrid_scheme_t<central, fweno_t | weno_t<rusanov_t>, ducros_t>, visc_lr<constant_viscosity_t | sgs_visc_t<constant_viscosity_t, wale_t>>, omni::compose of those); "
            "there is no CPU fallback");
    }
    template <typename gas_t> inline void fill(spb_flux_desc& d, const convective::totani_lr<gas_t>& f)
    { d.conv = SPB_CONV_TOTANI; fill_gas(d, f.gas); }
    template <typename gas_t, int order> inline void fill(spb_flux_desc& d, const convective::cent_keep_scheme_t<gas_t, order>& f)
    {
        static_assert(order == 2 || order == 4 || order == 6 || order == 8, "spade_b200: cent_keep orders 2, 4, 6, 8");
        d.conv = (order == 2) ? SPB_CONV_TOTANI : (order == 4) ? SPB_CONV_CENT_KEEP4 : (order == 6) ? SPB_CONV_CENT_KEEP6 : SPB_CONV_CENT_KEEP8;
        fill_gas(d, f.gas);
    }
    // fweno_t / weno_t with enable_smooth (nonlinear weights) or disable_smooth (the linear weights, convective.h:301-305,397-401)
    template <typename gas_t, convective::weno_smooth_indicator sm> inline void fill(spb_flux_desc& d, const convective::fweno_t<gas_t, sm>& f)
    { d.conv = SPB_CONV_FWENO; d.weno_linear = (sm == convective::disable_smooth) ? 1 : 0; fill_gas(d, f.gas); }
    // weno_t<rusanov_t> (convective.h:256-333, flux_funcs.h:9-53): the same reconstruction as fweno_t on precomputed split
    // fluxes (the nonlinear weights are algebraically identical); runs on the fweno_t kernel, agrees to round-off
    template <typename gas_t, convective::weno_smooth_indicator sm> inline void fill(spb_flux_desc& d, const convective::weno_t<convective::rusanov_t<gas_t>, sm>& f)
    { d.conv = SPB_CONV_FWENO; d.weno_linear = (sm == convective::disable_smooth) ? 1 : 0; fill_gas(d, f.flux_func.gas); }
    template <typename s0_t, typename gas_t, convective::weno_smooth_indicator sm, typename float_t, typename tag_t>
    inline void fill(spb_flux_desc& d, const convective::hybrid_scheme_t<s0_t, convective::weno_t<convective::rusanov_t<gas_t>, sm>, state_sensor::ducros_t<float_t>, tag_t>& f)
    {
        fill(d, f.scheme0);
        d.weno_linear = (sm == convective::disable_smooth) ? 1 : 0;
        d.diss = SPB_DISS_FWENO;
        d.blend = tag_t::value ? SPB_BLEND_FULL_FLUX : SPB_BLEND_DISS_FLUX;
        d.sensor_eps = f.blender.epsilon;
    }
    template <typename s0_t, typename gas_t, convective::weno_smooth_indicator sm, typename float_t, typename tag_t>
    inline void fill(spb_flux_desc& d, const convective::hybrid_scheme_t<s0_t, convective::fweno_t<gas_t, sm>, state_sensor::ducros_t<float_t>, tag_t>& f)
    {
        fill(d, f.scheme0);
        d.weno_linear = (sm == convective::disable_smooth) ? 1 : 0;
        d.diss = SPB_DISS_FWENO;
        d.blend = tag_t::value ? SPB_BLEND_FULL_FLUX : SPB_BLEND_DISS_FLUX;
        d.sensor_eps = f.blender.epsilon;
    }
    template <typename float_t, typename gas_t>
    inline void fill(spb_flux_desc& d, const viscous::visc_lr<viscous_laws::constant_viscosity_t<float_t>, gas_t>& f)
    {
        d.visc = 1; d.mu = f.vlaw.visc; d.beta = f.vlaw.beta; d.prandtl_inv = f.vlaw.prandtl_inv; fill_gas(d, f.gas);
    }
    // LES closure: visc_lr<sgs_visc_t<constant_viscosity_t, wale_t>> (viscous_laws.h:175-216, subgrid_scale.h:25-91)
    template <typename float_t, typename wgas_t, typename gas_t>
    inline void fill(spb_flux_desc& d, const viscous::visc_lr<viscous_laws::sgs_visc_t<viscous_laws::constant_viscosity_t<float_t>, subgrid_scale::wale_t<float_t, wgas_t>>, gas_t>& f)
    {
        d.visc = 1; d.mu = f.vlaw.lam.visc; d.beta = f.vlaw.lam.beta; d.prandtl_inv = f.vlaw.lam.prandtl_inv; fill_gas(d, f.gas);
        d.sgs = SPB_SGS_WALE; d.sgs_cw = f.vlaw.turb.cw; d.sgs_delta = f.vlaw.turb.delta; d.sgs_prt = f.vlaw.turb.prt;
    }
    template <typename k0_t> inline void fill(spb_flux_desc& d, const omni::composite_kernel_t<k0_t>& f) { fill(d, f.kern); }
    template <typename k0_t, typename k1_t, typename... ks_t>
    inline void fill(spb_flux_desc& d, const omni::composite_kernel_t<k0_t, k1_t, ks_t...>& f) { fill(d, f.kern); fill(d, f.next); }

    template <typename F> inline spb_flux_desc flux_desc(const F& f)
    {
        spb_flux_desc d{};
        d.conv = SPB_CONV_NONE; d.diss = SPB_DISS_NONE; d.blend = SPB_BLEND_FULL_FLUX; d.visc = 0;
        d.gamma = 1.4; d.R = 287.15; d.prandtl_inv = 1.0;
        d.sgs = SPB_SGS_NONE; d.sgs_prt = 1.0;
        fill(d, f);
        return d;
    }

    // ---------------------------------------------------------------- grid handle (device image of grid_geometry_t)
    // one host thread per GPU (compute_pool.h:497-514): the cache is thread-local. An entry is keyed by the grid's address and
    // the exchange-cell counts and carries a fingerprint of the geometry (cells per block, block count, first and last block
    // box); a grid rebuilt at the same address (a stack grid in a resolution loop, an AMR grid refined in place) fails the
    // fingerprint and its handle is rebuilt. b200::invalidate(grid) drops the entries of a grid whose lifetime ends.
    struct grid_entry_t { spb_grid* h = nullptr; int nx[3] = {0, 0, 0}; std::size_t nlb = 0; double box[12] = {0}; };
    using grid_key_t = std::tuple<const void*, int, int, int>;
    inline std::map<grid_key_t, grid_entry_t>& grid_cache() { thread_local std::map<grid_key_t, grid_entry_t> cache; return cache; }
    template <typename grid_t> inline void invalidate(const grid_t& grid)
    {
        auto& cache = grid_cache();
        for (auto it = cache.begin(); it != cache.end();)
        {
            if (std::get<0>(it->first) == (const void*)&grid) { spb_grid_destroy(it->second.h); it = cache.erase(it); }
            else ++it;
        }
    }
    template <typename array_t> inline spb_grid* grid_handle(const array_t& arr)
    {
        auto& cache = grid_cache();
        const auto& grid = arr.get_grid();
        const auto ng = arr.get_num_exchange();
        const std::size_t nlb = grid.get_num_local_blocks();
        const grid_key_t key{(const void*)&grid, ng[0], ng[1], ng[2]};
        grid_entry_t fp;
        fp.nlb = nlb;
        for (int d = 0; d < 3; ++d) fp.nx[d] = grid.get_num_cells(d);
        if (nlb > 0)
        {
            const auto b0 = grid.get_bounding_box(utils::tag[partition::local](std::size_t(0)));
            const auto b1 = grid.get_bounding_box(utils::tag[partition::local](nlb - 1));
            for (int d = 0; d < 3; ++d) { fp.box[2*d] = b0.min(d); fp.box[2*d + 1] = b0.max(d); fp.box[6 + 2*d] = b1.min(d); fp.box[7 + 2*d] = b1.max(d); }
        }
        auto it = cache.find(key);
        if (it != cache.end())
        {
            const grid_entry_t& e = it->second;
            bool same = e.nlb == fp.nlb;
            for (int d = 0; d < 3; ++d) same = same && e.nx[d] == fp.nx[d];
            for (int i = 0; i < 12; ++i) same = same && e.box[i] == fp.box[i];
            if (same) return e.h;
            spb_grid_destroy(e.h);                                       // stale: another grid lived at this address
            cache.erase(it);
        }
        using coord_sys_t = typename array_t::grid_type::coord_sys_type;
        constexpr bool is_identity = std::same_as<coord_sys_t, coords::identity<typename array_t::grid_type::coord_type>>;
        static_assert(is_identity || coords::diagonal_coordinate_system<coord_sys_t>,
            "spade_b200: coords::identity or coords::diagonal_coords (dense coordinate systems are not implemented)");
        std::vector<double> bbox(6*nlb);
        for (std::size_t lb = 0; lb < nlb; ++lb)
        {
            const auto bx = grid.get_bounding_box(utils::tag[partition::local](lb));
            for (int d = 0; d < 3; ++d) { bbox[6*lb + 2*d] = bx.min(d); bbox[6*lb + 2*d + 1] = bx.max(d); }
        }
        const int nx[3] = {grid.get_num_cells(0), grid.get_num_cells(1), grid.get_num_cells(2)};
        const int g[3]  = {ng[0], ng[1], ng[2]};
        spb_grid* h = nullptr;
        check(spb_grid_create(&h, nx, g, (int64_t)nlb, bbox.data()), "spb_grid_create");
        if constexpr (!is_identity)
        {
            // coords::diagonal_coords: the separable geometry as three 1-D tables per block and direction, evaluated with the
            // reference's own mapping objects. `area` follows info::metric to the letter: coord_deriv at the MAPPED cell
            // centre (omni/infos/info_metric.h:31 passes grid.get_coords(idx)); `jac` follows calc_jacobian: coord_deriv
            // at the computational centre (flux_div_basic.h:49-50); positions as grid_geometry.h:57-70.
            const auto& cs = grid.get_coord_sys();
            std::vector<double> area[3], jac[3], face[3];
            auto fill_dir = [&](const int d, const auto& map1d)
            {
                const int np = nx[d] + 2*g[d];
                area[d].resize(nlb*np); jac[d].resize(nlb*np); face[d].resize(nlb*(np + 1));
                for (std::size_t lb = 0; lb < nlb; ++lb)
                {
                    const double lo = bbox[6*lb + 2*d], dx = (bbox[6*lb + 2*d + 1] - bbox[6*lb + 2*d])/nx[d];
                    for (int i = 0; i <= np; ++i)
                    {
                        double rf = double(i - g[d]) + 0.5; rf -= 0.5;
                        face[d][lb*(np + 1) + i] = map1d.coord_deriv(lo + rf*dx);
                        if (i == np) break;
                        const double xc = lo + (double(i - g[d]) + 0.5)*dx;
                        jac[d][lb*np + i]  = map1d.coord_deriv(xc);
                        area[d][lb*np + i] = map1d.coord_deriv(map1d.map(xc));
                    }
                }
            };
            fill_dir(0, cs.xcoord); fill_dir(1, cs.ycoord); fill_dir(2, cs.zcoord);
            spb_metric_desc md;
            for (int d = 0; d < 3; ++d) { md.area[d] = area[d].data(); md.jac[d] = jac[d].data(); md.face[d] = face[d].data(); }
            check(spb_grid_set_metric(h, &md), "spb_grid_set_metric");
        }
        fp.h = h;
        cache[key] = fp;
        return h;
    }

    template <typename array_t> inline double* dev_ptr(array_t& a) { return (double*)(&a.data[0]); }
    template <typename array_t> inline const double* dev_ptr(const array_t& a) { return (const double*)(&a.data[0]); }

    template <typename array_t> constexpr void require_supported_array()
    {
        static_assert(on_gpu<array_t>, "spade_b200: arrays must live on device::gpu (the CPU path is the reference itself)");
        static_assert(std::same_as<typename array_t::value_type, double>, "spade_b200: fp64 only");
        static_assert(array_t::alias_type::size() == 5, "spade_b200: 5-variable states (prim_t / cons_t / flux_t)");
        // the kernels read the reference's DEFAULT memory order (mem_map::linear_t, variable fastest; core/mem_map.h:467-497);
        // an array built with mem_map::tiled / tiled_small (mem_map.h:501-650) has another order and would be misread
        static_assert(std::same_as<typename array_t::mem_map_type, mem_map::linear_t<5>>,
            "spade_b200: arrays must use the default mem_map::linear memory map (tiled maps are not implemented)");
    }

    // ---------------------------------------------------------------- exchange
    // In-process multi-GPU (the reference's own model: one host thread per GPU inside one process, compute_pool.h:497-514):
    // every rank publishes the device pointers of its per-peer receive buffers in this process-wide table; a rank then packs
    // its message for peer p STRAIGHT INTO p's receive buffer over NVLink (spb_exchange_pack_peer, peer access enabled once per
    // thread) instead of the reference's pack -> cudaMemcpyPeer -> unpack (exchange_message.h:14-56, compute_pool.h:93-99).
    struct peer_slot_t { double* buf[2] = {nullptr, nullptr}; unsigned long long* flags = nullptr; double* stage = nullptr; };
    struct peer_table_t
    {
        std::mutex mut;
        std::map<std::tuple<int, int, int>, peer_slot_t> slot;     // (exchange id, owner rank, sending peer) -> the owner's receive side
    };
    inline peer_table_t& peer_table() { static peer_table_t t; return t; }

    // side stream (highest priority) + the two events that order it against the thread's main (legacy default) stream
    struct stream_pair_t
    {
        cudaStream_t side = nullptr; cudaEvent_t ev_main = nullptr, ev_side = nullptr;
        cudaEvent_t ev_b[2] = {nullptr, nullptr}, ev_i[2] = {nullptr, nullptr};      // boundary / interior kernel of stage s done (s & 1)
        void fork() { cudaEventRecord(ev_main, nullptr); cudaStreamWaitEvent(side, ev_main, 0); }      // side continues after main
        void join() { cudaEventRecord(ev_side, side);    cudaStreamWaitEvent(nullptr, ev_side, 0); }   // main continues after side
    };
    inline stream_pair_t& streams()
    {
        thread_local stream_pair_t sp;
        if (!sp.side)
        {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (cudaStreamCreateWithPriority(&sp.side, cudaStreamNonBlocking, hi) != cudaSuccess
                || cudaEventCreateWithFlags(&sp.ev_main, cudaEventDisableTiming) != cudaSuccess
                || cudaEventCreateWithFlags(&sp.ev_side, cudaEventDisableTiming) != cudaSuccess
                || cudaEventCreateWithFlags(&sp.ev_b[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&sp.ev_b[1], cudaEventDisableTiming) != cudaSuccess
                || cudaEventCreateWithFlags(&sp.ev_i[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&sp.ev_i[1], cudaEventDisableTiming) != cudaSuccess)
                throw except::sp_exception("spade_b200: could not create the side stream of the overlapped schedule");
        }
        return sp;
    }

    // Move-only owner of a plan and of this rank's receive buffers / arrival flags (grid::make_exchange's handle is a value
    // type too, make_exchange.h:98-110; copies of THIS handle would alias device memory, so they are not allowed).
    // Messages between the GPUs of the process (the reference's model: one host thread per GPU, compute_pool.h:497-514) do
    // not pass through the host: begin() packs each message STRAIGHT INTO the neighbour's receive buffer over NVLink
    // (spb_exchange_pack_peer, peer access enabled once per thread) and raises the neighbour's arrival flag behind it in
    // stream order (spb_flag_signal); finish() makes the stream wait on this rank's own flags (spb_flag_wait) and unpacks.
    // No host barrier, no spb_sync: the reference's three barriers per exchange (compute_pool.h:166,216,222) are gone. Two
    // receive buffers per neighbour, used alternately, are enough: a rank can only pack message s+2 after it has received
    // s+1, which its neighbour sends after having unpacked s. GPUs without peer access fall back to pack -> cudaMemcpyPeer
    // -> unpack behind host barriers, like the reference (exchange_message.h:14-56, compute_pool.h:93-99).
    template <typename array_t> struct arr_exchange_t
    {
        using grid_type = typename array_t::grid_type;
        spb_exchange* plan = nullptr;
        int rank = 0, size = 1, id = 0;
        std::vector<peer_slot_t> own;       // [peer]: where messages from `peer` land (this rank's GPU)
        std::vector<peer_slot_t> remote;    // [peer]: the slot of this rank on `peer`'s GPU
        std::vector<double*> sendstage;     // [peer]: local send buffer when `peer` cannot be written directly
        bool all_direct = true, wired = false;
        unsigned long long seq = 0;
        std::vector<std::pair<int64_t, int64_t>> runs_first, runs_second;    // rank-boundary block runs, and the rest
        bool runs_ready = false;

        arr_exchange_t() = default;
        arr_exchange_t(const arr_exchange_t&) = delete;
        arr_exchange_t& operator=(const arr_exchange_t&) = delete;
        arr_exchange_t(arr_exchange_t&& o) noexcept { swap(o); }
        arr_exchange_t& operator=(arr_exchange_t&& o) noexcept { if (this != &o) { release(); swap(o); } return *this; }
        ~arr_exchange_t() { release(); }
        void swap(arr_exchange_t& o) noexcept
        {
            std::swap(plan, o.plan); std::swap(rank, o.rank); std::swap(size, o.size); std::swap(id, o.id);
            own.swap(o.own); remote.swap(o.remote); sendstage.swap(o.sendstage);
            std::swap(all_direct, o.all_direct); std::swap(wired, o.wired); std::swap(seq, o.seq);
            runs_first.swap(o.runs_first); runs_second.swap(o.runs_second); std::swap(runs_ready, o.runs_ready);
        }
        void release() noexcept
        {
            if (wired)
            {
                auto& tab = peer_table();
                std::lock_guard<std::mutex> lk(tab.mut);
                for (int p = 0; p < size; ++p) tab.slot.erase({id, rank, p});
            }
            for (auto& sl: own) { for (auto* b: sl.buf) if (b) spb_dev_free(b); if (sl.flags) spb_dev_free(sl.flags); if (sl.stage) spb_dev_free(sl.stage); }
            for (auto* b: sendstage) if (b) spb_dev_free(b);
            own.clear(); remote.clear(); sendstage.clear(); wired = false;
            if (plan) { spb_exchange_destroy(plan); plan = nullptr; }
        }

        template <typename group_t> void wire(group_t& group)
        {
            own.assign(size, peer_slot_t{}); remote.assign(size, peer_slot_t{}); sendstage.assign(size, nullptr);
            auto& tab = peer_table();
            const int mydev = group.device_id();
            int direct_here = 1;
            for (int p = 0; p < size; ++p)
            {
                if (p == rank) continue;
                const int64_t nr = spb_exchange_recv_cells(plan, p), ns = spb_exchange_send_cells(plan, p);
                if (nr > 0)
                {
                    for (int par = 0; par < 2; ++par) check(spb_dev_alloc((void**)&own[p].buf[par], sizeof(double)*5*nr), "spb_dev_alloc (receive buffer)");
                    check(spb_dev_alloc((void**)&own[p].flags, 2*sizeof(unsigned long long)), "spb_dev_alloc (arrival flags)");
                }
                const int pdev = group.pid(p).device_id;
                int can = 1;
                if (pdev != mydev)
                {
                    cudaDeviceCanAccessPeer(&can, mydev, pdev);
                    if (can) { const cudaError_t e = cudaDeviceEnablePeerAccess(pdev, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0; cudaGetLastError(); }
                }
                if (!can) direct_here = 0;
                if (!can && ns > 0) check(spb_dev_alloc((void**)&sendstage[p], sizeof(double)*5*ns), "spb_dev_alloc (send staging)");
                std::lock_guard<std::mutex> lk(tab.mut);
                peer_slot_t pub = own[p];
                pub.stage = nullptr;
                tab.slot[{id, rank, p}] = pub;
            }
            // every rank must take the same path: direct only if EVERY pair of the group can be written directly
            all_direct = group.reduce(direct_here, [](const int a, const int b) { return a < b ? a : b; }) != 0;
            group.sync();
            for (int p = 0; p < size; ++p)
            {
                if (p == rank) continue;
                std::lock_guard<std::mutex> lk(tab.mut);
                auto it = tab.slot.find({id, p, rank});
                if (it != tab.slot.end()) remote[p] = it->second;
            }
            if (!all_direct)
            {
                // staged path: publish the send staging buffers so that the receivers can pull from them
                { std::lock_guard<std::mutex> lk(tab.mut); for (int p = 0; p < size; ++p) if (p != rank) tab.slot[{-1 - id, rank, p}].stage = sendstage[p]; }
                group.sync();
                std::lock_guard<std::mutex> lk(tab.mut);
                for (int p = 0; p < size; ++p) if (p != rank) { auto it = tab.slot.find({-1 - id, p, rank}); if (it != tab.slot.end()) own[p].stage = it->second.stage; }
            }
            group.sync();
            wired = true;
        }

        void block_runs(const int64_t nlb)
        {
            if (runs_ready) return;
            std::vector<unsigned char> mask((std::size_t)std::max<int64_t>(nlb, 1), 0);
            check(spb_exchange_boundary_blocks(plan, nlb, mask.data()), "spb_exchange_boundary_blocks");
            runs_first.clear(); runs_second.clear();
            for (int64_t b = 0; b < nlb;)
            {
                int64_t e = b;
                while (e < nlb && mask[e] == mask[b]) ++e;
                (mask[b] ? runs_first : runs_second).push_back({b, e});
                b = e;
            }
            runs_ready = true;
        }

        // first half of the off-rank exchange on a raw device buffer in the array's layout: the messages leave on `stream`
        template <typename group_t> void begin(const double* q, group_t& group, cudaStream_t stream)
        {
            if (size == 1) return;
            if (!wired) wire(group);
            ++seq;
            const int par = int(seq & 1ull);
            for (int p = 0; p < size; ++p)
            {
                if (p == rank || spb_exchange_send_cells(plan, p) == 0) continue;
                if (all_direct)
                {
                    check(spb_exchange_pack_peer(plan, q, p, remote[p].buf[par], stream), "spb_exchange_pack_peer");
                    check(spb_flag_signal(remote[p].flags + par, seq, stream), "spb_flag_signal");
                }
                else check(spb_exchange_pack(plan, q, p, sendstage[p], stream), "spb_exchange_pack");
            }
        }
        // second half: wait for the neighbours' messages and unpack them on `stream`
        template <typename group_t> void finish(double* q, group_t& group, cudaStream_t stream)
        {
            if (size == 1) return;
            const int par = int(seq & 1ull);
            if (!all_direct)
            {
                check(spb_sync(stream), "spb_sync");
                group.sync();                                           // every staging buffer has been written
            }
            for (int p = 0; p < size; ++p)
            {
                const int64_t nr = spb_exchange_recv_cells(plan, p);
                if (p == rank || nr == 0) continue;
                if (all_direct) check(spb_flag_wait(own[p].flags + par, seq, stream), "spb_flag_wait");
                else cudaMemcpyPeerAsync(own[p].buf[par], group.device_id(), own[p].stage, group.pid(p).device_id, sizeof(double)*5*nr, stream);
                check(spb_exchange_unpack(plan, q, p, own[p].buf[par], stream), "spb_exchange_unpack");
            }
            if (!all_direct)
            {
                check(spb_sync(stream), "spb_sync");
                group.sync();                                           // staging buffers may be overwritten by the next exchange
            }
        }
        template <typename group_t> void exchange_messages(double* q, group_t& group, cudaStream_t stream = nullptr)
        {
            begin(q, group, stream);
            finish(q, group, stream);
        }

        void exchange(array_t& array, typename grid_type::group_type& group)
        {
            require_supported_array<array_t>();
            if (group.size() != 1) begin(dev_ptr(array), group, nullptr);
            check(spb_exchange_local(plan, dev_ptr(array), nullptr), "spb_exchange_local");      // overlaps the NVLink transfers
            if (group.size() != 1) finish(dev_ptr(array), group, nullptr);
            check(spb_sync(nullptr), "spb_sync");                       // reference semantics: visible on return (execute.h:85)
        }
    };

    // builds the plan from SPADE's own exchange_config_t (so AMR-free block topologies of any kind carry over), and
    // cross-checks nothing: the tables ARE the reference's
    template <typename array_t>
    inline arr_exchange_t<array_t> make_exchange(array_t& array, ctrs::array<bool, array_t::dim()>& periodic)
    {
        require_supported_array<array_t>();
        using namespace spade::udci;
        auto config = grid::get_exchg_config(array, periodic);
        const auto flatten = [&](const auto& list)
        {
            std::vector<int64_t> out; out.reserve(16*list.size());
            for (const auto& tr: list)
            {
                out.push_back(int64_t(tr.tag)); out.push_back(tr.rank_send); out.push_back(tr.rank_recv);
                out.push_back(int64_t(tr.glob_source_blk)); out.push_back(int64_t(tr.glob_dest_blk));
                for (int d = 0; d < 4; ++d) out.push_back(tr.source.min(d));
                for (int d = 0; d < 3; ++d) out.push_back(tr.source.size(d));
                for (int d = 0; d < 4; ++d) out.push_back(tr.dest.min(d));
            }
            return out;
        };
        // AMR grids: the patch_fill_t lists (transactions.h:136-234) as 26-field records
        const auto flatten_interp = [&](const auto& list)
        {
            std::vector<int64_t> out; out.reserve(26*list.size());
            for (const auto& pf: list)
            {
                const auto& tr = pf.patches;
                out.push_back(int64_t(pf.tag)); out.push_back(tr.rank_send); out.push_back(tr.rank_recv);
                out.push_back(int64_t(tr.glob_source_blk)); out.push_back(int64_t(tr.glob_dest_blk));
                for (int d = 0; d < 4; ++d) out.push_back(tr.source.min(d));
                for (int d = 0; d < 3; ++d) out.push_back(tr.source.size(d));
                for (int d = 0; d < 4; ++d) out.push_back(tr.dest.min(d));
                for (int d = 0; d < 3; ++d) out.push_back(tr.dest.size(d));
                for (int d = 0; d < 3; ++d) out.push_back(pf.i_coeff[d]);
                for (int d = 0; d < 3; ++d) out.push_back(pf.i_incr[d]);
                out.push_back(0);
            }
            return out;
        };
        const auto send = flatten(config.send_data[0_c]);
        const auto recv = flatten(config.recv_data[0_c]);
        const auto isend = flatten_interp(config.send_data[1_c]);
        const auto irecv = flatten_interp(config.recv_data[1_c]);
        const auto& grid = array.get_grid();
        const auto ngv = array.get_num_exchange();
        const int nx[3] = {grid.get_num_cells(0), grid.get_num_cells(1), grid.get_num_cells(2)};
        const int ng[3] = {ngv[0], ngv[1], ngv[2]};
        arr_exchange_t<array_t> out;
        out.rank = grid.group().rank(); out.size = grid.group().size();
        thread_local int next_id = 0;                                   // every rank builds its handles in the same order
        out.id = next_id++;
        check(spb_exchange_create_from_tables(&out.plan, nx, ng, out.rank, out.size, send.data(), (int64_t)send.size()/16,
                                              recv.data(), (int64_t)recv.size()/16), "spb_exchange_create_from_tables");
        if (!isend.empty() || !irecv.empty())
            check(spb_exchange_add_interp(out.plan, isend.data(), (int64_t)isend.size()/26, irecv.data(), (int64_t)irecv.size()/26), "spb_exchange_add_interp");
        return out;
    }

    // ---------------------------------------------------------------- callbacks the integrator can recognise
    // The usual rhs callback of a SPADE solver, `[&](auto& rhs, const auto& q, const auto& t) { flux_div(q, rhs, f, traits); }`
    // (development/cuda-tgv/main.cc:174-186), and the usual periodic boundary callback `handle.exchange(q, pool)`, as named
    // types. Passed to integrator_t they behave exactly like the lambdas; because their types are known, integrate_advance
    // (below) runs flux_div + the RK stage update + the same-rank ghost exchange as ONE kernel per stage.
    template <typename flux_func_t> struct flux_div_rhs_t
    {
        flux_func_t flux_func;
        template <typename rhs_arr_t, typename sol_arr_t, typename time_t>
        void operator()(rhs_arr_t& rhs, const sol_arr_t& q, const time_t&) const
        {
            require_supported_array<sol_arr_t>();
            const spb_flux_desc d = flux_desc(flux_func);
            check(spb_flux_div(grid_handle(q), dev_ptr(q), dev_ptr(rhs), &d, 0, nullptr), "spb_flux_div");
            check(spb_sync(nullptr), "spb_sync");
        }
    };
    template <typename flux_func_t> inline flux_div_rhs_t<flux_func_t> flux_div_rhs(const flux_func_t& f) { return flux_div_rhs_t<flux_func_t>{f}; }

    // algs::boundary_fill on a raw device buffer in the layout of `arr` (defined below)
    template <typename arr_t> inline void boundary_fill_ptr(const arr_t& arr, double* ptr, const boundary::identifier_t& boundaries, const spb_bc_desc& d);
    struct mirror_kernel;
    inline spb_bc_desc mirror_desc(const mirror_kernel& k);

    // The boundary callback of a solver as a named type: `handle.exchange(q, pool);` and, for a wall-bounded solver, the
    // `algs::boundary_fill(q, boundaries, kern)` that follows it (bc = exchange + boundary_fill).
    template <typename handle_t, typename group_t> struct exchange_bc_t
    {
        handle_t* handle;
        group_t*  group;
        bool has_fill = false;
        boundary::identifier_t which{};
        spb_bc_desc bc{};
        template <typename sol_arr_t> void fill(const sol_arr_t& q, double* ptr) const { if (has_fill) boundary_fill_ptr(q, ptr, which, bc); }
        template <typename sol_arr_t, typename time_t>
        void operator()(sol_arr_t& q, const time_t&) const { handle->exchange(q, *group); fill(q, dev_ptr(q)); check(spb_sync(nullptr), "spb_sync"); }
    };
    template <typename handle_t, typename group_t> inline exchange_bc_t<handle_t, group_t> exchange_bc(handle_t& h, group_t& g) { return exchange_bc_t<handle_t, group_t>{&h, &g}; }
    template <typename handle_t, typename group_t>
    inline exchange_bc_t<handle_t, group_t> exchange_bc(handle_t& h, group_t& g, const boundary::identifier_t& boundaries, const mirror_kernel& kern)
    {
        exchange_bc_t<handle_t, group_t> out{&h, &g};
        out.has_fill = true; out.which = boundaries; out.bc = mirror_desc(kern);
        return out;
    }

    template <typename T> struct is_flux_div_rhs : std::false_type {};
    template <typename F> struct is_flux_div_rhs<flux_div_rhs_t<F>> : std::true_type {};
    template <typename T> struct is_exchange_bc : std::false_type {};
    template <typename H, typename G> struct is_exchange_bc<exchange_bc_t<H, G>> : std::true_type {};

    // scratch solution buffer of the fused stage kernels (q_out must differ from q_in), one per thread (= per GPU) and size
    inline double* scratch_buffer(std::size_t doubles)
    {
        thread_local std::map<std::size_t, double*> cache;
        auto it = cache.find(doubles);
        if (it != cache.end()) return it->second;
        double* p = nullptr;
        if (cudaMalloc((void**)&p, sizeof(double)*doubles) != cudaSuccess) throw except::sp_exception("spade_b200: cudaMalloc of the stage scratch buffer failed");
        cache[doubles] = p;
        return p;
    }

    // ---------------------------------------------------------------- boundary_fill / source_term
    // algs::boundary_fill(arr, boundaries, kern) for the kernels that are linear per variable (boundary_fill.h:104-129):
    // ghost[v] = a[v]*image[v] + b[v]; with use_normal the velocity component along the boundary normal uses a_normal
    struct mirror_kernel
    {
        double a[5] = {1, 1, 1, 1, 1}, b[5] = {0, 0, 0, 0, 0};
        bool use_normal = false; double a_normal = 1.0;
        static mirror_kernel noslip_isothermal(const double t_wall) { mirror_kernel k; k.a[1] = k.a[2] = k.a[3] = k.a[4] = -1.0; k.b[1] = 2.0*t_wall; return k; }
        static mirror_kernel noslip_adiabatic() { mirror_kernel k; k.a[2] = k.a[3] = k.a[4] = -1.0; return k; }
        static mirror_kernel symmetry() { mirror_kernel k; k.use_normal = true; k.a_normal = -1.0; return k; }
    };
    template <typename arr_t> inline void boundary_fill_ptr(const arr_t& arr, double* ptr, const boundary::identifier_t& boundaries, const spb_bc_desc& d)
    {
        require_supported_array<arr_t>();
        const auto& grid = arr.get_grid();
        const auto& geom = grid.geometry(partition::local);
        for (int ib = 0; ib < 6; ++ib)
        {
            if (!boundaries(ib/2, ib%2)) continue;
            std::vector<int64_t> blocks;
            for (const auto lb: geom.boundary_blocks[ib].data(device::cpu)) blocks.push_back((int64_t)lb);
            if (blocks.empty()) continue;
            check(spb_boundary_fill(grid_handle(arr), ptr, ib/2, ib%2, blocks.data(), (int64_t)blocks.size(), &d, nullptr), "spb_boundary_fill");
        }
    }
    template <typename arr_t> inline void boundary_fill_desc(arr_t& arr, const boundary::identifier_t& boundaries, const spb_bc_desc& d)
    {
        boundary_fill_ptr(arr, dev_ptr(arr), boundaries, d);
        check(spb_sync(nullptr), "spb_sync");
    }
    inline spb_bc_desc mirror_desc(const mirror_kernel& k)
    {
        spb_bc_desc d{};
        d.kind = SPB_BC_MIRROR;
        for (int v = 0; v < 5; ++v) { d.a[v] = k.a[v]; d.b[v] = k.b[v]; }
        d.use_normal = k.use_normal ? 1 : 0; d.a_normal = k.a_normal;
        return d;
    }
    template <typename arr_t> inline void boundary_fill(arr_t& arr, const boundary::identifier_t& boundaries, const mirror_kernel& k)
    {
        boundary_fill_desc(arr, boundaries, mirror_desc(k));
    }
    template <typename arr_t, const int order> inline void boundary_fill(arr_t& arr, const boundary::identifier_t& boundaries, const boundary::extrap_t<order>&)
    {
        spb_bc_desc d{};
        d.kind = SPB_BC_EXTRAP; d.order = order;
        boundary_fill_desc(arr, boundaries, d);
    }
    // pde_algs::source_term(q, rhs, func) for the body-force kernel S = (0, f.u, fx, fy, fz) (source_term.h:25-51)
    struct body_force { double f[3] = {0, 0, 0}; };
    template <typename sol_arr_t, typename rhs_arr_t> inline void source_term(const sol_arr_t& q, rhs_arr_t& rhs, const body_force& bf)
    {
        require_supported_array<sol_arr_t>();
        spb_source_desc d{};
        d.kind = SPB_SRC_BODY_FORCE;
        for (int i = 0; i < 3; ++i) d.f[i] = bf.f[i];
        check(spb_source_term(grid_handle(q), dev_ptr(q), dev_ptr(rhs), &d, nullptr), "spb_source_term");
        check(spb_sync(nullptr), "spb_sync");
    }

    // ---------------------------------------------------------------- reductions
    // algs::transform_reduce(array, make_reduction(array, f, op)) (transform_reduce.h:43-191) for the closed set of element
    // kernels of the C ABI, with op = algs::max or algs::sum; the cross-rank step is the group's own reduce (transform_reduce.h:171-190)
    template <typename gas_t> struct wavespeed { gas_t gas; };             // sqrt(gamma R T) + |u|   (CFL, cuda-tgv/main.cc:152-161)
    template <typename gas_t> struct kinetic_energy { gas_t gas; };        // 0.5 rho |u|^2
    struct variable { int ivar = 0; };                                     // q[ivar]
    struct abs_variable { int ivar = 0; };                                 // |q[ivar]|
    template <typename op_t> constexpr int reduce_op()
    {
        static_assert(std::same_as<op_t, algs::max_t> || std::same_as<op_t, algs::sum_t>, "spade_b200: transform_reduce with algs::max or algs::sum");
        return std::same_as<op_t, algs::max_t> ? SPB_RED_MAX : SPB_RED_SUM;
    }
    template <typename array_t, typename op_t>
    inline double reduce_impl(const array_t& q, const int fn, const int ivar, const double gamma, const double R, const op_t& op)
    {
        require_supported_array<array_t>();
        double out = 0.0;
        check(spb_reduce(grid_handle(q), dev_ptr(q), reduce_op<op_t>(), fn, ivar, gamma, R, &out, nullptr), "spb_reduce");
        return q.get_grid().group().reduce(out, [&](const double a, const double b) { return double(op(a, b)); });
    }
    template <typename array_t, typename gas_t, typename op_t>
    inline double transform_reduce(const array_t& q, const wavespeed<gas_t>& f, const op_t& op)
    { return reduce_impl(q, SPB_FN_WAVESPEED, 0, f.gas.get_gamma(), f.gas.get_R(), op); }
    template <typename array_t, typename gas_t, typename op_t>
    inline double transform_reduce(const array_t& q, const kinetic_energy<gas_t>& f, const op_t& op)
    { return reduce_impl(q, SPB_FN_KINETIC, 0, f.gas.get_gamma(), f.gas.get_R(), op); }
    template <typename array_t, typename op_t>
    inline double transform_reduce(const array_t& q, const variable& f, const op_t& op) { return reduce_impl(q, SPB_FN_VAR, f.ivar, 1.4, 287.15, op); }
    template <typename array_t, typename op_t>
    inline double transform_reduce(const array_t& q, const abs_variable& f, const op_t& op) { return reduce_impl(q, SPB_FN_ABSVAR, f.ivar, 1.4, 287.15, op); }

    // ---------------------------------------------------------------- checkpoints and visualisation files
    // io::binary_write / io::binary_read (io/io_native.h:18-56): a headerless file in which global block lb occupies the bytes
    // [lb*B, (lb+1)*B), B = bytes of one padded block in the array's own memory order. The device layout IS that order, so a
    // rank's share is one contiguous slab: one cudaMemcpy and one pwrite / pread per rank. Files are interchangeable with the
    // reference's (same bytes) — a SPADE checkpoint restarts here and the other way round.
    template <typename array_t> inline void binary_write(const std::string& filename, const array_t& arr)
    {
        require_supported_array<array_t>();
        const auto& grid = arr.get_grid();
        auto& group = grid.group();
        const std::size_t nlb = grid.get_num_local_blocks(), total = arr.data.size();
        const std::size_t per_block = nlb ? total/nlb : 0;
        std::vector<double> host(total);
        if (total && cudaMemcpy(host.data(), dev_ptr(arr), sizeof(double)*total, cudaMemcpyDeviceToHost) != cudaSuccess) throw except::sp_exception("spade_b200: binary_write: device -> host copy failed");
        if (group.isroot())
        {
            std::FILE* f = std::fopen(filename.c_str(), "wb");
            if (!f) throw except::sp_exception("spade_b200: binary_write: cannot create " + filename);
            std::fclose(f);
        }
        group.sync();
        std::FILE* f = std::fopen(filename.c_str(), "r+b");
        if (!f) throw except::sp_exception("spade_b200: binary_write: cannot open " + filename);
        std::size_t nw = 0;
        for (std::size_t lb = 0; lb < nlb; ++lb)
        {
            const std::size_t lb_glob = grid.get_partition().to_global(utils::tag[partition::local](lb)).value;
            std::fseek(f, long(lb_glob*per_block*sizeof(double)), SEEK_SET);
            nw += std::fwrite(host.data() + lb*per_block, sizeof(double), per_block, f);
        }
        std::fclose(f);
        if (nw != total) throw except::sp_exception("spade_b200: binary_write: short write to " + filename);
        group.sync();
    }
    template <typename array_t> inline void binary_read(const std::string& filename, array_t& arr)
    {
        require_supported_array<array_t>();
        const auto& grid = arr.get_grid();
        const std::size_t nlb = grid.get_num_local_blocks(), total = arr.data.size();
        const std::size_t per_block = nlb ? total/nlb : 0;
        std::vector<double> host(total);
        std::FILE* f = std::fopen(filename.c_str(), "rb");
        if (!f) throw except::sp_exception("spade_b200: binary_read: cannot open " + filename);
        std::size_t nr = 0;
        for (std::size_t lb = 0; lb < nlb; ++lb)
        {
            const std::size_t lb_glob = grid.get_partition().to_global(utils::tag[partition::local](lb)).value;
            std::fseek(f, long(lb_glob*per_block*sizeof(double)), SEEK_SET);
            nr += std::fread(host.data() + lb*per_block, sizeof(double), per_block, f);
        }
        std::fclose(f);
        if (nr != total) throw except::sp_exception("spade_b200: binary_read: " + filename + " does not hold this rank's blocks");
        if (total && cudaMemcpy(dev_ptr(arr), host.data(), sizeof(double)*total, cudaMemcpyHostToDevice) != cudaSuccess) throw except::sp_exception("spade_b200: binary_read: host -> device copy failed");
    }
}

namespace spade::pde_algs
{
    // the new algorithm tag, next to tags.h:24-32
    static struct tb200_t : public fdiv_alg_base_t {} b200;

    namespace detail_b200
    {
        template <typename traits_t> struct tag_of
        {
            using raw = decltype(algs::get_trait(std::declval<const traits_t&>(), fdiv_alg_base_t::trait_label()));
            using type = typename utils::remove_all<raw>::type;
        };
        template <typename traits_t> concept has_b200_tag = std::same_as<typename tag_of<traits_t>::type, tb200_t>;
    }

    // More constrained than the reference's dispatcher (flux_div.h:23-41), so overload resolution picks it whenever the
    // trait list carries pde_algs::b200. The binding a maintainer would add upstream instead is one line in that
    // dispatcher: `if constexpr (std::same_as<fdiv_tag_t, tb200_t>) b200::flux_div(prims, rhs, flux_func, traits);`
    template <
        grid::multiblock_array sol_arr_t,
        grid::multiblock_array rhs_arr_t,
        typename flux_func_t,
        typename alg_traits_t>
    requires
        grid::has_centering_type<sol_arr_t, grid::cell_centered> && detail_b200::has_b200_tag<alg_traits_t>
    static void flux_div(
        const sol_arr_t& prims,
        rhs_arr_t& rhs,
        const flux_func_t& flux_func,
        const alg_traits_t& traits)
    {
        b200::require_supported_array<sol_arr_t>();
        b200::require_supported_array<rhs_arr_t>();
        using namespace sym::literals;
        const auto& incr = algs::get_trait(traits, "pde_increment"_sym, increment);        // default: increment (flux_div_basic.h:32-35)
        using incr_mode_t = typename utils::remove_all<decltype(incr)>::type;
        const spb_flux_desc d = b200::flux_desc(flux_func);
        b200::check(spb_flux_div(b200::grid_handle(prims), b200::dev_ptr(prims), b200::dev_ptr(rhs), &d,
                                 incr_mode_t::increment_mode ? 1 : 0, nullptr), "spb_flux_div");
        b200::check(spb_sync(nullptr), "spb_sync");
    }
}

namespace spade::time_integration
{
    // Same call as advance.h:236-280; more specialised in its `data` parameter (integrator_data_t<...> instead of a
    // bare template parameter), so partial ordering prefers it for GPU arrays; the stage update goes through
    // spb_rk_update (the fused prim <-> cons update of advance.h:57-102) instead of algs::transform_inplace.
    template <typename axis_t, typename var_state_t, typename rhs_state_t, typename scheme_t, typename rhs_t, typename boundary_t,
              typename state_t, typename gas_t>
    requires (scheme_t::is_rk_specialization && b200::on_gpu<var_state_t>)
    void integrate_advance(axis_t& axis, integrator_data_t<var_state_t, rhs_state_t, scheme_t>& data, const scheme_t& scheme,
                           const rhs_t& rhs, const boundary_t& boundary, const fluid_state::state_transform_t<gas_t, state_t>& trans)
    {
        b200::require_supported_array<var_state_t>();
        static_assert(std::same_as<state_t, fluid_state::cons_t<double>>, "spade_b200: the fused update integrates conserved variables");
        auto& q = data.solution(0);
        constexpr int num_stages = scheme_t::table_type::rows();
        using numeric_type = typename axis_t::value_type;
        const auto& dt = axis.timestep();
        spb_grid* gh = b200::grid_handle(q);
        const double gamma = trans.gas.get_gamma(), R = trans.gas.get_R();

        const auto update = [&](const auto& prev_row, const auto& curr_row)
        {
            using prev_row_t = typename utils::remove_all<decltype(prev_row)>::type;
            using curr_row_t = typename utils::remove_all<decltype(curr_row)>::type;
            constexpr int nupdate = curr_row_t::length();
            const double* ks[nupdate];
            double coeff[nupdate];
            algs::static_for<0, nupdate>([&](const auto& idx)
            {
                constexpr int i = idx.value;
                using diff_t = typename detail::ratio_diff_t<typename curr_row_t::elem_t<i>, typename prev_row_t::elem_t<i>>::type;
                ks[i] = b200::dev_ptr(data.residual(i));
                coeff[i] = detail::nonzero_t<diff_t>::value ? double(detail::coeff_value_t<numeric_type, diff_t>::value()*dt) : 0.0;
            });
            b200::check(spb_rk_update(gh, b200::dev_ptr(q), ks, nupdate, coeff, gamma, R, nullptr), "spb_rk_update");
            b200::check(spb_sync(nullptr), "spb_sync");
        };

        {
            constexpr numeric_type c0 = detail::coeff_value_t<numeric_type, typename scheme_t::dt_type::elem_t<0>>::value();
            axis.time() += c0*dt;
            rhs(data.residual(0), q, axis.time());
            axis.time() -= c0*dt;
        }
        algs::static_for<1, num_stages>([&](const auto& i_substep)
        {
            constexpr int i = i_substep.value;
            update(typename scheme_t::table_type::elem_t<i - 1>(), typename scheme_t::table_type::elem_t<i>());
            constexpr numeric_type ci = detail::coeff_value_t<numeric_type, typename scheme_t::dt_type::elem_t<i>>::value();
            axis.time() += ci*dt;
            boundary(q, axis.time());
            rhs(data.residual(i), q, axis.time());
            axis.time() -= ci*dt;
        });
        update(typename scheme_t::table_type::elem_t<num_stages - 1>(), typename scheme_t::accum_type());
        axis.time() += dt;
        boundary(q, axis.time());
    }

    // The same call again, selected when the rhs callback is b200::flux_div_rhs_t (and, for the ghost fusion and the overlapped
    // schedule, the boundary callback b200::exchange_bc_t): one kernel per stage and block range,
    // spb_flux_div_rk_stage[_exchange] (include/spade_b200.h). The stage plan comes from spb_rk_fused_plan and the supported
    // functor sets from spb_flux_div_rk_stage_supported — the same two calls the Python host makes, so the two host sides
    // cannot diverge. Falls back to the two-kernel path for schemes / functor sets outside that pattern.
    // With more than one rank (GPU) the schedule is the Python host's: the stage kernel runs on the rank-boundary blocks
    // (spb_exchange_boundary_blocks) on a high-priority side stream, their messages are packed straight into the
    // neighbours' buffers behind them, and the rank-interior blocks are advanced on the main stream AT THE SAME TIME; the
    // unpack waits on stream-ordered arrival flags. No host barrier and no stream synchronisation inside the step.
    template <typename axis_t, typename var_state_t, typename rhs_state_t, typename scheme_t, typename flux_func_t, typename boundary_t,
              typename state_t, typename gas_t>
    requires (scheme_t::is_rk_specialization && b200::on_gpu<var_state_t>)
    void integrate_advance(axis_t& axis, integrator_data_t<var_state_t, rhs_state_t, scheme_t>& data, const scheme_t& scheme,
                           const b200::flux_div_rhs_t<flux_func_t>& rhs, const boundary_t& boundary,
                           const fluid_state::state_transform_t<gas_t, state_t>& trans)
    {
        b200::require_supported_array<var_state_t>();
        static_assert(std::same_as<state_t, fluid_state::cons_t<double>>, "spade_b200: the fused update integrates conserved variables");
        constexpr int n = scheme_t::table_type::rows();
        using numeric_type = typename axis_t::value_type;
        const spb_flux_desc fd = b200::flux_desc(rhs.flux_func);

        // diffs[i][j]: coefficient of k_j in the update that follows stage i (last row: the accumulation row), advance.h:47-55,84-92
        double diffs[n*n];
        algs::static_for<0, n>([&](const auto& ii)
        {
            constexpr int i = ii.value;
            algs::static_for<0, n>([&](const auto& jj)
            {
                constexpr int j = jj.value;
                using prev_t = typename scheme_t::table_type::template elem_t<i>::template elem_t<j>;
                if constexpr (i + 1 < n)
                {
                    using curr_t = typename scheme_t::table_type::template elem_t<i + 1>::template elem_t<j>;
                    using diff_t = typename detail::ratio_diff_t<curr_t, prev_t>::type;
                    diffs[i*n + j] = detail::nonzero_t<diff_t>::value ? double(detail::coeff_value_t<numeric_type, diff_t>::value()) : 0.0;
                }
                else
                {
                    using curr_t = typename scheme_t::accum_type::template elem_t<j>;
                    using diff_t = typename detail::ratio_diff_t<curr_t, prev_t>::type;
                    diffs[i*n + j] = detail::nonzero_t<diff_t>::value ? double(detail::coeff_value_t<numeric_type, diff_t>::value()) : 0.0;
                }
            });
        });
        spb_stage_plan plan[n];
        const bool ok = spb_flux_div_rk_stage_supported(&fd) && spb_rk_fused_plan(n, diffs, plan) == 0;
        if (!ok)
        {
            // outside the fused pattern: same two-kernel path as for an opaque rhs callback
            const auto as_lambda = [&](auto& rr, const auto& qq, const auto& t) { rhs(rr, qq, t); };
            integrate_advance(axis, data, scheme, as_lambda, boundary, trans);
            return;
        }

        // time at which the boundary callback after stage i is evaluated: t + c_{i+1} dt, t + dt after the last one (advance.h:264-279)
        double tfrac[n];
        algs::static_for<0, n>([&](const auto& ii)
        {
            constexpr int i = ii.value;
            if constexpr (i + 1 < n) tfrac[i] = double(detail::coeff_value_t<numeric_type, typename scheme_t::dt_type::template elem_t<i + 1>>::value());
            else tfrac[i] = 1.0;
        });
        const auto t_start = axis.time();

        auto& q = data.solution(0);
        spb_grid* gh = b200::grid_handle(q);
        const std::size_t nd = q.data.size();
        double* bufs[2] = {b200::dev_ptr(q), b200::scratch_buffer(nd)};
        const double dt = double(axis.timestep());
        const int64_t nlb = (int64_t)q.get_grid().get_num_local_blocks();
        constexpr bool known_bc = b200::is_exchange_bc<boundary_t>::value;
        spb_exchange* fuse = nullptr;
        bool overlap = false;
        if constexpr (known_bc)
        {
            fuse = boundary.handle->plan;                                 // same-rank ghosts in the kernel; messages below
            overlap = boundary.group->size() > 1;
        }
        thread_local bool fuse_refused = false;
        // Deferred unpack (several ranks, no AMR interpolation, no wall fill, every peer writable directly): the off-rank ghost
        // cells of stage s are only read by the rank-boundary blocks of stage s+1, so their flag wait + unpack moves to the
        // side stream in front of that boundary kernel and the rank-interior kernel of stage s+1 starts without waiting for any
        // message — a neighbour GPU that runs late stalls the short chain unpack -> boundary kernel -> pack, which has the
        // whole interior kernel to catch up. Same schedule as api.py::integrator_t._advance_fused.
        bool defer = false;
        if constexpr (known_bc)
        {
            if (overlap)
            {
                auto& h = *boundary.handle;
                if (!h.wired) h.wire(*boundary.group);
                defer = h.all_direct && !boundary.has_fill && !fuse_refused
                        && spb_exchange_num_interp_send(h.plan) == 0 && spb_exchange_num_interp_recv(h.plan) == 0;
            }
        }
        double* pending = nullptr;                   // buffer whose off-rank ghost cells still wait for their messages
        int cur = 0;
        for (int i = 0; i < n; ++i)
        {
            const spb_stage_plan& st = plan[i];
            spb_stage_desc sd{};
            sd.nin = st.nin;
            for (int a = 0; a < st.nin; ++a) { sd.in[a] = b200::dev_ptr(data.residual(st.in[a])); sd.cq[a] = st.cq[a]*dt; sd.co[a] = st.co[a]; }
            sd.cq_self = st.cq_self*dt; sd.co_self = st.co_self;
            sd.out = st.out >= 0 ? b200::dev_ptr(data.residual(st.out)) : nullptr;
            bool ghosts_done = false;
            // one launch per part of the local blocks (boundary / interior / all), whatever their order in memory
            const auto launch = [&](const int part, cudaStream_t stream)
            {
                if (fuse && !fuse_refused)
                {
                    const int rc = spb_flux_div_rk_stage_part(gh, bufs[cur], bufs[1 - cur], &fd, &sd, fuse, 1, part, stream);
                    if (rc == SPB_ERR_UNSUPPORTED) fuse_refused = true;          // same-rank injections not canonical: separate exchange from now on
                    else { b200::check(rc, "spb_flux_div_rk_stage_part"); ghosts_done = true; return; }
                }
                b200::check(spb_flux_div_rk_stage_part(gh, bufs[cur], bufs[1 - cur], &fd, &sd, fuse, 0, part, stream), "spb_flux_div_rk_stage_part");
            };
            bool deferred_stage = false;
            if constexpr (known_bc)
            {
                if (overlap && defer)
                {
                    auto& sp = b200::streams();
                    auto& h = *boundary.handle;
                    if (i == 0) sp.fork();                                        // everything enqueued before this step
                    if (pending) { h.finish(pending, *boundary.group, sp.side); pending = nullptr; }
                    if (i > 0) cudaStreamWaitEvent(sp.side, sp.ev_i[(i - 1) & 1], 0);   // same-rank ghosts of the boundary blocks written by the previous interior kernel
                    launch(SPB_PART_BOUNDARY, sp.side);
                    if (ghosts_done)
                    {
                        cudaEventRecord(sp.ev_b[i & 1], sp.side);
                        h.begin(bufs[1 - cur], *boundary.group, sp.side);
                        if (i > 0) cudaStreamWaitEvent(nullptr, sp.ev_b[(i - 1) & 1], 0);   // ... and of the interior blocks by the previous boundary kernel
                        launch(SPB_PART_INTERIOR, nullptr);
                        cudaEventRecord(sp.ev_i[i & 1], nullptr);
                        pending = bufs[1 - cur];
                        deferred_stage = true;
                    }
                    else
                    {
                        // the plan refused the ghost fusion at the first launch: finish this stage the undeferred way
                        sp.join();
                        launch(SPB_PART_INTERIOR, nullptr);
                        h.begin(bufs[1 - cur], *boundary.group, nullptr);
                        defer = false;
                    }
                }
                else if (overlap)
                {
                    auto& sp = b200::streams();
                    auto& h = *boundary.handle;
                    sp.fork();
                    launch(SPB_PART_BOUNDARY, sp.side);
                    h.begin(bufs[1 - cur], *boundary.group, sp.side);             // the messages of the boundary blocks leave
                    launch(SPB_PART_INTERIOR, nullptr);
                    sp.join();
                }
                else launch(SPB_PART_ALL, nullptr);
            }
            else launch(SPB_PART_ALL, nullptr);
            cur = 1 - cur;
            axis.time() = t_start + tfrac[i]*dt;
            if constexpr (known_bc)
            {
                if (deferred_stage)
                {
                    if (i + 1 == n)
                    {
                        // end of the step: the last messages are unpacked and the main stream sees the whole state
                        auto& sp = b200::streams();
                        boundary.handle->finish(pending, *boundary.group, sp.side);
                        pending = nullptr;
                        sp.join();
                    }
                }
                else
                {
                    // the callback's work on the raw stage buffer (the result may sit in the scratch buffer): same-rank ghosts unless the
                    // kernel wrote them, the ghost cells fed by other ranks, then the wall fills of a channel solver
                    if (!ghosts_done) b200::check(spb_exchange_local(boundary.handle->plan, bufs[cur], nullptr), "spb_exchange_local");
                    else b200::check(spb_exchange_local_interp(boundary.handle->plan, bufs[cur], nullptr), "spb_exchange_local_interp");   // AMR: what the kernel left
                    if (overlap) boundary.handle->finish(bufs[cur], *boundary.group, nullptr);
                    boundary.fill(q, bufs[cur]);
                }
            }
            else
            {
                // opaque boundary callback: it needs a SPADE array, so the state is brought back into q first
                if (cur == 1)
                {
                    cudaMemcpyAsync(bufs[0], bufs[1], sizeof(double)*nd, cudaMemcpyDeviceToDevice, nullptr);
                    cur = 0;
                }
                boundary(q, axis.time());
            }
        }
        if (cur == 1) cudaMemcpyAsync(bufs[0], bufs[1], sizeof(double)*nd, cudaMemcpyDeviceToDevice, nullptr);    // odd number of stages
        axis.time() = t_start + dt;
        b200::check(spb_sync(nullptr), "spb_sync");
    }

    // ssprk3_opt: the 2-register SSPRK3 of advance.h:359-402 (detail::opt_rk3_s0/s1/s2, advance.h:286-354) for device::gpu arrays;
    // more specialised in `data` than the reference's overload, so partial ordering prefers it. Same operation order as the
    // reference: stage 1 overwrites r0 with (dt/6)(r0 + r1).
    template <typename axis_t, typename var_state_t, typename rhs_state_t, typename rhs_t, typename boundary_t, typename state_t, typename gas_t>
    requires (b200::on_gpu<var_state_t>)
    void integrate_advance(axis_t& axis, integrator_data_t<var_state_t, rhs_state_t, tspecial_rk3_t>& data, const tspecial_rk3_t&, const rhs_t& rhs,
                           const boundary_t& boundary, const fluid_state::state_transform_t<gas_t, state_t>& trans)
    {
        b200::require_supported_array<var_state_t>();
        static_assert(std::same_as<state_t, fluid_state::cons_t<double>>, "spade_b200: ssprk3_opt integrates conserved variables");
        auto& q  = data.solution(0);
        auto& r0 = data.residual(0);
        auto& r1 = data.residual(1);
        using numeric_type = typename axis_t::value_type;
        const auto& dt = axis.timestep();
        const numeric_type tc[3] = {numeric_type(0.0), numeric_type(1.0), numeric_type(0.5)};
        spb_grid* gh = b200::grid_handle(q);
        const double gamma = trans.gas.get_gamma(), R = trans.gas.get_R();
        const auto stage = [&](const int s)
        {
            b200::check(spb_ssprk3_stage(gh, s, b200::dev_ptr(q), b200::dev_ptr(r0), b200::dev_ptr(r1), double(dt), gamma, R, nullptr), "spb_ssprk3_stage");
            b200::check(spb_sync(nullptr), "spb_sync");
        };
        axis.time() += tc[0]*dt;
        rhs(r0, q, axis.time());
        axis.time() -= tc[0]*dt;
        stage(0);
        axis.time() += tc[1]*dt;
        boundary(q, axis.time());
        rhs(r1, q, axis.time());
        axis.time() -= tc[1]*dt;
        stage(1);
        axis.time() += tc[2]*dt;
        boundary(q, axis.time());
        rhs(r1, q, axis.time());
        axis.time() -= tc[2]*dt;
        stage(2);
        axis.time() += dt;
        boundary(q, axis.time());
    }

    // Generic integrate_advance (advance.h:109-230) with time_integration::identity_transform on device::gpu arrays: the array
    // itself is integrated. Every `resid *= c; sol (+|-)= resid; resid *= 1/c` triple of the reference (three passes, each
    // followed by a device synchronisation) is ONE kernel (spb_axpy_roundtrip), bit-identical including the round trip of the
    // residual. Tables with a second copy of the solution (rk2hs_t, ssprk3hs_t; explicit.h:33) skip the subtraction pass.
    template <typename axis_t, typename var_state_t, typename rhs_state_t, typename scheme_t, typename rhs_t, typename boundary_t>
    requires (scheme_t::is_rk_specialization && b200::on_gpu<var_state_t>)
    void integrate_advance(axis_t& axis, integrator_data_t<var_state_t, rhs_state_t, scheme_t>& data, const scheme_t& scheme,
                           const rhs_t& rhs, const boundary_t& boundary, const identity_transform_t&)
    {
        b200::require_supported_array<var_state_t>();
        constexpr bool second = scheme_t::var_size() > 1;
        constexpr int n = scheme_t::table_type::rows();
        using numeric_type = typename axis_t::value_type;
        const auto& dt = axis.timestep();
        const auto axpy = [&](auto& sol, auto& resid, const numeric_type c, const int subtract)
        {
            b200::check(spb_axpy_roundtrip(b200::grid_handle(sol), b200::dev_ptr(sol), b200::dev_ptr(resid), double(c), subtract, nullptr), "spb_axpy_roundtrip");
        };
        algs::static_for<0, n>([&](const auto& ii)
        {
            constexpr int i = ii.value;
            auto& sol = data.solution(second ? 1 : 0);
            if constexpr (second)
                cudaMemcpyAsync(b200::dev_ptr(sol), b200::dev_ptr(data.solution(0)), sizeof(double)*sol.data.size(), cudaMemcpyDeviceToDevice, nullptr);
            const auto sweep = [&](const int subtract)
            {
                algs::static_for<0, i>([&](const auto& jj)
                {
                    constexpr int j = jj.value;
                    using coeff_t = typename scheme_t::table_type::template elem_t<i>::template elem_t<j>;
                    if constexpr (detail::nonzero_t<coeff_t>::value)
                    {
                        constexpr numeric_type coeff = detail::coeff_value_t<numeric_type, coeff_t>::value();
                        axpy(sol, data.residual(j), dt*coeff, subtract);
                    }
                });
            };
            sweep(0);
            b200::check(spb_sync(nullptr), "spb_sync");
            constexpr numeric_type tcoef = detail::coeff_value_t<numeric_type, typename scheme_t::dt_type::template elem_t<i>>::value();
            axis.time() += tcoef*dt;
            if constexpr (i > 0) boundary(sol, axis.time());
            rhs(data.residual(i), sol, axis.time());
            axis.time() -= tcoef*dt;
            if constexpr (!second) sweep(1);
        });
        auto& new_solution = data.solution(0);
        algs::static_for<0, n>([&](const auto& ii)
        {
            constexpr int i = ii.value;
            using coeff_t = typename scheme_t::accum_type::template elem_t<i>;
            if constexpr (detail::nonzero_t<coeff_t>::value)
            {
                constexpr numeric_type coeff = detail::coeff_value_t<numeric_type, coeff_t>::value();
                axpy(new_solution, data.residual(i), coeff*dt, 0);
            }
        });
        b200::check(spb_sync(nullptr), "spb_sync");
        axis.time() += dt;
        boundary(new_solution, axis.time());
    }
}
