/* spade_b200.h — C ABI of libspade_b200.so: the B200 (sm_100a) implementation of SPADE's RHS hot
 * path (flux_div + ghost exchange + RK stage update + reductions).
 *
 * Every entry point takes plain pointers and sizes. Device memory is owned by the caller (SPADE's
 * grid_array / device_vector, or a torch tensor): the library never allocates persistent state
 * except inside the opaque handles created and destroyed here. There is NO CPU fallback: every
 * compute entry point launches hand-written CUDA kernels and returns an error if no device is
 * present.
 *
 * Array layout (bit-exact with the reference's default mem_map::linear_t, reference
 * src/core/mem_map.h:484-496, src/grid/grid_array.h:252-257):
 *     off(v,i,j,k,lb) = v + 5*((i+g0) + (n0+2*g0)*((j+g1) + (n1+2*g1)*((k+g2) + (n2+2*g2)*lb)))
 * v: 0..4 = (p,T,u,v,w) for primitive arrays, (continuity, energy, x/y/z momentum) for residuals
 * (reference src/navier-stokes/fluid_state.h:11-77); i,j,k may run over [-g, n+g).
 *
 * Error convention: 0 = success; nonzero = failure (cudaError_t value or SPB_ERR_*), message from
 * spb_last_error(). The reference throws except::sp_exception after a failed CUDA call
 * (reference src/dispatch/execute.h:86-96); the C++ shim (include/spade_b200_shim.hpp) turns a
 * nonzero return into that exception.
 *
 * Threading: one host thread (or process) per GPU, each having selected its device; calls on
 * different devices are independent (reference src/parallel/compute_pool.h:497-514).
 * Streams: `stream` is a cudaStream_t passed as void* (NULL = default stream). All calls are
 * asynchronous with respect to the host unless stated otherwise; the reference's synchronous
 * semantics (src/dispatch/execute.h:85) are recovered with spb_sync().
 */
#ifndef SPADE_B200_H
#define SPADE_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SPB_NVAR 5

enum {
    SPB_ERR_BAD_ARG      = 10001,
    SPB_ERR_UNSUPPORTED  = 10002,
    SPB_ERR_NO_DEVICE    = 10003,
    SPB_ERR_DRIVER       = 10004
};

/* ---- flux functor description -------------------------------------------------------------
 * Closed set of the reference functor types the kernels implement; the C++ shim recognises the
 * composed functor type at compile time and fills this POD from its by-value members.
 *   conv: convective scheme          reference src/navier-stokes/convective.h
 *   diss: dissipative scheme blended by the sensor (hybrid_scheme_t), reference hybrid_scheme.h:15-47
 *   visc: viscous::visc_lr with viscous_laws::constant_viscosity_t, reference viscous.h:14-113 */
enum { SPB_CONV_NONE = 0, SPB_CONV_TOTANI = 1 /* totani_lr, convective.h:54-94 */,
       SPB_CONV_CENT_KEEP4 = 2 /* cent_keep<4>, convective.h:97-192 */,
       SPB_CONV_FWENO = 3 /* fweno_t alone, convective.h:336-497 */,
       SPB_CONV_CENT_KEEP6 = 4, SPB_CONV_CENT_KEEP8 = 5 /* cent_keep<6>, cent_keep<8>: 3 / 4 exchange cells, no hybrid */ };
enum { SPB_DISS_NONE = 0, SPB_DISS_FWENO = 1 /* hybrid_scheme_t(conv, fweno_t, ducros_t, tag) */ };
enum { SPB_BLEND_FULL_FLUX = 0 /* (1-a)F0 + a F1 */, SPB_BLEND_DISS_FLUX = 1 /* F0 + a F1 */ };
/* LES closure of visc_lr: viscous_laws::sgs_visc_t(constant_viscosity_t, subgrid_scale::wale_t(gas, cw, delta, prt)),
 * reference src/navier-stokes/viscous_laws.h:175-216, subgrid_scale.h:25-91: mu += mu_t, beta -= 0.66666666667 mu_t,
 * alpha += mu_t/prt with mu_t from the face gradient and the face density. */
enum { SPB_SGS_NONE = 0, SPB_SGS_WALE = 1 };

typedef struct spb_flux_desc
{
    int    conv;        /* SPB_CONV_*  */
    int    diss;        /* SPB_DISS_*  */
    int    blend;       /* SPB_BLEND_* (used when diss != NONE) */
    int    visc;        /* 0 / 1 : add viscous::visc_lr */
    double gamma, R;    /* fluid_state::ideal_gas_t, reference src/navier-stokes/gas.h:35-49 */
    double mu;          /* constant_viscosity_t::visc */
    double beta;        /* constant_viscosity_t::beta   (= -2 mu/3 as stored by the reference) */
    double prandtl_inv; /* constant_viscosity_t::prandtl_inv */
    double sensor_eps;  /* state_sensor::ducros_t::epsilon, reference state_sensor.h:21-43 */
    int    sgs;         /* SPB_SGS_* (needs visc) */
    double sgs_cw, sgs_delta, sgs_prt;   /* wale_t::cw, delta, prt */
    int    weno_linear; /* 1: fweno_t / weno_t<.., disable_smooth> (convective.h:248-252): the linear weights 1/3, 2/3, 2/3, 1/3 instead of the
                           nonlinear ones; 0 (enable_smooth) is the default of the reference */
} spb_flux_desc;

/* ---- grid ------------------------------------------------------------------------------------
 * Local (per-rank) block geometry, replacing the device image of grid_geometry_t
 * (reference src/grid/grid_geometry.h:16-100, src/grid/cartesian_grid.h:114-165).
 * bbox: host array [nlb][6] = xmin,xmax,ymin,ymax,zmin,zmax of each local block in computational
 * coordinates; dx = (max-min)/nx and inv_dx = 1.0/dx are formed exactly as the reference does.
 * The RHS kernels read the spacings of a block as refinement LEVELS: per direction the distinct values of inv_dx (values that
 * differ by no more than the rounding of the block bounds explains, 4 eps |x|max / block size, are one level; at most 16 per
 * direction, else the RHS calls return SPB_ERR_UNSUPPORTED). A lattice with one level per direction is "uniform". */
typedef struct spb_grid spb_grid;
int  spb_grid_create(spb_grid** out, const int nx[3], const int ng[3], int64_t nlb, const double* bbox_host);
void spb_grid_destroy(spb_grid* g);
int64_t spb_grid_array_size(const spb_grid* g);                    /* doubles in one 5-variable array */
/* Host-only (no device needed): the refinement levels spb_grid_create derives from the block boxes. lev_n[d] = number of
 * distinct inv_dx along d (-1: more than 16), lev_inv = [3][16] doubles, lev_of_block (may be null) = [nlb] packed indices
 * l0 | l1 << 8 | l2 << 16, round_tol (may be null) = [3] relative spread that the rounding of the block bounds explains. */
int  spb_grid_spacing_levels(const int nx[3], int64_t nlb, const double* bbox_host, int lev_n[3], double* lev_inv,
                             int* lev_of_block, double* round_tol);
int64_t spb_grid_offset(const spb_grid* g, int v, int i, int j, int k, int64_t lb);

/* ---- general coordinates: replaces the geometry of coords::diagonal_coords ---------------------
 * reference src/core/coord_system.h:65-90 (one 1-D mapping x_d(xi_d) per direction), calc_normal_vector :250-267,
 * calc_jacobian :295-302, used by flux_div_basic.h:49-71 and info::metric (src/omni/infos/info_metric.h:24-32).
 * A diagonal system is separable, so the whole geometry is three 1-D tables per block and direction of the
 * coordinate derivative m_d = dx_d/dxi_d (host arrays, copied by the call):
 *   area[d]  [nlb][n_d + 2 g_d]      m_d at the cell centres AS info::metric EVALUATES IT. The reference passes the
 *                                    mapped position into coord_deriv (info_metric.h:31: grid.get_coords(idx)); a
 *                                    caller that wants the reference's numbers fills this table the same way, one
 *                                    that wants the consistent metric passes the same values as jac[d].
 *   jac[d]   [nlb][n_d + 2 g_d]      m_d at the computational cell centres (calc_jacobian, flux_div_basic.h:49-50)
 *   face[d]  [nlb][n_d + 2 g_d + 1]  m_d at the computational face positions (entry i = lower face of padded cell i)
 * With a metric set, spb_flux_div / spb_flux_div_rk_stage compute
 *   rhs(c) = J(c) sum_d (F_lower - F_upper)/dxi_d,   J = 1/(m_0 m_1 m_2),   F = flux with metric vector (m_t1 m_t2) e_d,
 * and the face gradient of visc_lr / ducros_t is transformed as d/dx_d = (1/m_d) d/dxi_d (face[d] for the normal
 * direction, jac[t] for the tangential ones). That transform is NOT in the reference (info_gradient.h:83
 * static_asserts coords::identity): the convective functors match the reference on stretched grids, the viscous and
 * sensor terms are this library's completion. spb_source_term divides by J like source_term.h:38-46.
 * The one-kernel ghost fusion (spb_flux_div_rk_stage_exchange with a plan) is for identity coordinates; with a metric
 * it returns SPB_ERR_UNSUPPORTED and the caller runs spb_flux_div_rk_stage + spb_exchange_local.
 * m == NULL returns the grid to coords::identity. */
typedef struct spb_metric_desc
{
    const double* area[3];
    const double* jac[3];
    const double* face[3];
} spb_metric_desc;
int spb_grid_set_metric(spb_grid* g, const spb_metric_desc* m);
int spb_grid_has_metric(const spb_grid* g);

/* ---- RHS: replaces pde_algs::flux_div(q, rhs, flux_func, traits) ------------------------------
 * reference src/pde-algs/flux-div/flux_div.h:23-41, flux_div_basic.h:17-77.
 * rhs(cell) (+)= sum_dir (F_lowerface - F_upperface) / dx_dir on interior cells (identity coords,
 * Jacobian 1). increment = 0 is the `overwrite` trait: interior cells are overwritten (the ghost
 * cells of rhs, which the reference zero-fills, are left untouched — they are never read).
 * q must have its ghost cells filled (exchange) to the stencil depth of the scheme. */
int spb_flux_div(const spb_grid* g, const double* q_dev, double* rhs_dev, const spb_flux_desc* flux,
                 int increment, void* stream);
/* Same, restricted to local blocks [lb_begin, lb_end): used to overlap the exchange of rank-boundary
 * blocks with the RHS of interior blocks. */
int spb_flux_div_blocks(const spb_grid* g, const double* q_dev, double* rhs_dev, const spb_flux_desc* flux,
                        int increment, int64_t lb_begin, int64_t lb_end, void* stream);

/* ---- fused RK stage: flux_div + stage update in one pass over q ------------------------------
 * Replaces the pair  rhs(k_i, q, t); detail::transform_advance_to(...)  of integrate_advance
 * (reference src/time-integration/advance.h:254-275, 57-102) when the rhs callback is flux_div itself:
 * with r = flux_div(q_in) on a cell,
 *     q_out = prim( cons(q_in) + cq_self*r + cq[0]*in[0] + cq[1]*in[1] )      (interior cells; ghosts of q_out untouched)
 *     out   = co_self*r + co[0]*in[0] + co[1]*in[1]                           (if out != NULL; may alias in[a])
 * q is read once and the residual never makes a round trip through HBM before the update. q_out must be a
 * different array from q_in (neighbouring cells still read q_in). Implemented for the one-ghost-cell functor
 * set (totani_lr and/or visc_lr); other descriptors return SPB_ERR_UNSUPPORTED and the caller uses
 * spb_flux_div + spb_rk_update. The caller folds dt and the Butcher coefficients into cq (see
 * spade_b200/api.py::integrator_t for rk4_t). */
/* 1 if the fused stage exists for this functor set (totani_lr and/or visc_lr; visc_lr with hybrid(totani_lr | cent_keep<4>,
 * fweno_t) or cent_keep<4|6|8>; the WALE closure on those up to cent_keep<4>), 0 otherwise. Host-only. */
int spb_flux_div_rk_stage_supported(const spb_flux_desc* flux);
typedef struct spb_stage_desc
{
    int           nin;          /* number of residual registers read (0..2) */
    const double* in[2];
    double        cq_self, cq[2];
    double*       out;          /* residual register written, or NULL */
    double        co_self, co[2];
} spb_stage_desc;
int spb_flux_div_rk_stage(const spb_grid* g, const double* q_in_dev, double* q_out_dev, const spb_flux_desc* flux,
                          const spb_stage_desc* stage, int64_t lb_begin, int64_t lb_end, void* stream);

/* Same, and the finished q_out planes are also stored into the ghost cells of the same-rank neighbour blocks
 * (the same-rank transactions of `exch`, reference src/grid/make_exchange.h:166-203): after the call q_out needs no
 * spb_exchange_local, only the messages of other ranks (spb_exchange_pack / unpack). The ghost values are bit-identical
 * to a separate exchange of q_out. With lb_begin/lb_end the ghost cells fed by blocks outside the range are not
 * written. The one-ghost-cell functor set sends the ghosts from a dedicated warp (TMA stores of the staged plane; 2 exchange
 * cells along i), every other functor set stores them from the threads that own the cells (any number of exchange cells).
 * SPB_ERR_UNSUPPORTED if the plan holds same-rank INJECTION transactions other than the 26 canonical boxes; the caller then
 * uses spb_flux_div_rk_stage + spb_exchange_local. Interpolation transactions (AMR) are never done by the kernel: after
 * a fused call the caller runs spb_exchange_local_interp. exch == NULL is spb_flux_div_rk_stage. */
typedef struct spb_exchange spb_exchange;
int spb_flux_div_rk_stage_exchange(const spb_grid* g, const double* q_in_dev, double* q_out_dev, const spb_flux_desc* flux,
                                   const spb_stage_desc* stage, spb_exchange* exch, int64_t lb_begin, int64_t lb_end,
                                   void* stream);

/* Stage plan of the fused path for a Butcher table (the ONE planner both host sides use: include/spade_b200_shim.hpp and
 * spade_b200/api.py). `diffs` is row-major [n][n]: diffs[i][j] = a_{i+1,j} - a_{i,j} (last row: b_j - a_{n-1,j}) evaluated
 * in double exactly as the reference forms its update coefficients (ratio_diff_t + coeff_value_t,
 * src/time-integration/advance.h:47-55, 84-92), WITHOUT dt. plan[i] describes stage i: which residual registers the stage
 * kernel reads (`in`, at most two), with which coefficients they enter q_out (cq) and the written register (co), and which
 * register it writes (`out`, -1 = none). A final update that needs more than two earlier residuals gets their combination
 * C prepared by the stage before it in register n-2 (rk4: C = k0/6 + k1/3 - 2 k2/3, so an RK4 step reads 5 and writes 3
 * residual planes instead of 10 and 4). Returns SPB_ERR_UNSUPPORTED when a stage would need more than two inputs (the
 * caller then runs flux_div + spb_rk_update), SPB_ERR_BAD_ARG for n outside 1..8. Host-only, no GPU needed.
 * After a fused step the registers differ from the reference's: k_{n-1} is never written and register n-2 may hold C. */
typedef struct spb_stage_plan
{
    int    nin;            /* residual registers read (0..2) */
    int    in[2];          /* their indices */
    double cq[2], co[2];   /* coefficients into q_out (multiply by dt) and into the written register */
    double cq_self;        /* coefficient of this stage's own rhs in q_out (multiply by dt) */
    double co_self;        /* coefficient of this stage's own rhs in the written register */
    int    out;            /* register written, -1 = none */
} spb_stage_plan;
int spb_rk_fused_plan(int n, const double* diffs, spb_stage_plan* plan);

/* The fused stage on ONE PART of the local blocks in ONE launch, whatever their order in memory: part SPB_PART_BOUNDARY = the
 * blocks spb_exchange_boundary_blocks marks (sources of off-rank sends), SPB_PART_INTERIOR = the rest, SPB_PART_ALL = every
 * block. The overlapped schedule runs BOUNDARY on a side stream, sends the messages behind it and runs INTERIOR at the
 * same time on the main stream. On AMR grids the rank-boundary blocks are scattered in local order, so contiguous ranges
 * (lb_begin, lb_end) would mean one launch per run. fuse_ghosts != 0 also stores the same-rank injection ghosts from the
 * kernel (SPB_ERR_UNSUPPORTED if the plan cannot be fused: retry with fuse_ghosts = 0 and spb_exchange_local). */
enum { SPB_PART_ALL = 0, SPB_PART_BOUNDARY = 1, SPB_PART_INTERIOR = 2 };
int spb_flux_div_rk_stage_part(const spb_grid* g, const double* q_in_dev, double* q_out_dev, const spb_flux_desc* flux,
                               const spb_stage_desc* stage, spb_exchange* exch, int fuse_ghosts, int part, void* stream);

/* ---- RK stage update: replaces detail::transform_advance_to -----------------------------------
 * reference src/time-integration/advance.h:57-102: per interior cell
 *   w = cons(q); w += sum_j coeff[j]*k_j  (only j with coeff[j] != 0, in order); q = prim(w)
 * coeff[j] = (a_ij - a_{i-1,j})*dt as formed by the caller (the shim forms it like advance.h:84-92). */
int spb_rk_update(const spb_grid* g, double* q_dev, const double* const* k_dev, int nk, const double* coeff,
                  double gamma, double R, void* stream);
/* 2-register SSPRK3 stages opt_rk3_s0/s1/s2, reference src/time-integration/advance.h:286-354.
 * stage 1 overwrites r0 with (dt/6)(r0+r1) exactly like the reference. */
int spb_ssprk3_stage(const spb_grid* g, int stage, double* q_dev, double* r0_dev, const double* r1_dev,
                     double dt, double gamma, double R, void* stream);

/* Generic integrate_advance (reference src/time-integration/advance.h:109-230, the path taken when `trans` is not a
 * state_transform_t, e.g. time_integration::identity_transform): its building block
 *     resid *= c;  sol += resid  (sol -= resid if subtract);  resid *= 1.0/c
 * (advance.h:149-153, 190-194, 216-221; the scalar passes touch every element incl. exchange cells, the array pass the
 * interior, grid_array.h:289-321,359-369) as ONE pass instead of three, bit-identical including the round trip of resid. */
int spb_axpy_roundtrip(const spb_grid* g, double* sol_dev, double* resid_dev, double c, int subtract, void* stream);

/* ---- reduction: replaces algs::transform_reduce(array, make_reduction(array, f, op)) ----------
 * reference src/algs/transform_reduce.h:43-191. Closed set of element kernels `f`. */
enum { SPB_RED_MAX = 0, SPB_RED_SUM = 1 };
enum { SPB_FN_WAVESPEED = 0 /* sqrt(gamma R T) + |u|  (CFL, development/cuda-tgv/main.cc:152-161) */,
       SPB_FN_VAR = 1 /* q[ivar] */, SPB_FN_ABSVAR = 2 /* |q[ivar]| */, SPB_FN_KINETIC = 3 /* 0.5 rho |u|^2 */ };
/* Synchronous: returns the reduced value over the interior cells of all local blocks in *out_host. */
int spb_reduce(const spb_grid* g, const double* q_dev, int op, int fn, int ivar, double gamma, double R,
               double* out_host, void* stream);

/* ---- ghost exchange: replaces make_exchange / arr_exchange_t::exchange -------------------------
 * reference src/grid/make_exchange.h:111-421, exchange_config.h:286-419, get_transaction.h:11-97.
 * A plan holds this rank's sorted send and receive transaction lists; lists are bit-identical to
 * exchange_config_t::send_data[0] / recv_data[0] of the reference (same order, boxes and tags). */
typedef struct spb_exchange spb_exchange;
/* Uniform cartesian block lattice partitioned like partition::block_partition_t
 * (reference src/grid/partition.h:27-84, src/grid/cartesian_blocks.h:40-105). Host-only, no GPU needed. */
int  spb_exchange_create(spb_exchange** out, const int nblocks[3], const int nx[3], const int ng[3],
                         const int periodic[3], int rank, int nranks);
/* From transaction tables produced elsewhere (e.g. marshalled from SPADE's exchange_config_t by the shim);
 * each transaction is 16 int64: tag, rank_send, rank_recv, glob_src_blk, glob_dst_blk,
 * src.min[0..3], src.size[0..2], dst.min[0..3]   (index 3 = local block id, -1 if not owned). */
int  spb_exchange_create_from_tables(spb_exchange** out, const int nx[3], const int ng[3], int rank, int nranks,
                                     const int64_t* send, int64_t nsend, const int64_t* recv, int64_t nrecv);
/* AMR (reference src/grid/get_transaction.h:100-263, transactions.h:136-234): adds the interpolation transactions
 * (patch_fill_t lists exchange_config_t::send_data[1] / recv_data[1]) to a plan made from tables. 26 int64 each: the 16
 * fields above with source = the donor region, then dest.size[0..2], i_coeff[0..2], i_incr[0..2], 0. A destination cell
 * (ix,iy,iz) of the box is 1/8 of the sum of the 2^3 donors at source.min + ((ix << (i_coeff+1)) >> 1) + d*i_incr
 * (fine -> coarse: mean of the covering fine cells; coarse -> fine: injection), make_exchange.h:208-314. In a peer message
 * the interpolation section follows the injection section. Call before the first device call of the plan. */
int  spb_exchange_add_interp(spb_exchange* e, const int64_t* send, int64_t nsend, const int64_t* recv, int64_t nrecv);
int64_t spb_exchange_num_interp_send(const spb_exchange* e);
int64_t spb_exchange_num_interp_recv(const spb_exchange* e);
void spb_exchange_destroy(spb_exchange* e);
int64_t spb_exchange_num_send(const spb_exchange* e);
int64_t spb_exchange_num_recv(const spb_exchange* e);
/* copies the tables out (host); offs gets 6 int64 per peer: send_message_size, recv_message_size,
 * send_rank_offsets, send_rank_sizes, recv_rank_offsets, recv_rank_sizes (cells / transactions). */
int  spb_exchange_tables(const spb_exchange* e, int64_t* send, int64_t* recv, int64_t* offs);
int64_t spb_exchange_local_blocks(const spb_exchange* e);          /* blocks owned by this rank */
int64_t spb_exchange_first_block(const spb_exchange* e);           /* global id of local block 0 */
/* cells this rank sends to / receives from `peer` (message sizes in cells; x5 doubles). Element
 * order inside the message is the reference's: transaction order, ix + nx*(iy + ny*iz), v fastest. */
int64_t spb_exchange_send_cells(const spb_exchange* e, int peer);
int64_t spb_exchange_recv_cells(const spb_exchange* e, int peer);
/* mask[lb] = 1 for every local block that is the source of an off-rank send transaction — injection AND interpolation
 * (patch_fill_t donors) — 0 otherwise; nlb = number of local blocks. The overlapped schedule advances these blocks first,
 * sends their messages and advances the rest while the messages fly. Host-only. */
int spb_exchange_boundary_blocks(const spb_exchange* e, int64_t nlb, unsigned char* mask);
/* same-rank transactions: q(dst) = q(src) on the device (make_exchange.h:166-203). */
int spb_exchange_local(spb_exchange* e, double* q_dev, void* stream);
/* only the same-rank INTERPOLATION transactions (AMR): what remains after spb_flux_div_rk_stage_exchange has stored the
 * same-level (injection) ghosts from inside the stage kernel. Every source cell of a transaction is an interior cell, so
 * the split gives the same ghost values as spb_exchange_local. No-op on plans without interpolation lists. */
int spb_exchange_local_interp(spb_exchange* e, double* q_dev, void* stream);
/* pack q into the contiguous message for `peer` (make_exchange.h:136-164) / unpack (340-369). */
int spb_exchange_pack(spb_exchange* e, const double* q_dev, int peer, double* sendbuf_dev, void* stream);
int spb_exchange_unpack(spb_exchange* e, double* q_dev, int peer, const double* recvbuf_dev, void* stream);
/* One-sided variant over NVLink peer memory: pack straight into the peer's receive buffer
 * (peer_recvbuf_dev is a pointer into the peer GPU's memory, mapped with cudaIpcOpenMemHandle or
 * cudaDeviceEnablePeerAccess). Same element order as spb_exchange_pack. */
int spb_exchange_pack_peer(spb_exchange* e, const double* q_dev, int peer, double* peer_recvbuf_dev, void* stream);

/* One process per GPU: the receive buffers are exported through CUDA IPC (spb_ipc_export / spb_ipc_import), the pack kernel of
 * a rank stores its message straight into the neighbour's buffer over NVLink (spb_exchange_pack_peer) and raises the
 * neighbour's flag (spb_flag_signal, stream-ordered after the pack); the neighbour's stream waits on its own flag
 * (spb_flag_wait) before spb_exchange_unpack. No SM of either GPU is held by a communication kernel while the RHS kernel
 * runs, unlike NCCL send/recv (reference counterpart: exchange_message_t::send_all with cudaMemcpyPeer,
 * exchange_message.h:14-56, compute_pool.h:93-99). spb_dev_alloc returns zero-filled cudaMalloc memory (IPC-shareable). */
int spb_dev_alloc(void** out, size_t bytes);
int spb_dev_free(void* p);
int spb_ipc_export(const void* dev_ptr, unsigned char handle[64]);
int spb_ipc_import(const unsigned char handle[64], void** out);
int spb_ipc_close(void* p);
int spb_flag_signal(unsigned long long* flag_dev, unsigned long long value, void* stream);
int spb_flag_wait(const unsigned long long* flag_dev, unsigned long long value, void* stream);

/* ---- domain-boundary ghost fill: replaces algs::boundary_fill(arr, boundaries, kern) ------------
 * reference src/grid/boundary_fill.h:32-133. One call fills ONE boundary (idir, pm): pm = 0 the lower, 1 the upper
 * face of the block lattice along idir. `blocks_host` lists the local blocks on that face
 * (grid_geometry_t::boundary_blocks[2*idir+pm], reference src/grid/cartesian_grid.h:139-146). Every cell beyond the
 * face is written, including the exchange cells of the two tangential directions, exactly like the reference; the
 * caller fills several boundaries in the reference's order xmin, xmax, ymin, ymax, zmin, zmax (a later boundary reads
 * ghosts an earlier one wrote). The user kernel is one of a closed set:
 *   SPB_BC_MIRROR  kern(image, idir): ghost[v] = a[v]*image[v] + b[v] with image the mirror cell through the face
 *                  (-1 - i below, 2 n - (i + 1) above); if use_normal the velocity component along idir uses a_normal.
 *                  no-slip isothermal wall: a = (1,-1,-1,-1,-1), b = (0, 2 T_wall, 0, 0, 0); no-slip adiabatic:
 *                  a = (1,1,-1,-1,-1); symmetry / slip: a = 1, a_normal = -1; Dirichlet: a = 0, b = value.
 *   SPB_BC_EXTRAP  boundary::extrapolate<order>: Lagrange extrapolation through the order+1 cells next to the face.
 * Ghost values are bit-identical to the reference's CPU result (no FMA contraction in the kernel). */
enum { SPB_BC_MIRROR = 0, SPB_BC_EXTRAP = 1 };
typedef struct spb_bc_desc
{
    int    kind;          /* SPB_BC_* */
    int    order;         /* SPB_BC_EXTRAP */
    double a[5], b[5];    /* SPB_BC_MIRROR */
    int    use_normal;
    double a_normal;
} spb_bc_desc;
int spb_boundary_fill(const spb_grid* g, double* q_dev, int idir, int pm, const int64_t* blocks_host, int64_t nblocks,
                      const spb_bc_desc* bc, void* stream);

/* ---- source term: replaces pde_algs::source_term(q, rhs, source_term_func) ----------------------
 * reference src/pde-algs/source_term.h:25-51: rhs(cell) += S(q(cell))/jac on interior cells (jac = 1). Closed set:
 *   SPB_SRC_BODY_FORCE  S = (0, f.u, f[0], f[1], f[2])   (the forcing of a channel run)
 *   SPB_SRC_CONSTANT    S = f[0..4] */
enum { SPB_SRC_BODY_FORCE = 0, SPB_SRC_CONSTANT = 1 };
typedef struct spb_source_desc { int kind; double f[5]; } spb_source_desc;
int spb_source_term(const spb_grid* g, const double* q_dev, double* rhs_dev, const spb_source_desc* src, void* stream);

/* ---- utilities ---------------------------------------------------------------------------------- */
const char* spb_last_error(void);
int  spb_device_count(void);
int  spb_sync(void* stream);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches) */
int64_t spb_launch_count(void);
const char* spb_version(void);

#ifdef __cplusplus
}
#endif
#endif
