// Internal declarations shared by the translation units of libspade_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <atomic>
#include "../../include/spade_b200.h"

namespace spb
{
    void set_error(const std::string& msg);
    extern std::atomic<int64_t> g_launches;

    inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
    {
        set_error(std::string(what) + ": " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
        return int(e);
    }

    // Resolved through the runtime (cudaGetDriverEntryPoint) so the library links without libcuda.
    typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    encode_tiled_fn get_encode_tiled();
}

struct spb_exchange;
namespace spb
{
    // Same-rank neighbour table of an exchange plan for the fused stage kernel (spb_exchange.cu): d_nbr[27*lb + e],
    // e = (ex+1) + 3*(ey+1) + 9*(ez+1), is the local block whose ghost cells receive block lb's cells in direction
    // (ex,ey,ez) through a same-rank transaction, -1 if there is none (off-rank, domain boundary, e = 13).
    // SPB_ERR_UNSUPPORTED if a same-rank transaction is not one of the 26 canonical injection boxes.
    int exchange_fuse_table(spb_exchange* e, const int nx[3], const int ng[3], int64_t nlb, const int** d_nbr);
    // Device list of the local blocks of one part of the overlapped schedule (spb_exchange.cu): part 1 = the source blocks of
    // off-rank sends (injection and interpolation), part 2 = the rest. Built once per plan.
    int exchange_block_list(spb_exchange* e, int64_t nlb, int part, const int** d_list, int64_t* count);
    // Block list the RHS launchers use instead of the contiguous range [lb_begin, lb_end) when set (one launch for a scattered
    // set of blocks: AMR grids, whose rank-boundary blocks are not contiguous in local order). Set and cleared by the C entry
    // point around the launch; thread-local like the rest of the per-thread (= per-GPU) state.
    struct BlockList { const int* dev = nullptr; int64_t count = 0; };
    BlockList& current_block_list();
}

#define SPB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return spb::cuda_fail(e__, #call, __FILE__, __LINE__); } while (0)
#define SPB_LAUNCH_CHECK() do { spb::g_launches.fetch_add(1, std::memory_order_relaxed); cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return spb::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); } while (0)

struct spb_grid
{
    int     nx[3];
    int     ng[3];
    int     np[3];              // padded extents nx + 2 ng
    int64_t nlb;
    int64_t block_stride;       // doubles per block = 5*np0*np1*np2
    int     device;
    std::vector<double> dx_host;      // [nlb][3]
    std::vector<double> inv_dx_host;  // [nlb][3]
    double* inv_dx_dev;               // [nlb][3]
    // relative spread of the spacings of a direction that the rounding of the block bounds alone explains: a block's size is
    // formed as (lo + bsize) - lo (cartesian_blocks.h:66-71), exact to 2 ulp of the COORDINATE, i.e. 2 eps |x|/size relative
    double  spacing_round_tol[3] = {0.0, 0.0, 0.0};
    // refinement levels: the distinct inverse spacings of each direction (values that differ by no more than the rounding
    // tolerance above are one level) and, per block, the packed level indices l0 | l1 << 8 | l2 << 16. lev_n[d] = -1: more
    // than SPB_MAX_LEVELS distinct spacings in direction d (the narrow kernel then refuses a non-uniform launch)
    int     lev_n[3] = {0, 0, 0};
    double  lev_inv[3][16] = {};
    int*    lev_dev = nullptr;
    int     num_sms;
    // general coordinates (spb_grid_set_metric): [nlb][3 directions][3 rows][metric_lm] doubles, rows = area metric,
    // 1/jacobian metric, 1/face metric; null for coords::identity
    double* metric_dev = nullptr;
    int     metric_lm = 0;
    // scratch of spb_reduce (partials | result | counter), owned by the handle so a reduction never allocates
    mutable double* red_scratch = nullptr;
    mutable int64_t red_cap = 0;
    // device copies of the block lists of the six domain boundaries last passed to spb_boundary_fill
    mutable std::vector<int64_t> bnd_blocks_host[6];
    mutable int64_t* bnd_blocks_dev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
