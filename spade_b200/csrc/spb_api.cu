// Grid handle, error string, utilities of the C ABI (include/spade_b200.h).
#include "spb_common.cuh"
#include <mutex>
#include <cmath>
#include <cstring>

namespace spb
{
    BlockList& current_block_list() { thread_local BlockList b; return b; }

    static thread_local std::string t_error;
    std::atomic<int64_t> g_launches{0};
    void set_error(const std::string& msg) { t_error = msg; }

    encode_tiled_fn get_encode_tiled()
    {
        static encode_tiled_fn fn = nullptr;
        static std::once_flag once;
        std::call_once(once, []
        {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess
                && qres == cudaDriverEntryPointSuccess)
                fn = (encode_tiled_fn)p;
        });
        return fn;
    }
}

extern "C"
{
    const char* spb_last_error(void) { return spb::t_error.c_str(); }
    const char* spb_version(void) { return "spade_b200 0.1 (sm_100a)"; }
    int64_t spb_launch_count(void) { return spb::g_launches.load(); }

    int spb_device_count(void)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
        return n;
    }

    int spb_sync(void* stream)
    {
        SPB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        return 0;
    }

    // Host-only: the distinct inverse spacings of each direction. The reference forms dx = box.size/num_cell block by block from
    // the rounded block bounds (cartesian_blocks.h:66-71, cartesian_grid.h:134-135), exact to 2 ulp of the COORDINATE: spacings
    // that differ by no more than 4 eps |x|max / (smallest block size) (at least 8 eps, at most 1e-13) are one level.
    int spb_grid_spacing_levels(const int nx[3], int64_t nlb, const double* bbox_host, int lev_n[3], double* lev_inv,
                                int* lev_of_block, double* round_tol)
    {
        if (!nx || nlb < 0 || (nlb > 0 && !bbox_host) || !lev_n || !lev_inv) { spb::set_error("spb_grid_spacing_levels: bad argument"); return SPB_ERR_BAD_ARG; }
        const double eps = 2.220446049250313e-16;
        for (int64_t lb = 0; lb < nlb && lev_of_block; ++lb) lev_of_block[lb] = 0;
        for (int d = 0; d < 3; ++d)
        {
            if (nx[d] < 1) { spb::set_error("spb_grid_spacing_levels: bad extents"); return SPB_ERR_BAD_ARG; }
            double xmax = 0.0, smin = 0.0;
            for (int64_t lb = 0; lb < nlb; ++lb)
            {
                const double lo = bbox_host[6*lb + 2*d], hi = bbox_host[6*lb + 2*d + 1];
                xmax = fmax(xmax, fmax(fabs(lo), fabs(hi)));
                smin = (lb == 0) ? fabs(hi - lo) : fmin(smin, fabs(hi - lo));
            }
            const double rt = (smin > 0.0) ? 4.0*eps*xmax/smin : 0.0;
            if (round_tol) round_tol[d] = rt;
            const double tol = fmin(fmax(8.0*eps, rt), 1e-13);
            int n = 0;
            for (int l = 0; l < 16; ++l) lev_inv[16*d + l] = 0.0;
            for (int64_t lb = 0; lb < nlb && n >= 0; ++lb)
            {
                const double a = 1.0/((bbox_host[6*lb + 2*d + 1] - bbox_host[6*lb + 2*d])/nx[d]);     // inv_dx as the reference forms it
                int l = 0;
                while (l < n && !(fabs(a - lev_inv[16*d + l]) <= tol*fabs(a))) ++l;
                if (l == n) { if (n == 16) { n = -1; break; } lev_inv[16*d + n++] = a; }
                if (lev_of_block) lev_of_block[lb] |= l << (8*d);
            }
            lev_n[d] = n;
        }
        return 0;
    }

    // reference: src/grid/cartesian_grid.h:114-136 (dx = box.size/num_cell; inv_dx = 1.0/dx)
    int spb_grid_create(spb_grid** out, const int nx[3], const int ng[3], int64_t nlb, const double* bbox_host)
    {
        if (!out || !nx || !ng || nlb < 0 || (nlb > 0 && !bbox_host)) { spb::set_error("spb_grid_create: bad argument"); return SPB_ERR_BAD_ARG; }
        for (int d = 0; d < 3; ++d) if (nx[d] < 1 || ng[d] < 0) { spb::set_error("spb_grid_create: bad extents"); return SPB_ERR_BAD_ARG; }
        if (spb_device_count() < 1) { spb::set_error("spb_grid_create: no CUDA device (this library has no CPU path)"); return SPB_ERR_NO_DEVICE; }
        spb_grid* g = new spb_grid();
        for (int d = 0; d < 3; ++d) { g->nx[d] = nx[d]; g->ng[d] = ng[d]; g->np[d] = nx[d] + 2*ng[d]; }
        g->nlb = nlb;
        g->block_stride = (int64_t)SPB_NVAR*g->np[0]*g->np[1]*g->np[2];
        g->dx_host.resize(3*nlb); g->inv_dx_host.resize(3*nlb);
        for (int64_t lb = 0; lb < nlb; ++lb)
            for (int d = 0; d < 3; ++d)
            {
                const double size = bbox_host[6*lb + 2*d + 1] - bbox_host[6*lb + 2*d];
                const double dx = size/nx[d];
                g->dx_host[3*lb+d] = dx;
                g->inv_dx_host[3*lb+d] = 1.0/dx;
            }
        // refinement levels per direction (an AMR grid has a handful; a uniform lattice one)
        std::vector<int> lev(nlb > 0 ? nlb : 0, 0);
        spb_grid_spacing_levels(nx, nlb, bbox_host, g->lev_n, &g->lev_inv[0][0], nlb > 0 ? lev.data() : nullptr, g->spacing_round_tol);
        g->inv_dx_dev = nullptr;
        cudaError_t e = cudaGetDevice(&g->device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, g->device);
        if (e == cudaSuccess && nlb > 0) e = cudaMalloc(&g->inv_dx_dev, sizeof(double)*3*nlb);
        if (e == cudaSuccess && nlb > 0) e = cudaMemcpy(g->inv_dx_dev, g->inv_dx_host.data(), sizeof(double)*3*nlb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && nlb > 0) e = cudaMalloc((void**)&g->lev_dev, sizeof(int)*nlb);
        if (e == cudaSuccess && nlb > 0) e = cudaMemcpy(g->lev_dev, lev.data(), sizeof(int)*nlb, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { delete g; return spb::cuda_fail(e, "spb_grid_create", __FILE__, __LINE__); }
        *out = g;
        return 0;
    }

    // reference: src/core/coord_system.h:250-267 (calc_normal_vector), 295-302 (calc_jacobian)
    int spb_grid_set_metric(spb_grid* g, const spb_metric_desc* m)
    {
        if (!g) { spb::set_error("spb_grid_set_metric: null grid"); return SPB_ERR_BAD_ARG; }
        if (g->metric_dev) { cudaFree(g->metric_dev); g->metric_dev = nullptr; g->metric_lm = 0; }
        if (!m) return 0;
        int lm = 0;
        for (int d = 0; d < 3; ++d)
        {
            if (!m->area[d] || !m->jac[d] || !m->face[d]) { spb::set_error("spb_grid_set_metric: null table"); return SPB_ERR_BAD_ARG; }
            lm = g->np[d] + 1 > lm ? g->np[d] + 1 : lm;
        }
        std::vector<double> tab((size_t)g->nlb*9*lm, 1.0);
        for (int64_t lb = 0; lb < g->nlb; ++lb)
            for (int d = 0; d < 3; ++d)
            {
                double* row = tab.data() + ((size_t)lb*3 + d)*3*lm;
                // Only the entries of interior cells and of their faces are ever read (a face's tangential metrics are those of
                // the interior cell row it belongs to). A mapping may fold or flatten beyond the domain boundary
                // (integrated_tanh_1D does): entries out there are kept when usable and neutralised otherwise.
                const int lo = g->ng[d], hi = g->ng[d] + g->nx[d];
                auto usable = [](const double x) { return x != 0.0 && std::isfinite(x); };
                for (int i = 0; i < g->np[d]; ++i)
                {
                    const double a = m->area[d][lb*g->np[d] + i], j = m->jac[d][lb*g->np[d] + i];
                    const bool used = i >= lo && i < hi;
                    if (used && (!usable(a) || !usable(j))) { spb::set_error("spb_grid_set_metric: coordinate derivatives of interior cells must be finite and nonzero"); return SPB_ERR_BAD_ARG; }
                    row[i] = usable(a) ? a : 1.0;
                    row[lm + i] = usable(j) ? 1.0/j : 1.0;
                }
                for (int i = 0; i <= g->np[d]; ++i)
                {
                    const double f = m->face[d][lb*(g->np[d] + 1) + i];
                    const bool used = i >= lo && i <= hi;
                    if (used && !usable(f)) { spb::set_error("spb_grid_set_metric: coordinate derivatives at the faces of interior cells must be finite and nonzero"); return SPB_ERR_BAD_ARG; }
                    row[2*lm + i] = usable(f) ? 1.0/f : 1.0;
                }
            }
        if (g->nlb == 0) return 0;
        SPB_CUDA(cudaMalloc(&g->metric_dev, sizeof(double)*tab.size()));
        SPB_CUDA(cudaMemcpy(g->metric_dev, tab.data(), sizeof(double)*tab.size(), cudaMemcpyHostToDevice));
        g->metric_lm = lm;
        return 0;
    }
    int spb_grid_has_metric(const spb_grid* g) { return (g && g->metric_dev) ? 1 : 0; }

    // ---- peer-memory plumbing for one-process-per-GPU runs: raw (IPC-shareable) allocations, IPC handles, stream-ordered flags
    int spb_dev_alloc(void** out, size_t bytes)
    {
        if (!out) { spb::set_error("spb_dev_alloc: null argument"); return SPB_ERR_BAD_ARG; }
        SPB_CUDA(cudaMalloc(out, bytes ? bytes : 8));
        SPB_CUDA(cudaMemset(*out, 0, bytes ? bytes : 8));
        return 0;
    }
    int spb_dev_free(void* p) { if (p) SPB_CUDA(cudaFree(p)); return 0; }
    int spb_ipc_export(const void* dev_ptr, unsigned char handle[64])
    {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        cudaIpcMemHandle_t h;
        SPB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
        memcpy(handle, &h, 64);
        return 0;
    }
    int spb_ipc_import(const unsigned char handle[64], void** out)
    {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, 64);
        SPB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
        return 0;
    }
    int spb_ipc_close(void* p) { if (p) SPB_CUDA(cudaIpcCloseMemHandle(p)); return 0; }

    void spb_grid_destroy(spb_grid* g)
    {
        if (!g) return;
        if (g->metric_dev) cudaFree(g->metric_dev);
        if (g->inv_dx_dev) cudaFree(g->inv_dx_dev);
        if (g->lev_dev) cudaFree(g->lev_dev);
        if (g->red_scratch) cudaFree(g->red_scratch);
        for (int i = 0; i < 6; ++i) if (g->bnd_blocks_dev[i]) cudaFree(g->bnd_blocks_dev[i]);
        delete g;
    }

    int64_t spb_grid_array_size(const spb_grid* g) { return g->block_stride*g->nlb; }

    // reference: src/core/mem_map.h:484-496
    int64_t spb_grid_offset(const spb_grid* g, int v, int i, int j, int k, int64_t lb)
    {
        return v + (int64_t)SPB_NVAR*((i + g->ng[0]) + (int64_t)g->np[0]*((j + g->ng[1]) + (int64_t)g->np[1]*((k + g->ng[2]) + (int64_t)g->np[2]*lb)));
    }
}
