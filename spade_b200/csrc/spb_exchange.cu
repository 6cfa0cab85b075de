// Ghost-cell exchange: plan construction (host) and the pack / unpack / same-rank copy kernels.
// Replaces get_exchg_config + arr_exchange_t::exchange of the reference
// (src/grid/exchange_config.h:286-419, src/grid/get_transaction.h:11-97,
//  src/grid/cartesian_blocks.h:76-103, src/grid/partition.h:27-84, src/grid/make_exchange.h:111-410).
//
// The reference finds the transaction of a buffer element with a 16-way tag search and div/mod per
// element (make_exchange.h:14-63). Here the plan is flattened once on the host into work items
// (transaction, element range); a CTA streams one item with contiguous runs along i.
#include "spb_common.cuh"
#include <algorithm>
#include <cstdlib>

struct spb_trans
{
    int64_t f[16];   // tag, rank_send, rank_recv, glob_src, glob_dst, src.min[4], src.size[3], dst.min[4]
    // interpolation (patch_fill_t) only: dest.size[3], i_coeff[3], i_incr[3]; the source box is then the donor region
    int64_t dsize[3] = {0, 0, 0}, ic[3] = {0, 0, 0}, inc[3] = {0, 0, 0};
    bool interp = false;
    // cells of the message = destination volume (a patch_fill_t packs into the size the receiver expects, transactions.h:149-153)
    int64_t cells() const { return interp ? dsize[0]*dsize[1]*dsize[2] : f[9]*f[10]*f[11]; }
};

namespace spb
{
    struct DevTrans
    {
        long long src_off;   // offset (doubles) of the source box origin in q, -1 if not local
        long long dst_off;   // offset (doubles) of the destination box origin in q, -1 if not local
        long long buf_off;   // offset (doubles) of this transaction inside its peer message
        int bx, by, bz;      // destination box (= source box for injections)
        int interp;          // 1: patch_fill_t (AMR): destination cell = mean of 2^3 donors of the source box
        int ic[3], inc[3];   // i_coeff, i_incr of the patch_fill_t (reference src/grid/transactions.h:147)
    };
    struct WorkItem { int trans; int begin; };   // element range [begin, begin + ITEM) of a transaction

    constexpr int ITEM = 2048;                   // doubles per work item
}

struct spb_exchange
{
    int nx[3], ng[3], np[3];
    int rank, nranks;
    int64_t nlocal, first_block;
    std::vector<spb_trans> send, recv;                 // injection lists, sorted like the reference
    std::vector<spb_trans> isend, irecv;               // interpolation lists (AMR), same order rule; empty on uniform lattices
    std::vector<int64_t> inj_send_cells, inj_recv_cells;   // per peer: the injection section that precedes the interpolation section
    std::vector<int64_t> send_rank_off, send_rank_cnt, recv_rank_off, recv_rank_cnt;
    std::vector<int64_t> send_cells, recv_cells;       // per peer message size in cells
    // device side (built lazily on the first device call)
    bool dev_ready = false;
    spb::DevTrans* d_send = nullptr; spb::DevTrans* d_recv = nullptr;
    std::vector<std::vector<spb::WorkItem>> items_send, items_recv;   // per peer
    std::vector<spb::WorkItem*> d_items_send, d_items_recv;
    // same-rank neighbour table for the fused stage kernel: 0 = not built, 1 = built, -1 = plan not canonical
    int fuse_state = 0;
    int* d_nbr = nullptr;
    int64_t nbr_blocks = 0;
    // same-rank interpolation work items only (what is left to do after the stage kernel stored the injection ghosts)
    std::vector<spb::WorkItem> items_local_itp;
    spb::WorkItem* d_items_local_itp = nullptr;
    // device block lists of the overlapped schedule: [0] rank-boundary blocks, [1] the rest
    int* d_blist[2] = {nullptr, nullptr};
    int64_t blist_count[2] = {0, 0};
    int64_t blist_nlb = -1;
};

namespace spb
{
    // reference src/grid/partition.h:40-71
    static void partition_blocks(int64_t nglob, int nranks, std::vector<int64_t>& g2r, std::vector<int64_t>& g2l)
    {
        g2r.assign(nglob, 0); g2l.assign(nglob, 0);
        std::vector<int64_t> partial(nranks, 0);
        int64_t cur = 0, counter = 0;
        const int64_t per = nglob/nranks, extra = nglob - per*nranks;
        for (int64_t lb = 0; lb < nglob; ++lb)
        {
            g2r[lb] = cur; g2l[lb] = partial[cur]++;
            ++counter;
            const int64_t add = (extra > 0 && cur < extra) ? 1 : 0;
            if (counter == per + add) { counter = 0; cur = (cur + 1) % nranks; }
        }
    }

    static void finish_plan(spb_exchange* e)
    {
        // stable sort by (peer asc, tag desc): reference src/grid/exchange_config.h:379-397
        auto by_recv = [](const spb_trans& a, const spb_trans& b)
            { if (a.f[2] != b.f[2]) return a.f[2] < b.f[2]; if (a.f[0] != b.f[0]) return a.f[0] > b.f[0]; return a.cells() < b.cells(); };
        auto by_send = [](const spb_trans& a, const spb_trans& b)
            { if (a.f[1] != b.f[1]) return a.f[1] < b.f[1]; if (a.f[0] != b.f[0]) return a.f[0] > b.f[0]; return a.cells() < b.cells(); };
        std::stable_sort(e->send.begin(), e->send.end(), by_recv);
        std::stable_sort(e->recv.begin(), e->recv.end(), by_send);
        std::stable_sort(e->isend.begin(), e->isend.end(), by_recv);
        std::stable_sort(e->irecv.begin(), e->irecv.end(), by_send);
        const int n = e->nranks;
        e->send_rank_off.assign(n, 0); e->send_rank_cnt.assign(n, 0); e->recv_rank_off.assign(n, 0); e->recv_rank_cnt.assign(n, 0);
        e->send_cells.assign(n, 0); e->recv_cells.assign(n, 0);
        for (const auto& t: e->send) { e->send_rank_cnt[t.f[2]]++; e->send_cells[t.f[2]] += t.cells(); }
        for (const auto& t: e->recv) { e->recv_rank_cnt[t.f[1]]++; e->recv_cells[t.f[1]] += t.cells(); }
        e->inj_send_cells = e->send_cells; e->inj_recv_cells = e->recv_cells;
        // a peer message = injection section, then interpolation section (make_exchange.h:133-135,208-209)
        for (const auto& t: e->isend) e->send_cells[t.f[2]] += t.cells();
        for (const auto& t: e->irecv) e->recv_cells[t.f[1]] += t.cells();
        int64_t so = 0, ro = 0;
        for (int p = 0; p < n; ++p) { e->send_rank_off[p] = so; so += e->send_rank_cnt[p]; e->recv_rank_off[p] = ro; ro += e->recv_rank_cnt[p]; }
        e->dev_ready = false;
    }

    static long long box_origin(const spb_exchange* e, const int64_t* mn)   // mn = {i, j, k, lb}
    {
        if (mn[3] < 0) return -1;
        return 5ll*((mn[0] + e->ng[0]) + (long long)e->np[0]*((mn[1] + e->ng[1]) + (long long)e->np[1]*((mn[2] + e->ng[2]) + (long long)e->np[2]*mn[3])));
    }

    static void free_device(spb_exchange* e)
    {
        if (e->d_send) cudaFree(e->d_send);
        if (e->d_recv) cudaFree(e->d_recv);
        for (auto p: e->d_items_send) if (p) cudaFree(p);
        for (auto p: e->d_items_recv) if (p) cudaFree(p);
        if (e->d_items_local_itp) { cudaFree(e->d_items_local_itp); e->d_items_local_itp = nullptr; }
        e->d_send = nullptr; e->d_recv = nullptr; e->d_items_send.clear(); e->d_items_recv.clear();
    }

    static int build_device(spb_exchange* e)
    {
        if (e->dev_ready) return 0;
        free_device(e);
        const int n = e->nranks;
        // device list = injection transactions followed by interpolation transactions; the work items of a peer run over
        // its injection section first, then its interpolation section, like the reference's message layout
        auto build = [&](const std::vector<spb_trans>& inj, const std::vector<spb_trans>& itp, bool is_send, const std::vector<int64_t>& inj_cells,
                         DevTrans** d_list, std::vector<std::vector<WorkItem>>& items, std::vector<WorkItem*>& d_items) -> int
        {
            std::vector<DevTrans> h(inj.size() + itp.size());
            items.assign(n, {}); d_items.assign(n, nullptr);
            std::vector<long long> boff(n, 0);
            auto add = [&](const spb_trans& tr, size_t t, long long base)
            {
                const int p = (int)(is_send ? tr.f[2] : tr.f[1]);
                DevTrans& d = h[t];
                d.src_off = box_origin(e, &tr.f[5]);
                d.dst_off = box_origin(e, &tr.f[12]);
                d.buf_off = base + boff[p];
                d.interp = tr.interp ? 1 : 0;
                for (int k = 0; k < 3; ++k) { d.ic[k] = (int)tr.ic[k]; d.inc[k] = (int)tr.inc[k]; }
                d.bx = (int)(tr.interp ? tr.dsize[0] : tr.f[9]); d.by = (int)(tr.interp ? tr.dsize[1] : tr.f[10]); d.bz = (int)(tr.interp ? tr.dsize[2] : tr.f[11]);
                const long long nel = 5ll*tr.cells();
                boff[p] += nel;
                for (long long b = 0; b < nel; b += ITEM) items[p].push_back(WorkItem{(int)t, (int)b});
            };
            for (size_t t = 0; t < inj.size(); ++t) add(inj[t], t, 0);
            std::fill(boff.begin(), boff.end(), 0);
            for (size_t t = 0; t < itp.size(); ++t)
            {
                const int p = (int)(is_send ? itp[t].f[2] : itp[t].f[1]);
                add(itp[t], inj.size() + t, 5ll*inj_cells[p]);
            }
            for (int p = 0; p < n; ++p)
                if (!items[p].empty())
                {
                    SPB_CUDA(cudaMalloc((void**)&d_items[p], sizeof(WorkItem)*items[p].size()));
                    SPB_CUDA(cudaMemcpy(d_items[p], items[p].data(), sizeof(WorkItem)*items[p].size(), cudaMemcpyHostToDevice));
                }
            if (!h.empty())
            {
                SPB_CUDA(cudaMalloc((void**)d_list, sizeof(DevTrans)*h.size()));
                SPB_CUDA(cudaMemcpy(*d_list, h.data(), sizeof(DevTrans)*h.size(), cudaMemcpyHostToDevice));
            }
            return 0;
        };
        int rc = build(e->send, e->isend, true, e->inj_send_cells, &e->d_send, e->items_send, e->d_items_send);
        if (rc) return rc;
        // the same-rank interpolation items alone: device transaction ids continue after the injection list
        e->items_local_itp.clear();
        for (const auto& it: e->items_send[e->rank]) if (it.trans >= (int)e->send.size()) e->items_local_itp.push_back(it);
        if (!e->items_local_itp.empty())
        {
            SPB_CUDA(cudaMalloc((void**)&e->d_items_local_itp, sizeof(WorkItem)*e->items_local_itp.size()));
            SPB_CUDA(cudaMemcpy(e->d_items_local_itp, e->items_local_itp.data(), sizeof(WorkItem)*e->items_local_itp.size(), cudaMemcpyHostToDevice));
        }
        rc = build(e->recv, e->irecv, false, e->inj_recv_cells, &e->d_recv, e->items_recv, e->d_items_recv);
        if (rc) return rc;
        e->dev_ready = true;
        return 0;
    }

    int exchange_fuse_table(spb_exchange* e, const int nx[3], const int ng[3], int64_t nlb, const int** d_nbr)
    {
        if (!e || !d_nbr) { set_error("exchange_fuse_table: null argument"); return SPB_ERR_BAD_ARG; }
        for (int d = 0; d < 3; ++d)
            if (e->nx[d] != nx[d] || e->ng[d] != ng[d]) { set_error("fused exchange: the plan was made for another block shape"); return SPB_ERR_BAD_ARG; }
        if (e->fuse_state == 0 || (e->fuse_state == 1 && e->nbr_blocks != nlb))
        {
            if (e->d_nbr) { cudaFree(e->d_nbr); e->d_nbr = nullptr; }
            std::vector<int> nbr(27*(size_t)std::max<int64_t>(nlb, 1), -1);
            bool ok = true;
            for (const auto& t: e->send)
            {
                if (t.f[2] != e->rank) continue;                       // off-rank: packed into a message
                int code = 0, mul = 1;
                for (int d = 0; d < 3 && ok; ++d)
                {
                    const int64_t smin = t.f[5+d], size = t.f[9+d], dmin = t.f[12+d];
                    int ed;
                    if (size == nx[d] && smin == 0 && dmin == 0) ed = 0;
                    else if (size == ng[d] && smin == 0 && dmin == nx[d]) ed = -1;
                    else if (size == ng[d] && smin == nx[d] - ng[d] && dmin == -ng[d]) ed = 1;
                    else { ok = false; break; }
                    code += (ed + 1)*mul; mul *= 3;
                }
                const int64_t src = t.f[8], dst = t.f[15];
                if (!ok || code == 13 || src < 0 || src >= nlb || dst < 0 || dst >= nlb || nbr[27*src + code] != -1) { ok = false; break; }
                nbr[27*src + code] = (int)dst;
            }
            if (!ok) e->fuse_state = -1;
            else
            {
                SPB_CUDA(cudaMalloc((void**)&e->d_nbr, sizeof(int)*nbr.size()));
                SPB_CUDA(cudaMemcpy(e->d_nbr, nbr.data(), sizeof(int)*nbr.size(), cudaMemcpyHostToDevice));
                e->nbr_blocks = nlb;
                e->fuse_state = 1;
            }
        }
        if (e->fuse_state != 1)
        {
            set_error("fused exchange: the plan holds same-rank transactions that are not canonical injection boxes (use spb_exchange_local)");
            return SPB_ERR_UNSUPPORTED;
        }
        *d_nbr = e->d_nbr;
        return 0;
    }

    int exchange_block_list(spb_exchange* e, int64_t nlb, int part, const int** d_list, int64_t* count)
    {
        if (!e || !d_list || !count || (part != 1 && part != 2) || nlb < 0) { set_error("exchange_block_list: bad argument"); return SPB_ERR_BAD_ARG; }
        if (e->blist_nlb != nlb)
        {
            std::vector<unsigned char> mask((size_t)std::max<int64_t>(nlb, 1), 0);
            int rc = spb_exchange_boundary_blocks(e, nlb, mask.data()); if (rc) return rc;
            std::vector<int> lists[2];
            for (int64_t b = 0; b < nlb; ++b) lists[mask[b] ? 0 : 1].push_back((int)b);
            for (int k = 0; k < 2; ++k)
            {
                if (e->d_blist[k]) { cudaFree(e->d_blist[k]); e->d_blist[k] = nullptr; }
                e->blist_count[k] = (int64_t)lists[k].size();
                if (lists[k].empty()) continue;
                SPB_CUDA(cudaMalloc((void**)&e->d_blist[k], sizeof(int)*lists[k].size()));
                SPB_CUDA(cudaMemcpy(e->d_blist[k], lists[k].data(), sizeof(int)*lists[k].size(), cudaMemcpyHostToDevice));
            }
            e->blist_nlb = nlb;
        }
        *d_list = e->d_blist[part - 1];
        *count = e->blist_count[part - 1];
        return 0;
    }

    // MODE 0: q(dst) = q(src)   MODE 1: buf = q(src)   MODE 2: q(dst) = buf
    template <int MODE>
    __global__ void __launch_bounds__(256) exchange_kernel(double* q, const double* qsrc, double* __restrict__ buf,
        const double* __restrict__ bufsrc, const DevTrans* __restrict__ trans, const WorkItem* __restrict__ items, int np0, int np1)
    {
        const WorkItem it = items[blockIdx.x];
        const DevTrans tr = trans[it.trans];
        const int nel = 5*tr.bx*tr.by*tr.bz;
        const int row = 5*tr.bx;                       // contiguous doubles along i
        const long long pitch_j = 5ll*np0, pitch_k = 5ll*np0*np1;
        const int end = min(it.begin + ITEM, nel);
        // 16-byte path (uniform per work item): rows of an even number of doubles that start on even offsets — every box of a
        // lattice with an even padded row (2 exchange cells, even block sizes). A pair (el, el+1) with el even never leaves its row.
        const bool even = !tr.interp && ((row | (5*np0) | nel) & 1) == 0
            && (MODE == 2 || ((tr.src_off & 1) == 0 && (reinterpret_cast<uintptr_t>(qsrc) & 15) == 0))
            && (MODE == 1 || ((tr.dst_off & 1) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0))
            && (MODE == 0 || ((tr.buf_off & 1) == 0 && (reinterpret_cast<uintptr_t>(MODE == 1 ? buf : bufsrc) & 15) == 0));
        if (even)
        {
            for (int el = it.begin + 2*threadIdx.x; el < end; el += 512)
            {
                const int r = el / row, c = el - r*row;
                const int iz = r / tr.by, iy = r - iz*tr.by;
                const long long rel = c + iy*pitch_j + iz*pitch_k;
                const double2 val = (MODE == 2) ? *reinterpret_cast<const double2*>(bufsrc + tr.buf_off + el)
                                                : *reinterpret_cast<const double2*>(qsrc + tr.src_off + rel);
                if (MODE == 1) *reinterpret_cast<double2*>(buf + tr.buf_off + el) = val;
                else           *reinterpret_cast<double2*>(q + tr.dst_off + rel) = val;
            }
            return;
        }
        if (MODE != 2 && tr.interp)
        {
            // patch_fill_t (make_exchange.h:233-256, transactions.h:176-201), one destination CELL per thread: donor offset
            // ((ix << (i_coeff+1)) >> 1) + d*i_incr; per variable the 2^3 donors are added in the order d0 fastest and the sum is
            // multiplied by 1/8 (bit-identical to the reference). The index arithmetic is done once per cell, the 8 x 5 loads are
            // independent. A cell that straddles two work items is completed by both (each stores its own variables).
            for (int cell = it.begin/5 + threadIdx.x; 5*cell < end; cell += 256)
            {
                const int r = cell / tr.bx, ix = cell - r*tr.bx;
                const int iz = r / tr.by, iy = r - iz*tr.by;
                const int i0 = (ix << (tr.ic[0] + 1)) >> 1, j0 = (iy << (tr.ic[1] + 1)) >> 1, k0 = (iz << (tr.ic[2] + 1)) >> 1;
                const double* d0p = qsrc + tr.src_off + 5ll*i0 + j0*pitch_j + k0*pitch_k;
                const long long si = 5ll*tr.inc[0], sj = tr.inc[1]*pitch_j, sk = tr.inc[2]*pitch_k;
                double sum[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                #pragma unroll
                for (int iv = 0; iv < 8; ++iv)
                {
                    const double* dp = d0p + (iv & 1)*si + ((iv >> 1) & 1)*sj + ((iv >> 2) & 1)*sk;
                    #pragma unroll
                    for (int v = 0; v < 5; ++v) sum[v] = __dadd_rn(sum[v], dp[v]);
                }
                const long long rel = 5ll*ix + iy*pitch_j + iz*pitch_k;
                #pragma unroll
                for (int v = 0; v < 5; ++v)
                {
                    const int el = 5*cell + v;
                    if (el < it.begin || el >= end) continue;
                    if (MODE == 1) buf[tr.buf_off + el] = sum[v]*0.125;
                    else           q[tr.dst_off + rel + v] = sum[v]*0.125;
                }
            }
            return;
        }
        for (int el = it.begin + threadIdx.x; el < end; el += 256)
        {
            const int r = el / row, c = el - r*row;
            const int iz = r / tr.by, iy = r - iz*tr.by;
            const long long rel = c + iy*pitch_j + iz*pitch_k;
            const double val = (MODE == 2) ? bufsrc[tr.buf_off + el] : qsrc[tr.src_off + rel];
            if (MODE == 1) buf[tr.buf_off + el] = val;
            else q[tr.dst_off + rel] = val;
        }
    }
}

namespace spb
{
    // the stores of the kernels that precede this one in the stream have been performed (kernel boundary); the fence orders
    // them before the flag for an observer on another GPU
    __global__ void flag_signal_kernel(unsigned long long* flag, const unsigned long long value)
    {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(flag) = value;
        __threadfence_system();
    }
    __global__ void flag_wait_kernel(const unsigned long long* flag, const unsigned long long value)
    {
        const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(flag);
        while (*f < value) __nanosleep(200);
        __threadfence_system();
    }
}

extern "C"
{
    int spb_exchange_create(spb_exchange** out, const int nb[3], const int nx[3], const int ng[3], const int periodic[3], int rank, int nranks)
    {
        using namespace spb;
        if (!out || !nb || !nx || !ng || !periodic || nranks < 1 || rank < 0 || rank >= nranks) { set_error("spb_exchange_create: bad argument"); return SPB_ERR_BAD_ARG; }
        spb_exchange* e = new spb_exchange();
        for (int d = 0; d < 3; ++d) { e->nx[d] = nx[d]; e->ng[d] = ng[d]; e->np[d] = nx[d] + 2*ng[d]; }
        e->rank = rank; e->nranks = nranks;
        const int64_t nglob = (int64_t)nb[0]*nb[1]*nb[2];
        std::vector<int64_t> g2r, g2l;
        partition_blocks(nglob, nranks, g2r, g2l);
        e->nlocal = 0; e->first_block = -1;
        for (int64_t lb = 0; lb < nglob; ++lb) if (g2r[lb] == rank) { if (e->first_block < 0) e->first_block = lb; e->nlocal++; }
        if (e->first_block < 0) e->first_block = 0;
        // transactions in (global block asc, neighbour-table order): reference exchange_config.h:331-363
        for (int64_t lb = 0; lb < nglob; ++lb)
        {
            const int blk[3] = {(int)(lb % nb[0]), (int)((lb/nb[0]) % nb[1]), (int)(lb/((int64_t)nb[0]*nb[1]))};
            for (int dk = -1; dk <= 1; ++dk) for (int dj = -1; dj <= 1; ++dj) for (int di = -1; di <= 1; ++di)
            {
                if (di == 0 && dj == 0 && dk == 0) continue;
                const int edge[3] = {di, dj, dk};
                int nbr[3]; bool ignore = false;
                for (int d = 0; d < 3; ++d)
                {
                    nbr[d] = blk[d] + edge[d];                 // periodic wrap always applied, cartesian_blocks.h:84-92
                    if (nbr[d] < 0) nbr[d] += nb[d];
                    if (nbr[d] >= nb[d]) nbr[d] -= nb[d];
                    const bool is_min = blk[d] == 0, is_max = blk[d] == nb[d] - 1;
                    if (!periodic[d] && ((is_min && edge[d] == -1) || (is_max && edge[d] == 1))) ignore = true;   // exchange_config.h:336-342
                }
                if (ignore) continue;
                const int64_t lbn = nbr[0] + (int64_t)nb[0]*(nbr[1] + (int64_t)nb[1]*nbr[2]);
                if (g2r[lb] != rank && g2r[lbn] != rank) continue;
                spb_trans t;
                const int dsum = std::abs(di) + std::abs(dj) + std::abs(dk);
                int tag = dsum == 1 ? 3 : (dsum == 2 ? 12 : 16);  // get_transaction.h:33-51
                for (int d = 0; d < 3; ++d)
                {
                    if (dsum == 1 && edge[d] != 0) tag -= d;
                    if (dsum == 2 && edge[d] == 0) tag -= d;
                }
                t.f[0] = tag; t.f[1] = g2r[lb]; t.f[2] = g2r[lbn]; t.f[3] = lb; t.f[4] = lbn;
                for (int d = 0; d < 3; ++d)                       // get_transaction.h:54-86
                {
                    int smin, smax, dmin;
                    if (edge[d] == -1)     { smin = 0;             smax = ng[d]; dmin = nx[d]; }
                    else if (edge[d] == 0) { smin = 0;             smax = nx[d]; dmin = 0; }
                    else                   { smin = nx[d] - ng[d]; smax = nx[d]; dmin = -ng[d]; }
                    t.f[5+d] = smin; t.f[9+d] = smax - smin; t.f[12+d] = dmin;
                }
                t.f[8]  = g2r[lb]  == rank ? g2l[lb]  : -1;       // partition.h:107-111 (to_local of a foreign block)
                t.f[15] = g2r[lbn] == rank ? g2l[lbn] : -1;
                if (g2r[lb]  == rank) e->send.push_back(t);
                if (g2r[lbn] == rank) e->recv.push_back(t);
            }
        }
        finish_plan(e);
        *out = e;
        return 0;
    }

    int spb_exchange_create_from_tables(spb_exchange** out, const int nx[3], const int ng[3], int rank, int nranks,
                                        const int64_t* send, int64_t nsend, const int64_t* recv, int64_t nrecv)
    {
        using namespace spb;
        if (!out || !nx || !ng || nranks < 1 || rank < 0 || rank >= nranks || nsend < 0 || nrecv < 0) { set_error("spb_exchange_create_from_tables: bad argument"); return SPB_ERR_BAD_ARG; }
        spb_exchange* e = new spb_exchange();
        for (int d = 0; d < 3; ++d) { e->nx[d] = nx[d]; e->ng[d] = ng[d]; e->np[d] = nx[d] + 2*ng[d]; }
        e->rank = rank; e->nranks = nranks; e->nlocal = -1; e->first_block = -1;
        e->send.resize(nsend); e->recv.resize(nrecv);
        for (int64_t i = 0; i < nsend; ++i) std::copy(send + 16*i, send + 16*i + 16, e->send[i].f);
        for (int64_t i = 0; i < nrecv; ++i) std::copy(recv + 16*i, recv + 16*i + 16, e->recv[i].f);
        // the kernels trust these tables: ranks inside the group, boxes inside the padded block, local ids >= -1
        for (const auto* list: {&e->send, &e->recv})
            for (const auto& t: *list)
            {
                bool ok = t.f[1] >= 0 && t.f[1] < nranks && t.f[2] >= 0 && t.f[2] < nranks && t.f[8] >= -1 && t.f[15] >= -1;
                for (int d = 0; d < 3 && ok; ++d)
                    ok = t.f[9+d] >= 0 && t.f[5+d] >= -ng[d] && t.f[5+d] + t.f[9+d] <= nx[d] + ng[d]
                                       && t.f[12+d] >= -ng[d] && t.f[12+d] + t.f[9+d] <= nx[d] + ng[d];
                if (!ok) { set_error("spb_exchange_create_from_tables: a transaction has a rank outside the group or a box outside the padded block"); delete e; return SPB_ERR_BAD_ARG; }
            }
        finish_plan(e);
        *out = e;
        return 0;
    }

    int spb_exchange_add_interp(spb_exchange* e, const int64_t* send, int64_t nsend, const int64_t* recv, int64_t nrecv)
    {
        using namespace spb;
        if (!e || nsend < 0 || nrecv < 0 || (nsend > 0 && !send) || (nrecv > 0 && !recv)) { set_error("spb_exchange_add_interp: bad argument"); return SPB_ERR_BAD_ARG; }
        auto fill = [](std::vector<spb_trans>& list, const int64_t* tab, int64_t n)
        {
            list.resize(n);
            for (int64_t i = 0; i < n; ++i)
            {
                const int64_t* r = tab + 26*i;
                std::copy(r, r + 16, list[i].f);
                for (int d = 0; d < 3; ++d) { list[i].dsize[d] = r[16+d]; list[i].ic[d] = r[19+d]; list[i].inc[d] = r[22+d]; }
                list[i].interp = true;
            }
        };
        for (int64_t i = 0; i < nsend + nrecv; ++i)
        {
            const int64_t* r = (i < nsend) ? send + 26*i : recv + 26*(i - nsend);
            for (int d = 0; d < 3; ++d)
                if (r[19+d] < -1 || r[19+d] > 1 || r[22+d] < 0 || r[22+d] > 1 || r[16+d] < 0) { set_error("spb_exchange_add_interp: bad i_coeff / i_incr / size"); return SPB_ERR_BAD_ARG; }
            if (r[1] < 0 || r[1] >= e->nranks || r[2] < 0 || r[2] >= e->nranks) { set_error("spb_exchange_add_interp: rank out of range"); return SPB_ERR_BAD_ARG; }
        }
        fill(e->isend, send, nsend);
        fill(e->irecv, recv, nrecv);
        finish_plan(e);
        e->fuse_state = 0;
        return 0;
    }

    int64_t spb_exchange_num_interp_send(const spb_exchange* e) { return (int64_t)e->isend.size(); }
    int64_t spb_exchange_num_interp_recv(const spb_exchange* e) { return (int64_t)e->irecv.size(); }

    void spb_exchange_destroy(spb_exchange* e)
    {
        if (!e) return;
        spb::free_device(e);
        if (e->d_nbr) cudaFree(e->d_nbr);
        for (int k = 0; k < 2; ++k) if (e->d_blist[k]) cudaFree(e->d_blist[k]);
        delete e;
    }

    int64_t spb_exchange_num_send(const spb_exchange* e) { return (int64_t)e->send.size(); }
    int64_t spb_exchange_num_recv(const spb_exchange* e) { return (int64_t)e->recv.size(); }
    int64_t spb_exchange_local_blocks(const spb_exchange* e) { return e->nlocal; }
    int64_t spb_exchange_first_block(const spb_exchange* e) { return e->first_block; }
    int64_t spb_exchange_send_cells(const spb_exchange* e, int peer) { return (peer < 0 || peer >= e->nranks) ? -1 : e->send_cells[peer]; }
    int64_t spb_exchange_recv_cells(const spb_exchange* e, int peer) { return (peer < 0 || peer >= e->nranks) ? -1 : e->recv_cells[peer]; }

    int spb_exchange_tables(const spb_exchange* e, int64_t* send, int64_t* recv, int64_t* offs)
    {
        if (!e) { spb::set_error("spb_exchange_tables: null plan"); return SPB_ERR_BAD_ARG; }
        if (send) for (size_t i = 0; i < e->send.size(); ++i) std::copy(e->send[i].f, e->send[i].f + 16, send + 16*i);
        if (recv) for (size_t i = 0; i < e->recv.size(); ++i) std::copy(e->recv[i].f, e->recv[i].f + 16, recv + 16*i);
        if (offs) for (int p = 0; p < e->nranks; ++p)
        {
            offs[6*p+0] = e->send_cells[p];    offs[6*p+1] = e->recv_cells[p];
            offs[6*p+2] = e->send_rank_off[p]; offs[6*p+3] = e->send_rank_cnt[p];
            offs[6*p+4] = e->recv_rank_off[p]; offs[6*p+5] = e->recv_rank_cnt[p];
        }
        return 0;
    }

    int spb_exchange_boundary_blocks(const spb_exchange* e, int64_t nlb, unsigned char* mask)
    {
        if (!e || !mask || nlb < 0) { spb::set_error("spb_exchange_boundary_blocks: bad argument"); return SPB_ERR_BAD_ARG; }
        std::fill(mask, mask + nlb, (unsigned char)0);
        for (const auto* list: {&e->send, &e->isend})
            for (const auto& t: *list)
            {
                if (t.f[2] == e->rank) continue;
                if (t.f[8] < 0 || t.f[8] >= nlb) { spb::set_error("spb_exchange_boundary_blocks: a send transaction names a block outside [0, nlb)"); return SPB_ERR_BAD_ARG; }
                mask[t.f[8]] = 1;
            }
        return 0;
    }

    int spb_exchange_local(spb_exchange* e, double* q_dev, void* stream)
    {
        using namespace spb;
        if (!e || !q_dev) { set_error("spb_exchange_local: bad argument"); return SPB_ERR_BAD_ARG; }
        int rc = build_device(e); if (rc) return rc;
        const auto& items = e->items_send[e->rank];
        if (items.empty()) return 0;
        exchange_kernel<0><<<(unsigned)items.size(), 256, 0, (cudaStream_t)stream>>>(q_dev, q_dev, nullptr, nullptr, e->d_send, e->d_items_send[e->rank], e->np[0], e->np[1]);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_exchange_local_interp(spb_exchange* e, double* q_dev, void* stream)
    {
        using namespace spb;
        if (!e || !q_dev) { set_error("spb_exchange_local_interp: bad argument"); return SPB_ERR_BAD_ARG; }
        int rc = build_device(e); if (rc) return rc;
        if (e->items_local_itp.empty()) return 0;
        exchange_kernel<0><<<(unsigned)e->items_local_itp.size(), 256, 0, (cudaStream_t)stream>>>(q_dev, q_dev, nullptr, nullptr, e->d_send, e->d_items_local_itp, e->np[0], e->np[1]);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_exchange_pack(spb_exchange* e, const double* q_dev, int peer, double* sendbuf_dev, void* stream)
    {
        using namespace spb;
        if (!e || !q_dev || peer < 0 || peer >= e->nranks) { set_error("spb_exchange_pack: bad argument"); return SPB_ERR_BAD_ARG; }
        int rc = build_device(e); if (rc) return rc;
        const auto& items = e->items_send[peer];
        if (items.empty()) return 0;
        if (!sendbuf_dev) { set_error("spb_exchange_pack: null buffer"); return SPB_ERR_BAD_ARG; }
        exchange_kernel<1><<<(unsigned)items.size(), 256, 0, (cudaStream_t)stream>>>(nullptr, q_dev, sendbuf_dev, nullptr, e->d_send, e->d_items_send[peer], e->np[0], e->np[1]);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_exchange_pack_peer(spb_exchange* e, const double* q_dev, int peer, double* peer_recvbuf_dev, void* stream)
    {
        return spb_exchange_pack(e, q_dev, peer, peer_recvbuf_dev, stream);
    }

    // Stream-ordered flags in (peer) device memory: the sender raises the receiver's flag after the pack kernel that stored the
    // message into the receiver's buffer; the receiver's stream spins on its own flag before it unpacks. Values only grow.
    int spb_flag_signal(unsigned long long* flag_dev, unsigned long long value, void* stream)
    {
        if (!flag_dev) { spb::set_error("spb_flag_signal: null flag"); return SPB_ERR_BAD_ARG; }
        spb::flag_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value);
        SPB_LAUNCH_CHECK();
        return 0;
    }
    int spb_flag_wait(const unsigned long long* flag_dev, unsigned long long value, void* stream)
    {
        if (!flag_dev) { spb::set_error("spb_flag_wait: null flag"); return SPB_ERR_BAD_ARG; }
        spb::flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_exchange_unpack(spb_exchange* e, double* q_dev, int peer, const double* recvbuf_dev, void* stream)
    {
        using namespace spb;
        if (!e || !q_dev || peer < 0 || peer >= e->nranks) { set_error("spb_exchange_unpack: bad argument"); return SPB_ERR_BAD_ARG; }
        int rc = build_device(e); if (rc) return rc;
        const auto& items = e->items_recv[peer];
        if (items.empty()) return 0;
        if (!recvbuf_dev) { set_error("spb_exchange_unpack: null buffer"); return SPB_ERR_BAD_ARG; }
        exchange_kernel<2><<<(unsigned)items.size(), 256, 0, (cudaStream_t)stream>>>(q_dev, nullptr, nullptr, recvbuf_dev, e->d_recv, e->d_items_recv[peer], e->np[0], e->np[1]);
        SPB_LAUNCH_CHECK();
        return 0;
    }
}
