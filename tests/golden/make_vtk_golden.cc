// Generates tests/golden/vtk_small/ (dev container only): the reference's own io::output_vtk (io/io_vtk.h:276-288) and
// io::binary_write (io/io_native.h:40-47) of a 2x1x2 lattice of 8x4x4 blocks whose array holds the exactly representable
// values  q[offset] = (offset % 1013)*0.125 + (offset % 5)  (offset = position in the array's memory), through
// oracle/_ref/libspade_ref.so. A standalone program because the library's std::filesystem calls crash inside a Python
// process that already loaded another libstdc++.
//   g++ -o /tmp/make_vtk_golden tests/golden/make_vtk_golden.cc -Loracle/_ref -lspade_ref -Wl,-rpath,$PWD/oracle/_ref && /tmp/make_vtk_golden tests/golden/vtk_small
#include <cstdio>
#include <vector>
struct ref_cfg { int nblocks[3]; int ncells[3]; int ng; double bounds[6]; int periodic[3]; int scheme; double gamma, R, mu, prandtl, sensor_eps; int nranks; int integrator; double sgs_cw, sgs_delta, sgs_prt; };
extern "C" int ref_output_vtk(const ref_cfg*, const double*, const char*, const char*);
extern "C" long ref_array_size(const ref_cfg*);
extern "C" const char* ref_last_error();
int main(int argc, char** argv)
{
    if (argc < 2) { std::printf("usage: make_vtk_golden out_dir\n"); return 1; }
    ref_cfg c{};
    c.nblocks[0] = 2; c.nblocks[1] = 1; c.nblocks[2] = 2;
    c.ncells[0] = 8; c.ncells[1] = 4; c.ncells[2] = 4;
    c.ng = 2;
    const double hi[3] = {2.0, 1.0, 3.0};
    for (int d = 0; d < 3; ++d) { c.bounds[2*d] = 0.0; c.bounds[2*d + 1] = hi[d]; c.periodic[d] = 1; }
    c.gamma = 1.4; c.R = 287.15; c.mu = 1e-2; c.prandtl = 0.72; c.sensor_eps = 1e-2; c.nranks = 1;
    const long n = ref_array_size(&c);
    std::vector<double> q(n);
    for (long o = 0; o < n; ++o) q[o] = (o % 1013)*0.125 + (o % 5);
    const int rc = ref_output_vtk(&c, q.data(), argv[1], "sol");
    std::printf("rc=%d %s (%ld doubles)\n", rc, ref_last_error(), n);
    return rc;
}
