"""host-side cost of one stage launch (tensor-map encodes etc.): empty block range, so nothing runs on the device"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import spade_b200.api as sp
from spade_b200._lib import StageDesc
torch.cuda.set_device(0)
blocks = sp.cartesian_blocks_t((4, 4, 4), [0.0, 6.283185307179586] * 3)
grid = sp.cartesian_grid_t((32,) * 3, blocks, sp.identity(), sp.pool_t())
gas = sp.ideal_gas_t(1.4, 287.15)
flux = sp.flux_desc(sp.compose(sp.totani_lr(gas), sp.visc_lr(sp.constant_viscosity_t(1e-3, 0.72), gas)))
q = bench.device_state(sp, grid, torch); q2 = q.clone(); k = [sp.grid_array(grid, 0.0) for _ in range(3)]
ex = sp.make_exchange(q, (1, 1, 1))
sd = StageDesc(); sd.nin = 1; sd.inp[0] = k[0].data.data_ptr(); sd.cq[0] = 1e-7; sd.co[0] = 0.5; sd.cq_self, sd.co_self = 1e-7, 1.0; sd.out = k[2].data.data_ptr()
lib = sp.lib()
for name, h in (("stage+ghosts", ex._h), ("stage", None)):
    for rng in ((0, 0), (0, 1)):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(300):
            lib.spb_flux_div_rk_stage_exchange(q.h, C.c_void_p(q.data.data_ptr()), C.c_void_p(q2.data.data_ptr()), C.byref(flux), C.byref(sd), h, rng[0], rng[1], None)
        dt = (time.perf_counter() - t) / 300; torch.cuda.synchronize()
        print(f"{name} range={rng}: {dt*1e6:.1f} us host time per call")
